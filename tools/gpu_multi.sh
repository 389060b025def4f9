#!/bin/bash
# Multi-GPU leg on one box (run under `gpurun --gpus N`): config 5 wideband broadcast, host-ingest ceiling and the streams bench
# at N ranks; at N = 2 also the 2-rank NCCL test.   usage: bash tools/gpu_multi.sh N [TAG]
N=${1:-2}; TAG=${2:-r2z}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NCCL_DEBUG=$([ $N = 8 ] && echo INFO || echo WARN) timeout 240 $TR --nproc-per-node $N --master-port $((29600 + N)) bench.py --workload wideband --gpus $N --steps 48 --warmup 6 \
    > gpurun_out/${TAG}_wideband_n$N.log 2>&1; echo "wideband n=$N exit $?"
grep '^{' gpurun_out/${TAG}_wideband_n$N.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(' n_gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.0f MS/s' % d['value'], 'e2e %.0f' % d['e2e']['value'], d['rds_check'], 'bcast', d['config']['broadcast_bytes_per_step'])"
if [ $N = 8 ]; then grep -E "NCCL INFO (Channel|comm|Connected|ncclCommInitRank|Init)|NVLS|P2P" gpurun_out/${TAG}_wideband_n8.log | head -40 > gpurun_out/${TAG}_wideband_n8_nccl.txt; fi
timeout 200 $TR --nproc-per-node $N --master-port $((29700 + N)) tools/h2d_ceiling.py > gpurun_out/${TAG}_h2d_ceiling_n$N.log 2>&1
grep '^{' gpurun_out/${TAG}_h2d_ceiling_n$N.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(' N', d['world_size'], 'h2d-only %.1f GB/s' % d['h2d_only']['h2d_GBps_aggregate'], ' h2d+d2h %.1f + %.1f GB/s' % (d['h2d_plus_d2h']['h2d_GBps_aggregate'], d['h2d_plus_d2h']['d2h_GBps_aggregate']), ' ceiling %.0f MS/s' % d['h2d_plus_d2h']['iq_MSps_ceiling'])"
timeout 300 $TR --nproc-per-node $N --master-port $((29800 + N)) bench.py --gpus $N --steps 60 --warmup 6 --no-cpu-baseline > gpurun_out/${TAG}_streams_n$N.log 2>&1; echo "streams n=$N exit $?"
grep '^{' gpurun_out/${TAG}_streams_n$N.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(' n_gpus', d['n_gpus'], 'value %.0f MS/s' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f MS/s' % d['e2e']['value'], 'd2h', d['e2e']['d2h_bytes_per_step'])"
if [ $N = 2 ]; then timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_pytest_gpu_multi.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu_multi.log; fi
