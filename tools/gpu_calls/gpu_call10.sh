#!/bin/bash
# round 2, GPU call 10 (8 GPUs): config 5 wideband broadcast at 2/4/8 ranks, host-ingest ceiling at 1/2/4/8, streams bench at 8, 2-rank test
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  NCCL_DEBUG=$([ $n = 8 ] && echo INFO || echo WARN) timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --workload wideband --gpus $n --steps 48 --warmup 6 \
      > gpurun_out/r2_wideband_n$n.log 2>&1; echo "wideband n=$n exit $?"
  grep '^{' gpurun_out/r2_wideband_n$n.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(' n_gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.0f MS/s' % d['value'], 'e2e %.0f' % d['e2e']['value'], d['rds_check'], 'bcast', d['config']['broadcast_bytes_per_step'])"
done
grep -E "NCCL INFO (Channel|comm|Connected|ncclCommInitRank|Init)|NVLS|P2P" gpurun_out/r2_wideband_n8.log | head -40 > gpurun_out/r2_wideband_n8_nccl.txt; wc -l gpurun_out/r2_wideband_n8_nccl.txt
timeout 120 python tools/h2d_ceiling.py > gpurun_out/r2_h2d_ceiling.log 2>&1
for n in 2 4 8; do timeout 200 $TR --nproc-per-node $n --master-port $((29700 + n)) tools/h2d_ceiling.py >> gpurun_out/r2_h2d_ceiling.log 2>&1; done
grep '^{' gpurun_out/r2_h2d_ceiling.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(' N', d['world_size'], 'h2d-only %.1f GB/s' % d['h2d_only']['h2d_GBps_aggregate'], ' h2d+d2h %.1f + %.1f GB/s' % (d['h2d_plus_d2h']['h2d_GBps_aggregate'], d['h2d_plus_d2h']['d2h_GBps_aggregate']), ' ceiling %.0f MS/s' % d['h2d_plus_d2h']['iq_MSps_ceiling'])"
for n in 8 4 2; do
  timeout 400 $TR --nproc-per-node $n --master-port $((29800 + n)) bench.py --gpus $n --steps 60 --warmup 6 --no-cpu-baseline > gpurun_out/r2_streams_n$n.log 2>&1; echo "streams n=$n exit $?"
  grep '^{' gpurun_out/r2_streams_n$n.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(' n_gpus', d['n_gpus'], 'value %.0f MS/s' % d['value'], 'e2e %.0f MS/s' % d['e2e']['value'], 'd2h', d['e2e']['d2h_bytes_per_step'])"
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_gpu_multi.log 2>&1; tail -3 gpurun_out/pytest_gpu_multi.log
