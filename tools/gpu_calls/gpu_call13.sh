#!/bin/bash
# round 2, GPU call 13 (2 GPUs): i8 MMA rate microbenchmark, config 5 wideband at 2 ranks, 2-rank test
mkdir -p gpurun_out
(cd tools && timeout 120 ./umma_rate > ../gpurun_out/r2_umma_rate.txt 2>&1); cat gpurun_out/r2_umma_rate.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --workload wideband --gpus 2 --steps 48 --warmup 6 > gpurun_out/r2_wideband_n2.log 2>&1; echo "wideband n=2 exit $?"
grep '^{' gpurun_out/r2_wideband_n2.log | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(' n_gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.0f MS/s' % d['value'], 'e2e %.0f' % d['e2e']['value'], d['rds_check'], 'bcast', d['config']['broadcast_bytes_per_step'])"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_gpu_multi2.log 2>&1; tail -3 gpurun_out/pytest_gpu_multi2.log
