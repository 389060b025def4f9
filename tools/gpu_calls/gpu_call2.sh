#!/bin/bash
# round 2, GPU call 2: K1T timing, full GPU test-suite with K1T + symbol-wise K5, bench A/B, stream-connection experiment
mkdir -p gpurun_out
(cd tools && timeout 300 ./k1t_probe 1 0 > ../gpurun_out/k1t_probe2.log 2>&1; echo "exit $?" >> ../gpurun_out/k1t_probe2.log)
grep -E "PASS|FAIL|timing|exit" gpurun_out/k1t_probe2.log
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/pytest_gpu_r2b.log
for cfg in "default" "FMGPU_K1_FP32=1" "FMGPU_K5_LITERAL=1" "FMGPU_K1_FP32=1 FMGPU_K5_LITERAL=1"; do
  echo "=== bench $cfg" >> gpurun_out/bench_r2b.log
  env $(echo $cfg | sed 's/default//') timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline >> gpurun_out/bench_r2b.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2b.log'):
    if ln.startswith('==='): print(ln.strip())
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3))
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
        print('  rds', d['rds_check'])
PY
for m in none full; do for c in 8 32; do
  echo "=== prime $m connections $c" >> gpurun_out/prime_modes2.log
  CUDA_DEVICE_MAX_CONNECTIONS=$c FMGPU_PRIME_MODE=$m timeout 300 python tools/bisect_bench.py "plain_${m}_c$c" >> gpurun_out/prime_modes2.log 2>&1
done; done
grep -v "^$" gpurun_out/prime_modes2.log
