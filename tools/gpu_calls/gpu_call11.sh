#!/bin/bash
# round 2, GPU call 11 (1 GPU): K1T 3-CTA shape, recurrence-partition sweep, first-handle effect with the FP32 K1
mkdir -p gpurun_out
(cd tools && timeout 300 ./k1t_probe > ../gpurun_out/k1t_probe11.log 2>&1; echo "exit $?" >> ../gpurun_out/k1t_probe11.log)
grep -E "PASS|FAIL|timing|exit" gpurun_out/k1t_probe11.log
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "k1_tensor or cuda_graph" > gpurun_out/pytest_gpu_r2k.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2k.log
FMGPU_K1T_SHAPE=1 timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "k1_tensor" > gpurun_out/pytest_gpu_r2k_shape1.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2k_shape1.log
for cfg in "default" "FMGPU_K1T_SHAPE=1" "FMGPU_RECURRENCE_SMS=12" "FMGPU_RECURRENCE_SMS=8" "FMGPU_RECURRENCE_SMS=20"; do
  echo "=== bench $cfg" >> gpurun_out/bench_r2k.log
  env $(echo $cfg | sed 's/default//') timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline >> gpurun_out/bench_r2k.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2k.log'):
    if ln.startswith('==='): print(ln.strip())
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3), d['config']['sm_partition'])
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
PY
for m in none full; do
  echo "=== K1 FP32, prime mode $m" >> gpurun_out/prime_modes4.log
  FMGPU_K1_FP32=1 FMGPU_PRIME_MODE=$m timeout 300 python tools/bisect_bench.py "k1fp32_$m" >> gpurun_out/prime_modes4.log 2>&1
done
grep -v "^$" gpurun_out/prime_modes4.log | grep -v "==="
