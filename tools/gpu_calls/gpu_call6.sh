#!/bin/bash
# round 2, GPU call 6 (1 GPU): K4 v2 (persistent producer/consumer) validation + A/B, config-3 sample test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "k4_persistent or config3" > gpurun_out/pytest_gpu_r2f_k4.log 2>&1; echo "k4 tests exit $?"
grep -E "config 3 sample|passed|failed|FAILED|Error|Timeout" gpurun_out/pytest_gpu_r2f_k4.log | head
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/pytest_gpu_r2f.log
for cfg in "default" "FMGPU_K4_V1=1"; do
  echo "=== bench $cfg" >> gpurun_out/bench_r2f.log
  env $(echo $cfg | sed 's/default//') timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline >> gpurun_out/bench_r2f.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2f.log'):
    if ln.startswith('==='): print(ln.strip())
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3))
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
PY
bash tools/ncu_capture.sh r2f > gpurun_out/ncu_capture_r2f.log 2>&1
tail -2 gpurun_out/ncu_capture_r2f.log
