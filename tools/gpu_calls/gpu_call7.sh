#!/bin/bash
# round 2, GPU call 7 (1 GPU): K4 balanced FIR roles: tests + bench A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest exit $?"
grep -E "config 3 sample|config 1 vs|K3 fast|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2g.log | head
for cfg in "default" "FMGPU_K4_V1=1"; do
  echo "=== bench $cfg" >> gpurun_out/bench_r2g.log
  env $(echo $cfg | sed 's/default//') timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline >> gpurun_out/bench_r2g.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2g.log'):
    if ln.startswith('==='): print(ln.strip())
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3))
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
PY
