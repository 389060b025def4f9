#!/bin/bash
# round 2, GPU call 40 (1 GPU): last validation of the final library (priorities removed) + default bench (ring depth 3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r3e.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_r3e.log
timeout 600 python bench.py --steps 240 --warmup 6 > gpurun_out/bench_r3e.json 2>/dev/null; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r3e.json') if l.startswith('{')][-1])
print(' K=240: value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s  cpu %.0f MS/s depth %d' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3, (d.get('cpu_baseline') or {}).get('value') or 0, d['config']['pipeline_depth']))
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r3e_20steps.json 2>/dev/null; python -c "
import json
d = json.loads([l for l in open('gpurun_out/bench_r3e_20steps.json') if l.startswith('{')][-1]); print(' K=20: value %.1f GS/s ms/step %.4f' % (d['value']/1e3, d['ms_per_step']))"
