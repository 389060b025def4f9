#!/bin/bash
# round 2, GPU call 5 (1 GPU): wideband bench workload, h2d ceiling, config-3 sample test, multi-rank test (skips)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_multi.py -m gpu -q -s -k "config3 or wideband" > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest exit $?"
grep -E "config 3 sample|passed|failed|FAILED|Error|skipped" gpurun_out/pytest_gpu_r2e.log | head
timeout 600 python bench.py --workload wideband --steps 48 --warmup 6 > gpurun_out/bench_wideband_n1.log 2>&1; echo "wideband exit $?"; tail -c 2500 gpurun_out/bench_wideband_n1.log
timeout 300 python tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_n1.log 2>&1; cat gpurun_out/h2d_ceiling_n1.log | tail -2
timeout 600 python bench.py --steps 60 --warmup 6 --no-cpu-baseline > gpurun_out/bench_r2e.log 2>&1; python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2e.log'):
    if ln.startswith('{'):
        d = json.loads(ln); print('value %.1f GS/s  ms/step %.4f e2e %.1f GS/s d2h %d' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3, d['e2e']['d2h_bytes_per_step']))
PY
