#!/bin/bash
# round 2, GPU call 9 (1 GPU): rerun after the stream-alias fix: tests, bench A/B (K4 taps), config-2 sweep, prime-mode bisection
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "pytest exit $?"
grep -E "config 3 sample|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2i.log | head
for cfg in "default" "FMGPU_K4_V1=1"; do
  echo "=== bench $cfg" >> gpurun_out/bench_r2i.log
  env $(echo $cfg | sed 's/default//') timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline >> gpurun_out/bench_r2i.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2i.log'):
    if ln.startswith('==='): print(ln.strip())
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3))
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
PY
echo "=== config 2 sweep, graph replay (default)"; timeout 600 python -c "
import sys; sys.path.insert(0, 'tools'); sys.argv=['x']
import config_sweeps as c, json
rows = c.config2(blocks=(4096, 16384, 65536, 262144, 1048576))
json.dump(rows, open('gpurun_out/config2_graph.json', 'w'), indent=1)
" 2>&1 | grep config2
echo "=== config 2 sweep, FMGPU_NO_GRAPH=1"; FMGPU_NO_GRAPH=1 timeout 600 python -c "
import sys; sys.path.insert(0, 'tools'); sys.argv=['x']
import config_sweeps as c, json
rows = c.config2(blocks=(4096, 16384))
json.dump(rows, open('gpurun_out/config2_nograph.json', 'w'), indent=1)
" 2>&1 | grep config2
for m in none full nodestroy gctx gctx_alloc gctx_events; do
  echo "=== prime mode $m" >> gpurun_out/prime_modes3.log
  FMGPU_PRIME_MODE=$m timeout 300 python tools/bisect_bench.py "plain_$m" >> gpurun_out/prime_modes3.log 2>&1
done
grep -v "^$" gpurun_out/prime_modes3.log | grep -v "==="
