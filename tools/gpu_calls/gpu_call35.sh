#!/bin/bash
# round 2, GPU call 35 (1 GPU): K3 helper-warp version selected automatically for small grids; full suite; bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r3d.log 2>&1; echo "pytest exit $?"; grep -E "K3 fast|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r3d.log | head
for k in 20 240; do
  timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('K=$k: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3), 'serial', {k: round(v, 3) for k, v in d['stage_ms_serial'].items()})"
done
for cfg in "" "FMGPU_K3_SINGLE=1"; do
env $cfg timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 2>/dev/null | grep '^{' | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('[$cfg] wideband ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], d['rds_check'])"
done
for cfg in "" "FMGPU_K3_SINGLE=1"; do
echo "=== config 2 sweep [$cfg]"; env $cfg timeout 600 python -c "
import sys; sys.path.insert(0, 'tools'); sys.argv=['x']
import config_sweeps as c, json
rows = c.config2(blocks=(4096, 16384, 65536, 262144, 1048576))
json.dump(rows, open('gpurun_out/config2_r3d_${cfg:-default}.json', 'w'), indent=1)
" 2>&1 | grep config2
done
