#!/bin/bash
# round 2, GPU call 32 (1 GPU): K3 with a helper warp (k3_pll_duo)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -s > gpurun_out/pytest_gpu_r3b.log 2>&1; echo "pytest exit $?"; grep -E "K3 fast|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r3b.log | head
for cfg in "" "FMGPU_K3_SINGLE=1"; do
  for k in 20 240; do
    env $cfg timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('[$cfg] K=$k: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3), 'serial', {k: round(v, 3) for k, v in d['stage_ms_serial'].items()}, 'piped k3 %.3f k5 %.3f' % (d['stage_ms_pipelined']['k3_pll'], d['stage_ms_pipelined']['k5_bpsk']))"
  done
done
timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 2>/dev/null | grep '^{' | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('wideband ms/step %.4f' % d['ms_per_step'], 'chan ms %.4f' % d['roofline']['ms_per_launch'], d['rds_check'])"
