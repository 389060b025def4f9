#!/bin/bash
# round 2, GPU call 39 (1 GPU): ring depth 3 (less run-ahead of the first stages), with and without graded FIR priorities
mkdir -p gpurun_out
for cfg in "3:" "3:FMGPU_FIR_PRIO=1" "4:" "4:FMGPU_FIR_PRIO=1" "3:"; do
  depth=${cfg%%:*}; envs=${cfg#*:}
  for k in 20 240; do
    env $envs timeout 300 python bench.py --steps $k --warmup 5 --depth $depth --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('depth $depth [$envs] K=$k: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3))"
  done
done
