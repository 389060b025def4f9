#!/bin/bash
# round 2, GPU call 30 (1 GPU): graded stream priorities on the FIR partition (later stages first), K = 20 and K = 240
mkdir -p gpurun_out
for cfg in "" "FMGPU_FIR_PRIO=1" "" "FMGPU_FIR_PRIO=1"; do
  for k in 20 240; do
    env $cfg timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('[$cfg] K=$k: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3), {k: round(v, 3) for k, v in d['stage_ms_pipelined'].items()})"
  done
done
