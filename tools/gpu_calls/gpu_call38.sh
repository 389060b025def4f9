#!/bin/bash
# round 2, GPU call 38 (1 GPU): ring depth 4 / 6 / 8 for short (K = 20) and long (K = 240) runs
mkdir -p gpurun_out
for depth in 4 6 8 4 6; do
  for k in 20 240; do
    timeout 300 python bench.py --steps $k --warmup 5 --depth $depth --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('depth $depth K=$k: ms/step %.4f  value %.1f GS/s  e2e %.1f' % (d['ms_per_step'], d['value']/1e3, d['e2e']['value']/1e3))"
  done
done
