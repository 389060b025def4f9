#!/bin/bash
# round 2, GPU call 31 (1 GPU): K3 on SMs of its own inside the recurrence partition (FMGPU_K3_SMS), K = 20 and K = 240
mkdir -p gpurun_out
for cfg in "" "FMGPU_K3_SMS=8" "FMGPU_RECURRENCE_SMS=24 FMGPU_K3_SMS=8" "FMGPU_RECURRENCE_SMS=24 FMGPU_K3_SMS=16" "FMGPU_K3_SMS=8 FMGPU_FIR_PRIO=1"; do
  for k in 20 240; do
    env $cfg timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline 2>gpurun_out/err31.log | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('[$cfg] K=$k: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3), {k: round(v, 3) for k, v in d['stage_ms_pipelined'].items()})"
  done
  grep -m1 "fmgpu: K3" gpurun_out/err31.log
done
