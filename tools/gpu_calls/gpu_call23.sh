#!/bin/bash
# round 2, GPU call 23 (1 GPU): compute-sanitizer over small shapes of every kernel + partition sweep with the round-2 kernels
mkdir -p gpurun_out
timeout 120 python tools/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain smoke exit $?"; tail -2 gpurun_out/sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|sanitize_smoke" gpurun_out/sanitize_memcheck.log | head -12
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck exit $?"; grep -E "ERROR SUMMARY|Barrier|sanitize_smoke" gpurun_out/sanitize_synccheck.log | head -8
for cfg in "" "FMGPU_NO_PARTITION=1" "FMGPU_RECURRENCE_SMS=12" "FMGPU_RECURRENCE_SMS=20"; do
  env $cfg timeout 300 python bench.py --steps 120 --warmup 6 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('partition [$cfg]: ms/step %.4f  value %.1f GS/s  e2e %.1f' % (d['ms_per_step'], d['value']/1e3, d['e2e']['value']/1e3), {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})"
done
