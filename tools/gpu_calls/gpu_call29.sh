#!/bin/bash
# round 2, GPU call 29 (1 GPU): K2 with the next chunk's inputs prefetched by cp.async mid-chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x > gpurun_out/pytest_gpu_r3a.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r3a.log
for i in 1 2; do timeout 300 python bench.py --steps 240 --warmup 6 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('ms/step %.4f  value %.1f GS/s  e2e %.1f' % (d['ms_per_step'], d['value']/1e3, d['e2e']['value']/1e3), {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})"; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('20 steps: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3))"
