#!/bin/bash
# round 2, GPU call 24 (1 GPU): K1T with the next-but-one tile staged mid-epilogue
mkdir -p gpurun_out
timeout 120 tools/k1t_probe > gpurun_out/k1t_probe24.log 2>&1; echo "k1t_probe exit $?"; grep -E "PASS|FAIL|timing|worst" gpurun_out/k1t_probe24.log | tail -8
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "k1 or single_stream or golden or small_blocks or batched or cf32 or graph" > gpurun_out/pytest_gpu_r2w.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r2w.log
timeout 300 python bench.py --steps 240 --warmup 6 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('ms/step %.4f  value %.1f GS/s  e2e %.1f' % (d['ms_per_step'], d['value']/1e3, d['e2e']['value']/1e3), {k: round(v, 4) for k, v in d['stage_ms_serial'].items()}); print('roofline frac %.3f' % d['roofline']['frac'])"
