#!/bin/bash
# round 2, GPU call 3: K1T v2 (4-deep ring, 256 threads) + K5 v2 (branch-free, look-ahead window): tests, bench, ncu
mkdir -p gpurun_out
(cd tools && timeout 300 ./k1t_probe > ../gpurun_out/k1t_probe3.log 2>&1; echo "exit $?" >> ../gpurun_out/k1t_probe3.log)
grep -E "PASS|FAIL|timing|exit" gpurun_out/k1t_probe3.log
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/pytest_gpu_r2c.log
timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline > gpurun_out/bench_r2c.log 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2c.log'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3))
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
PY
bash tools/ncu_capture.sh r2c > gpurun_out/ncu_capture_r2c.log 2>&1
tail -3 gpurun_out/ncu_capture_r2c.log
