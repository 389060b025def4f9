#!/bin/bash
# round 2, GPU call 4: K1T v3 (lean loop), K5 v3 (smem ring), K3 fast pass: probes, tests, bench A/B, ncu
mkdir -p gpurun_out
(cd tools && timeout 300 ./k1t_probe > ../gpurun_out/k1t_probe4.log 2>&1; echo "exit $?" >> ../gpurun_out/k1t_probe4.log; timeout 300 ./k3_probe > ../gpurun_out/k3_probe4.log 2>&1)
grep -E "PASS|FAIL|timing|exit" gpurun_out/k1t_probe4.log; cat gpurun_out/k3_probe4.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 -s > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest exit $?"
grep -E "config 1 vs|config 3 sample|K3 fast|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2d.log | head -20
for cfg in "default" "FMGPU_K3_EXACT=1"; do
  echo "=== bench $cfg" >> gpurun_out/bench_r2d.log
  env $(echo $cfg | sed 's/default//') timeout 600 python bench.py --steps 120 --warmup 6 --no-cpu-baseline >> gpurun_out/bench_r2d.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2d.log'):
    if ln.startswith('==='): print(ln.strip())
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3))
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
        print('  piped ', {k: round(v, 4) for k, v in d['stage_ms_pipelined'].items()})
PY
bash tools/ncu_capture.sh r2d > gpurun_out/ncu_capture_r2d.log 2>&1
tail -2 gpurun_out/ncu_capture_r2d.log
