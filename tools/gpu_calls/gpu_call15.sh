#!/bin/bash
# round 2, GPU call 15 (1 GPU): K1T epilogue trim validation + full suite + final captures (tag r2n)
mkdir -p gpurun_out
if [ -x tools/k1t_probe ]; then timeout 120 tools/k1t_probe > gpurun_out/k1t_probe15.log 2>&1; echo "k1t_probe exit $?"; tail -4 gpurun_out/k1t_probe15.log; fi
timeout 1500 python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu_r2n.log 2>&1; echo "pytest exit $?"
grep -E "config 3 sample|config 1 vs|K3 fast|passed|failed|FAILED" gpurun_out/pytest_gpu_r2n.log | head
timeout 900 python bench.py --steps 240 --warmup 6 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2n.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s  cpu %.0f MS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3, (d.get('cpu_baseline') or {}).get('value') or 0))
        print('  roofline', d['roofline']['kernel'], d['roofline']['bound'], round(d['roofline']['achieved'], 1), d['roofline']['unit'], 'frac %.3f' % d['roofline']['frac'])
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2n_20steps.json 2>/dev/null; python -c "
import json
d = json.loads([l for l in open('gpurun_out/bench_r2n_20steps.json') if l.startswith('{')][-1]); print(' 20-step run: value %.1f GS/s ms/step %.4f' % (d['value']/1e3, d['ms_per_step']))"
bash tools/ncu_capture.sh r2n > gpurun_out/ncu_capture_r2n.log 2>&1
tail -2 gpurun_out/ncu_capture_r2n.log
# config-4 (wideband) launch list; green-context kernels cannot be profiled, so partition off
FMGPU_NO_PARTITION=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(k[1-7]_|k4b|chan_)" -s 40 -c 120 --csv --log-file gpurun_out/launches_wideband_r2n.csv python tools/chan_profile.py tensor 6 > gpurun_out/chan_launches_r2n.log 2>&1; tail -2 gpurun_out/chan_launches_r2n.log
