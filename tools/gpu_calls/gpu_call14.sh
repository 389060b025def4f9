#!/bin/bash
# round 2, GPU call 14 (1 GPU): channelizer v2b (N = 192 MMAs, branch-free epilogue, 2 epilogue warpgroups) + full suite + final captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_channelizer.py -m gpu -q -x > gpurun_out/pytest_gpu_r2m_chan.log 2>&1; echo "chan tests exit $?"; tail -3 gpurun_out/pytest_gpu_r2m_chan.log
timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 > gpurun_out/bench_wideband_r2m.log 2>&1
grep '^{' gpurun_out/bench_wideband_r2m.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('wideband ms/step %.4f' % d['ms_per_step'], 'chan ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['rds_check'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chan_mma" -s 2 -c 1 -f -o gpurun_out/r2m_chan_mma python tools/chan_profile.py tensor 2 > gpurun_out/chan_ncu_r2m.log 2>&1
ncu -i gpurun_out/r2m_chan_mma.ncu-rep --page raw --csv > gpurun_out/r2m_chan_mma_raw.csv 2>/dev/null
python tools/summarize_ncu.py gpurun_out/r2m_chan_mma_raw.csv | head -24
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(k[1-5]_|k6_rds$|k7_|k4b|chan_)" --csv --log-file gpurun_out/r2m_chan_launches.csv python tools/chan_profile.py tensor 4 > gpurun_out/chan_launches_r2m.log 2>&1; tail -2 gpurun_out/chan_launches_r2m.log
timeout 1500 python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu_r2m.log 2>&1; echo "pytest exit $?"
grep -E "config 3 sample|config 1 vs|K3 fast|passed|failed|FAILED" gpurun_out/pytest_gpu_r2m.log | head
timeout 900 python bench.py --steps 240 --warmup 6 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2m.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s  cpu %.0f MS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3, (d.get('cpu_baseline') or {}).get('value') or 0))
        print('  roofline', d['roofline']['kernel'], d['roofline']['bound'], round(d['roofline']['achieved'], 1), d['roofline']['unit'], 'frac %.3f' % d['roofline']['frac'])
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2m_20steps.json 2>/dev/null; python -c "
import json
d = json.loads([l for l in open('gpurun_out/bench_r2m_20steps.json') if l.startswith('{')][-1]); print(' 20-step run (the driver\'s): value %.1f GS/s ms/step %.4f' % (d['value']/1e3, d['ms_per_step']))"
bash tools/ncu_capture.sh r2m > gpurun_out/ncu_capture_r2m.log 2>&1
tail -2 gpurun_out/ncu_capture_r2m.log
