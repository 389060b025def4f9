#!/bin/bash
# Final single-GPU leg of a round: full GPU suite, smoke, bench lines, ncu captures.   usage: bash tools/gpu_final.sh [TAG]
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"
grep -E "config 3 sample|config 1 vs|K3 fast|passed|failed|FAILED" gpurun_out/pytest_gpu_$TAG.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --steps 240 --warmup 6 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python - <<PY
import json
for ln in open('gpurun_out/bench_$TAG.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print(' value %.1f GS/s  ms/step %.4f  e2e %.1f GS/s  cpu %.0f MS/s' % (d['value']/1e3, d['ms_per_step'], d['e2e']['value']/1e3, (d.get('cpu_baseline') or {}).get('value') or 0))
        print('  roofline', d['roofline']['kernel'], d['roofline']['bound'], round(d['roofline']['achieved'], 1), d['roofline']['unit'], 'frac %.3f' % d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
        print('  serial', {k: round(v, 4) for k, v in d['stage_ms_serial'].items()})
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_20steps.json 2>/dev/null; python -c "
import json
d = json.loads([l for l in open('gpurun_out/bench_${TAG}_20steps.json') if l.startswith('{')][-1]); print(' 20-step run: value %.1f GS/s ms/step %.4f' % (d['value']/1e3, d['ms_per_step']))"
timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 > gpurun_out/bench_wideband_$TAG.json 2>/dev/null
grep '^{' gpurun_out/bench_wideband_$TAG.json | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(' wideband ms/step %.4f' % d['ms_per_step'], 'chan ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['rds_check'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>/dev/null; grep '^{' gpurun_out/bench_reference_$TAG.json | cut -c1-400
bash tools/ncu_capture.sh $TAG > gpurun_out/ncu_capture_$TAG.log 2>&1; tail -1 gpurun_out/ncu_capture_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chan_mma" -s 2 -c 1 -f -o gpurun_out/${TAG}_chan_mma python tools/chan_profile.py tensor 2 > gpurun_out/chan_ncu_$TAG.log 2>&1
ncu -i gpurun_out/${TAG}_chan_mma.ncu-rep --page raw --csv > gpurun_out/${TAG}_chan_mma_raw.csv 2>/dev/null
FMGPU_NO_PARTITION=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(k[1-7]_|k4b|chan_)" -s 40 -c 120 --csv --log-file gpurun_out/launches_wideband_$TAG.csv python tools/chan_profile.py tensor 12 > gpurun_out/chan_launches_$TAG.log 2>&1; tail -1 gpurun_out/chan_launches_$TAG.log
