#!/bin/bash
# round 2, GPU call 12 (1 GPU): pipelined channelizer: tests, A/B timing, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_channelizer.py -m gpu -q -x > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "chan tests exit $?"; tail -4 gpurun_out/pytest_gpu_r2l.log
timeout 120 python tools/chan_profile.py tensor 48 > gpurun_out/chan_time_r2l.log 2>&1
FMGPU_CHAN_V1=1 timeout 120 python tools/chan_profile.py tensor 48 >> gpurun_out/chan_time_r2l.log 2>&1
cat gpurun_out/chan_time_r2l.log
timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 > gpurun_out/bench_wideband_r2l.log 2>&1
grep '^{' gpurun_out/bench_wideband_r2l.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('wideband ms/step %.4f' % d['ms_per_step'], 'chan ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['rds_check'])"
FMGPU_CHAN_V1=1 timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 > gpurun_out/bench_wideband_r2l_v1.log 2>&1
grep '^{' gpurun_out/bench_wideband_r2l_v1.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('wideband (v1 kernel) ms/step %.4f' % d['ms_per_step'], 'chan ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['rds_check'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chan_mma" -s 2 -c 1 -f -o gpurun_out/r2l_chan_mma python tools/chan_profile.py tensor 2 > gpurun_out/chan_ncu_r2l.log 2>&1
ncu -i gpurun_out/r2l_chan_mma.ncu-rep --page raw --csv > gpurun_out/r2l_chan_mma_raw.csv 2>/dev/null
python tools/summarize_ncu.py gpurun_out/r2l_chan_mma_raw.csv | head -24
grep -E "sm__pipe_tensor_cycles_active|imma_cycles" gpurun_out/r2l_chan_mma_raw.csv | head -0
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2l_chan_mma_raw.csv')))
for i, h in enumerate(rows[0]):
    if 'pipe_tensor' in h and 'realtime' in h: print(h, rows[1][i], rows[2][i])
PY
