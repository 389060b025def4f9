#!/bin/bash
# round 2, GPU call 19 (1 GPU): channelizer epilogue with rotation table + device RDS database extension test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_channelizer.py -m gpu -q -x > gpurun_out/pytest_gpu_r2s_chan.log 2>&1; echo "chan tests exit $?"; tail -3 gpurun_out/pytest_gpu_r2s_chan.log
timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 > gpurun_out/bench_wideband_r2s.log 2>&1
grep '^{' gpurun_out/bench_wideband_r2s.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('wideband ms/step %.4f' % d['ms_per_step'], 'chan ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['rds_check'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chan_mma" -s 2 -c 1 -f -o gpurun_out/r2s_chan_mma python tools/chan_profile.py tensor 2 > gpurun_out/chan_ncu_r2s.log 2>&1
ncu -i gpurun_out/r2s_chan_mma.ncu-rep --page raw --csv > gpurun_out/r2s_chan_mma_raw.csv 2>/dev/null
python tools/summarize_ncu.py gpurun_out/r2s_chan_mma_raw.csv | head -28
