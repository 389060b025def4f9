#!/bin/bash
# round 2, GPU call 34 (1 GPU): instruction-footprint experiment -- K3 fast pass and K6 packet loop rolled (FM_ROLLED build) vs unrolled
mkdir -p gpurun_out
R=$PWD/fm_radio_b200/libfmgpu_rolled.so
FMGPU_K3_SINGLE=1 FMGPU_LIB=$R timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "k3_fast or database or device_rds or single_stream or golden" > gpurun_out/pytest_gpu_r3c.log 2>&1; echo "pytest (rolled) exit $?"; tail -2 gpurun_out/pytest_gpu_r3c.log
for lib in "" "$R" "" "$R"; do
  for k in 20 240; do
    FMGPU_K3_SINGLE=1 FMGPU_LIB=$lib timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('[${lib##*/}] K=$k: ms/step %.4f  value %.1f GS/s' % (d['ms_per_step'], d['value']/1e3), 'serial k3 %.3f k6 %.3f' % (d['stage_ms_serial']['k3_pll'], d['stage_ms_serial']['k6_rds']), 'piped k3 %.3f k5 %.3f k6 %.3f' % (d['stage_ms_pipelined']['k3_pll'], d['stage_ms_pipelined']['k5_bpsk'], d['stage_ms_pipelined']['k6_rds']))"
  done
done
FMGPU_K3_SINGLE=1 FMGPU_LIB=$R timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 2>/dev/null | grep '^{' | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('rolled wideband ms/step %.4f' % d['ms_per_step'], d['rds_check'])"
FMGPU_K3_SINGLE=1 timeout 300 python bench.py --workload wideband --steps 48 --warmup 6 2>/dev/null | grep '^{' | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('unrolled wideband ms/step %.4f' % d['ms_per_step'], d['rds_check'])"
