#!/bin/bash
# round 2, GPU call 1: K1 tensor-core probe (descriptor variants), prime-mode bisection, channelizer profile
mkdir -p gpurun_out
cd tools
for v in "0 0" "1 0" "1 1"; do
  echo "=== k1t_probe $v" >> ../gpurun_out/k1t_probe.log
  timeout 180 ./k1t_probe $v >> ../gpurun_out/k1t_probe.log 2>&1
  echo "exit $?" >> ../gpurun_out/k1t_probe.log
done
cd ..
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi1.log 2>&1
for m in none full nodestroy gctx gctx_launch alloc; do
  echo "=== prime mode $m" >> gpurun_out/prime_modes.log
  FMGPU_PRIME_MODE=$m timeout 300 python tools/bisect_bench.py "plain_$m" >> gpurun_out/prime_modes.log 2>&1
done
echo "=== lazy-loading off, no prime" >> gpurun_out/prime_modes.log
CUDA_MODULE_LOADING=EAGER FMGPU_PRIME_MODE=none timeout 300 python tools/bisect_bench.py "eager_none" >> gpurun_out/prime_modes.log 2>&1
# channelizer: step breakdown (launch list) and full-set capture of the two channelizer kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_chan_launches.csv python tools/chan_profile.py tensor 4 > gpurun_out/chan_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chan_" -s 2 -c 1 -f -o gpurun_out/r2a_chan_mma python tools/chan_profile.py tensor 2 > gpurun_out/chan_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chan_" -s 2 -c 1 -f -o gpurun_out/r2a_chan_fp32 python tools/chan_profile.py fp32 2 > gpurun_out/chan_ncu2.log 2>&1
python tools/chan_profile.py tensor 24 > gpurun_out/chan_time.log 2>&1
python tools/chan_profile.py fp32 24 >> gpurun_out/chan_time.log 2>&1
ls -la gpurun_out | tail -20
tail -40 gpurun_out/k1t_probe.log
cat gpurun_out/prime_modes.log | grep -v "^$" | tail -20
