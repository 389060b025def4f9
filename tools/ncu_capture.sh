#!/bin/bash
# Run on the GPU box (under gpurun): launch list of a short bench + one full-set capture of each hot kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k[1-7]_|k4b|chan_" -s 24 -c 160 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline --clock-warmup-ms 0 --input-blocks 4 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k[1-7]_" -s 21 -c 7 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --clock-warmup-ms 0 --input-blocks 4 > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ls -la gpurun_out
