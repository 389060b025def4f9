#!/bin/bash
# Does the nvidia-smi clock sampler (or the clock warm-up) perturb the timed region?
for cfg in "25 500" "200 500" "5000 500" "25 0"; do
  set -- $cfg
  BENCH_SMI_MS=$1 python bench.py --no-cpu-baseline --clock-warmup-ms $2 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('smi_ms', $1, 'warm_ms', $2, round(d['value']), round(d['ms_per_step'], 4), d['clocks'])"
done
