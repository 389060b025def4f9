#!/bin/bash
# How the step time depends on the number of distinct input blocks kept in HBM (L2 residue of the u8 input).
for nb in 4 8 24; do
  python bench.py --no-cpu-baseline --input-blocks $nb 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('input_blocks', $nb, round(d['value']), round(d['ms_per_step'], 4), round(d['stage_ms_pipelined']['k1_fir4_discrim'], 4), d['rds_check']['streams_with_own_pi_decoded_on_device'])"
done
