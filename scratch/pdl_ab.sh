#!/bin/bash
# A/B of programmatic dependent launch inside ONE lease: LB_PDL=0 (fully serialised launches) vs the default.
mkdir -p gpurun_out
FAST="--no-cpu-baseline --no-gpu-eager --no-attention --no-parity"
for b in 512 64 16; do
  for pdl in 0 1 0 1; do
    LB_PDL=$pdl timeout 200 python bench.py --batch $b --steps 10 --warmup 3 $FAST 2> gpurun_out/pdl_ab_b${b}_p${pdl}.err | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print('batch $b pdl $pdl', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms', 'e2e', round(d['e2e']['value'], 1))
"
  done
done
