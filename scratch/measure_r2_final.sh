#!/bin/bash
# Round-2 final measurements (one lease): full GPU test suite, default bench (with baselines), the other BASELINE configs.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_final.log
timeout 300 python bench.py > gpurun_out/r2f_bench_b512.json 2> gpurun_out/r2f_bench_b512.err; echo "bench rc=$?"
F="--no-cpu-baseline --no-attention --no-parity"
timeout 120 python bench.py --batch 16 $F > gpurun_out/r2f_bench_b16.json 2>> gpurun_out/r2f_bench.err
timeout 120 python bench.py --batch 64 $F > gpurun_out/r2f_bench_b64.json 2>> gpurun_out/r2f_bench.err
F="--no-cpu-baseline --no-attention --no-parity --no-gpu-eager"
timeout 120 python bench.py --res 64 $F > gpurun_out/r2f_bench_res64_b512.json 2>> gpurun_out/r2f_bench.err
timeout 120 python bench.py --res 32 $F > gpurun_out/r2f_bench_res32_b512.json 2>> gpurun_out/r2f_bench.err
timeout 200 python bench.py --res 256 --depth 3 --batch 64 $F > gpurun_out/r2f_bench_res256_d3_b64.json 2>> gpurun_out/r2f_bench.err
for f in gpurun_out/r2f_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],1), d.get('hbm_peak_gib'))"; done
