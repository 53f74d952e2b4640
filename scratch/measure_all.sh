#!/bin/bash
# One gpurun call: the round's benchmark points (every JSON line lands in gpurun_out/).  usage: bash scratch/measure_all.sh <tag>
tag=${1:-r2}
out=gpurun_out
run() { name=$1; shift; python bench.py "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; echo "== $name rc=$?"; python scratch/bench_summary.py $out/${tag}_$name.json 0 2>&1 | head -3; }
run bench_b512 --steps 10 --warmup 3 --shapes-file $out/${tag}_shapes_b512.txt
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "== reference arm rc=$?"; tail -c 400 $out/${tag}_bench_reference_arm.json
run bench_b16 --batch 16 --eager-batch 16 --steps 20 --warmup 3 --no-attention --no-cpu-baseline
run bench_b64 --batch 64 --steps 20 --warmup 3 --no-attention --no-cpu-baseline
run bench_res64_b512 --res 64 --batch 512 --eager-batch 64 --steps 10 --warmup 3 --no-cpu-baseline
run bench_res32_b512 --res 32 --batch 512 --eager-batch 64 --steps 10 --warmup 3 --no-attention --no-cpu-baseline
run bench_res256_d3_b64 --res 256 --depth 3 --batch 64 --eager-batch 8 --attention-batch 16 --steps 5 --warmup 3 --no-cpu-baseline
