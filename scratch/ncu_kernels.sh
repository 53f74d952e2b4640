#!/bin/bash
# One ncu --set full pass over the step's kernels of interest (first launches of each, program order).  usage: bash scratch/ncu_kernels.sh <tag>
tag=${1:-r2}
python scratch/one_step.py 512 > gpurun_out/${tag}_one_step.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_conv_tc2|k_wgrad_halo|k_wgrad_tc|k_wgrad_reduce|k_gate_fwd_stats|k_softmax_pixels_fwd|k_norm_apply4|k_gate_bwd4" \
    -c 260 -f -o gpurun_out/${tag}_step_kernels python scratch/one_step.py 512 > gpurun_out/${tag}_ncu_step.log 2>&1
tail -2 gpurun_out/${tag}_ncu_step.log
