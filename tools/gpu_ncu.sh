#!/bin/bash
# ncu --set full of selected kernel groups (indices of b200ba_kernel_name) on the Venice shape.
# Usage: tools/gpu_ncu.sh <tag> <indices e.g. 2,3> [kernel regex]
tag=$1; idx=$2; rx=${3:-k_cg}
mkdir -p gpurun_out
timeout 300 python tools/prof_one.py $idx > gpurun_out/prof_plain_$tag.log 2>&1 || { echo plain run failed; cat gpurun_out/prof_plain_$tag.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s ${4:-2} -c ${5:-2} -o gpurun_out/prof_$tag -f python tools/prof_one.py $idx > gpurun_out/prof_ncu_$tag.log 2>&1
tail -5 gpurun_out/prof_ncu_$tag.log
