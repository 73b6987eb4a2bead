#!/bin/bash
# ncu --set full (with sources) of the by-point table sweep, one-row and two-row variants.  Usage: tools/gpu_ncu2.sh <tag>
tag=$1; mkdir -p gpurun_out
timeout 300 python tools/prof_one.py 2 > gpurun_out/prof_plain_$tag.log 2>&1 || { echo plain run failed; cat gpurun_out/prof_plain_$tag.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_point_pass_tab -s 2 -c 1 -o gpurun_out/ncu_${tag}_ilp1 -f python tools/prof_one.py 2 > gpurun_out/ncu_${tag}_ilp1.log 2>&1
B200BA_TAB_ILP=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_point_pass_tab -s 2 -c 1 -o gpurun_out/ncu_${tag}_ilp2 -f python tools/prof_one.py 2 > gpurun_out/ncu_${tag}_ilp2.log 2>&1
tail -2 gpurun_out/ncu_${tag}_ilp1.log gpurun_out/ncu_${tag}_ilp2.log
ls -la gpurun_out/*.ncu-rep
