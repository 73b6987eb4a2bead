#!/bin/bash
# Evidence round: bench line, ncu launch list of the SAME bench command, ncu --set full of every hot kernel.
# Usage: tools/gpu_profile.sh <tag>
tag=$1; mkdir -p gpurun_out
timeout 900 python bench.py --steps 16 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo bench failed; tail -20 gpurun_out/bench_$tag.err; exit 1; }
tail -c 600 gpurun_out/bench_$tag.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
wc -l gpurun_out/launches_$tag.csv
# (a) the four PCG sweeps (fp64 + fp32) with source-level stall data; (b) every other hot kernel, metrics only.
# Two reports so that what comes back stays under gpurun's 64 MiB limit (one --set full launch with sources is ~1.7 MB).
timeout 300 python tools/prof_one.py 2,3,14,15 > gpurun_out/prof_plain_$tag.log 2>&1 || { echo plain failed; cat gpurun_out/prof_plain_$tag.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cg_point_pass_tabc|k_cg_camera_pass' -c 16 -o gpurun_out/prof_${tag}_sweeps -f \
    python tools/prof_one.py 2,3,14,15 > gpurun_out/prof_ncu_${tag}_sweeps.log 2>&1
tail -2 gpurun_out/prof_ncu_${tag}_sweeps.log
timeout 900 ncu --set full --clock-control none -k regex:'k_linearize|k_lin_tab|k_schur|k_backsub_tab|k_cost_tab|k_cg_camera_step' -s 2 -c 24 -o gpurun_out/prof_${tag}_rest -f \
    python tools/prof_one.py 0,1,6,8,11 > gpurun_out/prof_ncu_${tag}_rest.log 2>&1
tail -2 gpurun_out/prof_ncu_${tag}_rest.log
ls -la gpurun_out/
