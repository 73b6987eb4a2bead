#!/bin/bash
# DRAM traffic of the PCG kernels with WARM caches: one-pass ncu (two DRAM counters + duration), no cache flush between kernels,
# on four consecutive PCG steps.  Usage: tools/gpu_warm_traffic.sh <tag>   (B200BA_LIB selects a variant library)
tag=${1:-x}; mkdir -p gpurun_out
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:'k_cg_point_pass|k_cg_camera' --csv --log-file gpurun_out/warm_$tag.csv env PROF_FP32=0 python tools/prof_one.py 12 > gpurun_out/warm_$tag.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/warm_$tag.csv")) if len(r) > 10 and r[0].isdigit()]
acc = collections.OrderedDict()
for r in rows:
    acc.setdefault((r[0], r[4].split("(")[0]), {})[r[-3]] = (r[-1], r[-2])
for (i, k), m in acc.items():
    print(i, k.replace("b200ba::", ""), {a: b for a, b in m.items()})
PY
