#!/bin/bash
# Refresh of the closing evidence after the last kernel changes (k_schur_diag_cameras, k_linearize_cameras): every GPU test, the default
# bench line, the ncu launch list of the bench command, ncu --set full of the two kernels.  Usage: tools/gpu_evidence_r2h.sh <tag>
tag=${1:-r2g}; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
timeout 1200 python bench.py > gpurun_out/bench_${tag}_venice.json 2> gpurun_out/bench_${tag}_venice.err; echo "venice rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}_venice.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-post > gpurun_out/ncu_${tag}_venice.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lin_tab|k_linearize_cameras|k_schur_diag' -c 12 -o gpurun_out/ncu_${tag}_hot -f env PROF_FP32=0 python tools/prof_one.py 0,1,6 > gpurun_out/ncu_${tag}_hot.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_${tag}_venice.json").read().splitlines() if l.startswith("{")][-1])
print("ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), [round(x, 4) for x in d["e2e"]["seconds_total_each_call"]], "set_problem", d["e2e"]["ms_set_problem"],
      "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["iteration"]["frac"], 3), "parity", d["parity"]["max_rel_cost_diff"], d["parity"]["pass"], "launches", d["gpu_launches"], d["clocks"])
print({k: round(v["frac"], 3) for k, v in d["roofline"]["per_kernel"].items()}, "fp32", d["pcg_fp32"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
PY
