#!/bin/bash
# Closing evidence of round 2 (last sessions) on one GPU box: every GPU test, the bench lines (Venice default, the small shapes, Final on
# one GPU), the ncu launch list of the Venice bench command, ncu --set full of the hot kernels, warm-cache DRAM traffic of the PCG kernels.
# Usage: tools/gpu_evidence_r2g.sh <tag>
tag=${1:-r2g}; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
timeout 1200 python bench.py > gpurun_out/bench_${tag}_venice.json 2> gpurun_out/bench_${tag}_venice.err; echo "venice rc=$?"
for s in demo_kf49 ladybug trafalgar; do
  timeout 600 python bench.py --shape $s --steps 16 --warmup 3 --no-post > gpurun_out/bench_${tag}_$s.json 2> gpurun_out/bench_${tag}_$s.err; echo "$s rc=$?"
done
timeout 900 python bench.py --shape final --steps 8 --warmup 3 --no-post --no-cpu-baseline > gpurun_out/bench_${tag}_final.json 2> gpurun_out/bench_${tag}_final.err; echo "final rc=$?"
python - <<PY
import json
for s in ("venice", "demo_kf49", "ladybug", "trafalgar", "final"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_${tag}_{s}.json").read().splitlines() if l.startswith("{")][-1])
        print(s, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), [round(x, 4) for x in d["e2e"]["seconds_total_each_call"]],
              "roof", round(d["roofline"]["frac"], 3), "parity", (d["parity"] or {}).get("max_rel_cost_diff"), (d["parity"] or {}).get("pass"), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        print("    cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k in ("value", "cores", "kind")}, "plan", d["config"].get("plan", {}).get("reduced_system", "")[:60])
    except Exception as e:
        print(s, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}_venice.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-post > gpurun_out/ncu_${tag}_venice.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cg_point_pass_tabc|k_cg_camera_pass|k_cg_camera_step|k_lin_tab|k_linearize_cameras|k_schur_diag|k_backsub_tab|k_cost_tab' -c 14 -o gpurun_out/ncu_${tag}_hot -f env PROF_FP32=0 python tools/prof_one.py 0,1,6,8,12 > gpurun_out/ncu_${tag}_hot.log 2>&1; echo "ncu full rc=$?"
bash tools/gpu_warm_traffic.sh $tag | tail -6
ls -la gpurun_out/*${tag}* | head -30
