#!/bin/bash
# Round-2 (closing sessions) GPU call: A/B of the per-kernel timings (variants/lib_*.so against the in-tree library), parity
# tests, bench line.  Usage: tools/gpu_r2d.sh <tag> [steps...]   steps: ab pytest parity bench benchfull launches
tag=${1:-r2d}; shift
mkdir -p gpurun_out
for step in "$@"; do
case $step in
ab)
  echo "== in-tree" > gpurun_out/ab_$tag.log
  timeout 300 python tools/prof_cg.py venice >> gpurun_out/ab_$tag.log 2>&1
  for so in sequencesfm_b200/variants/lib_*.so; do
    [ -f "$so" ] || continue
    echo "== $so" >> gpurun_out/ab_$tag.log
    B200BA_LIB=$PWD/$so timeout 300 python tools/prof_cg.py venice >> gpurun_out/ab_$tag.log 2>&1
  done
  cat gpurun_out/ab_$tag.log ;;
abx)
  : > gpurun_out/abx_$tag.log
  for x in 0 1; do for so in sequencesfm_b200/libb200ba.so sequencesfm_b200/variants/lib_*.so; do
    [ -f "$so" ] || continue
    echo "== XPOSE=$x $so" >> gpurun_out/abx_$tag.log
    B200BA_XPOSE=$x B200BA_LIB=$PWD/$so timeout 300 python tools/prof_cg.py venice >> gpurun_out/abx_$tag.log 2>&1
  done; done
  grep "==\|linearize_points\|cg_point_pass \|backsub\|cg_step\|rror" gpurun_out/abx_$tag.log ;;
slicew)
  : > gpurun_out/slicew_$tag.log
  for x in 0 1 2 3 4 5 6 8; do
    echo "== SLICE_W=$x/3" >> gpurun_out/slicew_$tag.log
    B200BA_SLICE_W=$x timeout 300 python tools/prof_cg.py venice >> gpurun_out/slicew_$tag.log 2>&1
  done
  grep "==\|linearize_points\|cg_point_pass \|backsub\|rror" gpurun_out/slicew_$tag.log ;;
parity)
  ( timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --durations=12 2>&1 | tail -30 ) > gpurun_out/pytest_parity_$tag.log; tail -22 gpurun_out/pytest_parity_$tag.log ;;
pytest)
  ( timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -30 ) > gpurun_out/pytest_$tag.log; tail -22 gpurun_out/pytest_$tag.log ;;
bench)
  timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-post > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_total_each_call"], "roof", d["roofline"]["frac"], d["roofline"]["iteration"]["frac"], d["clocks"])
    print("parity", {k: v for k, v in d["parity"].items() if k not in ("against", "cost_per_iteration")})
    print({k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel"].items()})
    print({k: round(v, 4) for k, v in d["roofline"].get("kernels_ms", {}).items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-3000:])
PY
  ;;
benchfull)
  timeout 1200 python bench.py > gpurun_out/benchfull_$tag.json 2> gpurun_out/benchfull_$tag.err; echo "benchfull rc=$?"; tail -c 600 gpurun_out/benchfull_$tag.json ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}_venice.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-post > gpurun_out/ncu_${tag}_venice.log 2>&1; echo "ncu launches rc=$?" ;;
*) echo "unknown step $step" ;;
esac
done
