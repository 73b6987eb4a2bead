#!/bin/bash
# One GPU-box call of round 2: parity tests, variant timings, stamps, bench line.  Usage: tools/gpu_call.sh <tag> [steps...]
tag=${1:-x}; shift
mkdir -p gpurun_out
for step in "$@"; do
case $step in
pytest)
  ( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 ) > gpurun_out/pytest_$tag.log; tail -12 gpurun_out/pytest_$tag.log ;;
variants)
  bash tools/gpu_variants.sh $tag ;;
ilp2test)
  ( B200BA_TAB_ILP=2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "linearisation or venice_shape_matches or ragged or fifty or symmetric" 2>&1 | tail -5 ) > gpurun_out/pytest_ilp2_$tag.log; tail -3 gpurun_out/pytest_ilp2_$tag.log ;;
stamps2)
  B200BA_LIB=$PWD/sequencesfm_b200/variants/lib_stamps2.so timeout 300 python tools/stamps.py venice > gpurun_out/stamps2_$tag.log 2>&1; cat gpurun_out/stamps2_$tag.log; cp gpurun_out/stamps_venice.npy gpurun_out/stamps2_venice.npy ;;
stamps)
  B200BA_LIB=$PWD/sequencesfm_b200/variants/lib_stamps.so timeout 300 python tools/stamps.py venice > gpurun_out/stamps_$tag.log 2>&1; cat gpurun_out/stamps_$tag.log ;;
bench)
  timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_total_each_call"], "roof", d["roofline"]["frac"], d["roofline"]["iteration"]["frac"], d["clocks"])
    print("parity", {k: v for k, v in d["parity"].items() if k not in ("against",)})
    print("config", d["config"])
    print({k: round(v, 4) for k, v in d["roofline"]["kernels_ms"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-3000:])
PY
  ;;
benchfull)
  timeout 1200 python bench.py > gpurun_out/benchfull_$tag.json 2> gpurun_out/benchfull_$tag.err; tail -c 1500 gpurun_out/benchfull_$tag.json ;;
*) echo "unknown step $step" ;;
esac
done
