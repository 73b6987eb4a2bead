#!/bin/bash
# One GPU-box round: parity tests, per-kernel timings, bench line.  Usage: tools/gpu_round.sh <tag> [pytest-args]
tag=${1:-x}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
timeout 300 python tools/prof_cg.py venice > gpurun_out/kern_$tag.log 2>&1
cat gpurun_out/kern_$tag.log
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["iteration"]["frac"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-2000:])
PY
