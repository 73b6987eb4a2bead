#!/bin/bash
# Final-shape (13 682 / 4 456 117 / 28 987 644) bench lines at the rank counts given.  Usage: tools/gpu_final.sh <tag> <N> [<N> ...]
tag=$1; shift; mkdir -p gpurun_out
port=29600
for N in "$@"; do
  port=$((port+1))
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --shape final --steps 8 --warmup 3 --no-cpu-baseline --no-fp32 --no-post > gpurun_out/bench_final_${tag}_$N.json 2> gpurun_out/bench_final_${tag}_$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --shape final --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_final_${tag}_$N.json 2> gpurun_out/bench_final_${tag}_$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_final_${tag}_$N.json").read().splitlines() if l.startswith("{")][-1])
    print("final n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "cg_step", d["roofline"]["kernels_ms"].get("cg_step"), "parity", d["parity"].get("max_rel_cost_diff"), d["parity"].get("pass"), "roof", round(d["roofline"]["frac"],3))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_final_${tag}_$N.err").read()[-1500:])
PY
done
