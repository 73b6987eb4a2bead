#!/bin/bash
# Multi-GPU round on an N-GPU box: parity of the point-partitioned solve (peer windows and NCCL) + bench lines.
# Usage: tools/gpu_multi.sh <tag> <N>
tag=$1; N=${2:-2}; mkdir -p gpurun_out
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== p2p ladybug" > gpurun_out/multi_$tag.log
run 29511 tools/multi_gpu_check.py ladybug 12 >> gpurun_out/multi_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/multi_$tag.log
echo "== nccl ladybug" >> gpurun_out/multi_$tag.log
B200BA_NO_P2P=1 run 29512 tools/multi_gpu_check.py ladybug 12 >> gpurun_out/multi_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/multi_$tag.log
echo "== p2p venice" >> gpurun_out/multi_$tag.log
run 29513 tools/multi_gpu_check.py venice 6 >> gpurun_out/multi_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/multi_$tag.log
grep "==\|^{\|rc=\|Error\|error" gpurun_out/multi_$tag.log | cut -c1-600
run 29514 bench.py --gpus $N --steps 16 --warmup 3 > gpurun_out/bench_${tag}_p2p.json 2> gpurun_out/bench_${tag}_p2p.err; echo "bench p2p rc=$?"
B200BA_NO_P2P=1 run 29515 bench.py --gpus $N --steps 16 --warmup 3 > gpurun_out/bench_${tag}_nccl.json 2> gpurun_out/bench_${tag}_nccl.err; echo "bench nccl rc=$?"
python - <<PY
import json
for k in ("p2p","nccl"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_${tag}_%s.json"%k).read().splitlines() if l.startswith("{")][-1])
        print(k, "n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "cg_step", d["roofline"]["kernels_ms"].get("cg_step"))
    except Exception as e:
        print(k, "failed", e); print(open("gpurun_out/bench_${tag}_%s.err"%k).read()[-1500:])
PY
