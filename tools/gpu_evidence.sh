#!/bin/bash
# Round-2 closing evidence on one GPU box: parity tests, bench lines (Venice default + the small shapes), ncu launch lists of the
# same commands, ncu --set full captures of the explicit plan's kernels.  Usage: tools/gpu_evidence.sh <tag>
tag=${1:-r2b}; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
timeout 1200 python bench.py > gpurun_out/bench_${tag}_venice.json 2> gpurun_out/bench_${tag}_venice.err; echo "venice rc=$?"
for s in demo_kf49 ladybug trafalgar; do
  timeout 600 python bench.py --shape $s --steps 16 --warmup 3 --no-post > gpurun_out/bench_${tag}_$s.json 2> gpurun_out/bench_${tag}_$s.err; echo "$s rc=$?"
done
python - <<PY
import json
for s in ("venice", "demo_kf49", "ladybug", "trafalgar"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_${tag}_{s}.json").read().splitlines() if l.startswith("{")][-1])
        print(s, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), [round(x, 4) for x in d["e2e"]["seconds_total_each_call"]],
              "roof", round(d["roofline"]["frac"], 3), "parity", (d["parity"] or {}).get("max_rel_cost_diff"), (d["parity"] or {}).get("pass"), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        print("    cpu", {k: v for k, v in (d["cpu_baseline"] or {}).items() if k in ("value", "cores", "kind")}, "plan", d["config"].get("plan", {}).get("reduced_system", "")[:60])
    except Exception as e:
        print(s, "failed", e)
PY
# launch lists (ncu, warm caches are NOT kept: per-launch times are cold-cache and serialised - shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}_venice.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-post > gpurun_out/ncu_${tag}_venice.log 2>&1; echo "ncu venice rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}_demo.csv python bench.py --shape demo_kf49 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32 --no-post > gpurun_out/ncu_${tag}_demo.log 2>&1; echo "ncu demo rc=$?"
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0,'.')
from sequencesfm_b200 import binding
from sequencesfm_b200.synth import make_problem, load_problem_npz
name=sys.argv[1]
p = load_problem_npz('tests/golden/demo_kf49.npz') if name=="demo" else make_problem(name,1)
opt=binding.default_options(max_iterations=6,min_iterations=10**6,gradient_threshold=0.0,poll_every=64)
with binding.Engine(opt) as e:
    e.set_problem_from(p); e.begin(); e.iterate(4)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dense|k_precond|k_camera_reduce_unpack' -s 8 -c 8 -o gpurun_out/ncu_${tag}_dense_demo -f python /tmp/one.py demo > gpurun_out/ncu_${tag}_dense_demo.log 2>&1; echo "ncu full demo rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dense' -s 4 -c 4 -o gpurun_out/ncu_${tag}_dense_traf -f python /tmp/one.py trafalgar > gpurun_out/ncu_${tag}_dense_traf.log 2>&1; echo "ncu full trafalgar rc=$?"
ls -la gpurun_out/*${tag}*
