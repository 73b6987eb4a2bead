#!/bin/bash
# GPU-box call for the explicit reduced-system plan: plan-vs-plan check, step slopes, ncu launch list (warm caches).
# Usage: tools/gpu_dense.sh <tag> [check] [slope] [ncu] [pytest]
tag=${1:-x}; shift
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0,'.')
from sequencesfm_b200 import binding
from sequencesfm_b200.synth import make_problem, load_problem_npz
name=sys.argv[1]
p = load_problem_npz('tests/golden/demo_kf49.npz') if name=="demo" else make_problem(name,1)
opt=binding.default_options(max_iterations=6,min_iterations=10**6,gradient_threshold=0.0,poll_every=64)
with binding.Engine(opt) as e:
    e.set_problem_from(p); e.begin(); e.iterate(5)
PY
for step in "$@"; do
case $step in
check) timeout 600 python tools/dense_check.py > gpurun_out/dense_check_$tag.log 2>&1; python - <<PY
import json
for line in open("gpurun_out/dense_check_$tag.log"):
    try: name, js = line.split(" ", 1); r = json.loads(js)
    except Exception: print(line.rstrip()); continue
    print(name, "plan", r["fixed50"]["plan"], "fixed50 diff8 %.1e all %.1e" % (r["fixed50"]["max_rel_cost_diff_first8"], r["fixed50"]["max_rel_cost_diff"]),
          "exact diff8 %.1e all %.1e" % (r["exact"]["max_rel_cost_diff_first8"], r["exact"]["max_rel_cost_diff"]), "accept_eq", r["fixed50"]["accept_equal"], r["exact"]["accept_equal"],
          "ms/it", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r["ms_per_lm_iteration"].items()})
PY
;;
slope) timeout 300 python tools/step_slope.py 0 demo ladybug trafalgar > gpurun_out/slope_dense_$tag.log 2>&1; cat gpurun_out/slope_dense_$tag.log ;;
ncu)
for s in demo trafalgar; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_dense_${s}_$tag.csv python /tmp/one.py $s > gpurun_out/ncu_dense_$s.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_dense_${s}_$tag.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    if r[ui]=="ns": v/=1000
    elif r[ui]=="ms": v*=1000
    k=r[ki].split("(")[0]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print("$s total us", round(tot,1))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]: print(f"  {k[:60]:60s} n={a[0]:4d} avg={a[1]/a[0]:8.2f} us  sum={a[1]:9.1f}")
PY
done ;;
pytest) ( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_$tag.log; tail -5 gpurun_out/pytest_$tag.log ;;
esac
done
