#!/usr/bin/env python
"""Turn the closing evidence run of round 2 (tools/gpu_evidence.sh <tag>) into tracked files under profiles/:
bench lines (Venice + the small shapes), launch-list summaries of the Venice and demo-sequence bench commands, and the
ncu --set full digest of the explicit plan's kernels.   python tools/make_profiles_r2b.py <tag>"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2b"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for s in ("venice", "demo_kf49", "ladybug", "trafalgar"):
    line = [l for l in open(os.path.join(G, f"bench_{tag}_{s}.json")).read().splitlines() if l.startswith("{")][-1]
    json.dump(json.loads(line), open(os.path.join(P, f"bench_{tag}_{s}.json"), "w"), indent=1)
for s, cmd, end in (("venice", "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-post", "k_lm_decide"),
                    ("demo", "python bench.py --shape demo_kf49 --steps 4 --warmup 3 --no-cpu-baseline --no-fp32 --no-post", "k_lm_decide")):
    rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}_{s}.csv"))) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(P, f"launches_{tag}_{s}.csv"), "w", newline="") as f:
        w = csv.writer(f); w.writerow(["id", "kernel", "block", "grid", "ns"])
        for r in rows:
            w.writerow([r[0], r[4].replace("b200ba::", "")[:90], r[7], r[8], r[-1]])
    names = [r[4].split("(")[0].replace("b200ba::", "").replace("void ", "") for r in rows]
    ends = [i for i, n in enumerate(names) if n == end]
    summ = {"command": f"ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 {cmd}",
            "note": "per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes"}
    if len(ends) >= 3:
        a, b = ends[1] + 1, ends[2] + 1
        it = collections.OrderedDict()
        for i in range(a, b):
            x = it.setdefault(names[i], [0, 0.0]); x[0] += 1; x[1] += float(rows[i][-1])
        tot = sum(x[1] for x in it.values())
        summ["one_lm_iteration"] = {"launches": b - a, "total_us": tot / 1e3,
                                    "kernels": {k: {"launches": x[0], "mean_us": x[1] / x[0] / 1e3, "share": x[1] / tot} for k, x in sorted(it.items(), key=lambda kv: -kv[1][1])}}
    json.dump(summ, open(os.path.join(P, f"launches_{tag}_{s}_summary.json"), "w"), indent=1)
    print(s, summ.get("one_lm_iteration", {}).get("launches"), "launches per iteration,", round(summ.get("one_lm_iteration", {}).get("total_us", 0), 1), "us under ncu")
allrows = []
for rep, what in ((f"ncu_{tag}_dense_demo", "demo sequence (49 cameras)"), (f"ncu_{tag}_dense_traf", "Trafalgar shape (257 cameras)")):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    tmp = os.path.join("/tmp", rep + ".json")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_digest.py"), path, "6", "--json", tmp], stdout=subprocess.DEVNULL)
    for d in json.load(open(tmp)):
        d["problem"] = what
        allrows.append(d)
json.dump(allrows, open(os.path.join(P, f"ncu_{tag}_dense_summary.json"), "w"), indent=1)
print(len(allrows), "ncu --set full launches digested")
