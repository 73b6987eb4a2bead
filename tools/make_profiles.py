#!/usr/bin/env python
"""Turn one evidence round (tools/gpu_profile.sh <tag>) into the tracked files under profiles/.

    python tools/make_profiles.py <tag> <name>      e.g.  r2i r1_final

writes profiles/bench_<name>.json, profiles/launches_<name>.csv (+ _summary.json: per-kernel launch count, mean and share
of the LM iteration) and profiles/ncu_<name>_summary.json (last captured launch of every kernel, key metrics,
dram traffic per launch)."""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")

# bench line
line = [l for l in open(os.path.join(G, f"bench_{tag}.json")).read().splitlines() if l.startswith("{")][-1]
json.dump(json.loads(line), open(os.path.join(P, f"bench_{name}.json"), "w"), indent=1)

# launch list of the same command (ncu --metrics gpu__time_duration.sum): cold-cache, serialised
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 10 and r[0].isdigit()]
with open(os.path.join(P, f"launches_{name}.csv"), "w", newline="") as f:
    w = csv.writer(f); w.writerow(["id", "kernel", "block", "grid", "ns"])
    for r in rows:
        w.writerow([r[0], r[4].replace("b200ba::", ""), r[7], r[8], r[-1]])
agg = collections.OrderedDict()
for r in rows:
    k = r[4].split("(")[0].replace("b200ba::", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
tot = sum(a[1] for a in agg.values())
summ = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
        "note": "first 1200 launches of the run (set-up + the first LM iterations); per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
        "total_us": tot / 1e3,
        "kernels": {k: {"launches": a[0], "mean_us": a[1] / a[0] / 1e3, "share": a[1] / tot} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
# one full LM iteration = the launches between two consecutive k_lm_decide (the list also holds set-up and bench.py's
# per-kernel timing section, so the share that has to agree with bench.py's `share_of_step` is taken inside an iteration)
names = [r[4].split("(")[0].replace("b200ba::", "") for r in rows]
ends = [i for i, n in enumerate(names) if n == "k_lm_decide"]
cg = None
if len(ends) >= 3:
    a, b = ends[1] + 1, ends[2] + 1
    it = collections.OrderedDict()
    for i in range(a, b):
        x = it.setdefault(names[i], [0, 0.0]); x[0] += 1; x[1] += float(rows[i][-1])
    it_tot = sum(x[1] for x in it.values())
    cg = sum(x[1] for k, x in it.items() if k in ("k_cg_point_pass_tab", "k_cg_point_pass_tabc", "k_cg_point_pass_tab2", "k_cg_camera_pass", "k_cg_camera_step")) / it_tot
    summ["one_lm_iteration"] = {"launches": b - a, "total_us": it_tot / 1e3, "pcg_step_share": cg,
                                "kernels": {k: {"launches": x[0], "mean_us": x[1] / x[0] / 1e3, "share": x[1] / it_tot} for k, x in sorted(it.items(), key=lambda kv: -kv[1][1])}}
json.dump(summ, open(os.path.join(P, f"launches_{name}_summary.json"), "w"), indent=1)

# ncu --set full digests (one report, or the two the evidence script writes to stay under gpurun's 64 MiB limit)
reps = [os.path.join(G, f"prof_{tag}{suf}.ncu-rep") for suf in ("", "_sweeps", "_rest")]
allrows = []
for rep in reps:
    if not os.path.exists(rep):
        continue
    tmp = os.path.join("/tmp", os.path.basename(rep) + ".json")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_digest.py"), rep, "4", "--json", tmp], stdout=subprocess.DEVNULL)
    allrows += json.load(open(tmp))
last = collections.OrderedDict()
for d in allrows:
    last[d["Kernel Name"]] = d
out = []
for k, d in last.items():
    rd = float(d["dram__bytes_read.sum"].split()[0]); wr = float(d["dram__bytes_write.sum"].split()[0])
    unit = d["dram__bytes_read.sum"].split()[1]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[unit]
    wunit = d["dram__bytes_write.sum"].split()[1]
    wscale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[wunit]
    d["dram_traffic_bytes_per_launch"] = rd * scale + wr * wscale
    out.append(d)
json.dump(out, open(os.path.join(P, f"ncu_{name}_summary.json"), "w"), indent=1)
print("pcg step share inside one LM iteration of the launch list:", cg)
for d in out:
    print(d["Kernel Name"][:40].ljust(40), d["gpu__time_duration.sum"], round(d["dram_traffic_bytes_per_launch"] / 1e6, 1), "MB")
