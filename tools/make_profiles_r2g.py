#!/usr/bin/env python
"""Turn the closing evidence of round 2's last sessions (tools/gpu_evidence_r2g.sh, tools/gpu_r2d.sh, tools/gpu_warm_traffic.sh,
tools/micro_fp64.cu) into tracked files under profiles/:  python tools/make_profiles_r2g.py [tag]
  bench_<tag>_{venice,demo_kf49,ladybug,trafalgar,final_1gpu,venice_2gpu}.json   bench lines
  launches_<tag>_venice.csv + _summary.json                                      ncu launch list of the Venice bench command
  ncu_<tag>_summary.json                                                         ncu --set full digests of the hot kernels
  warm_<tag>_traffic.json                                                        DRAM traffic of the PCG kernels with warm caches
  whatif_<tag>.json                                                              timing experiments on the by-point PCG sweep
  micro_fp64_<tag>.txt                                                           DFMA latency / issue-rate microbenchmark"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2g"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


for s, dst in (("venice", "venice"), ("demo_kf49", "demo_kf49"), ("ladybug", "ladybug"), ("trafalgar", "trafalgar"), ("final", "final_1gpu")):
    src = os.path.join(G, f"bench_{tag}_{s}.json")
    if os.path.exists(src):
        json.dump(last_json(src), open(os.path.join(P, f"bench_{tag}_{dst}.json"), "w"), indent=1)
src = os.path.join(G, "bench_r2f_p2p.json")
if os.path.exists(src):
    json.dump(last_json(src), open(os.path.join(P, f"bench_{tag}_venice_2gpu.json"), "w"), indent=1)

# ---- launch list ----
src = os.path.join(G, f"launches_{tag}_venice.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(P, f"launches_{tag}_venice.csv"), "w", newline="") as f:
        w = csv.writer(f); w.writerow(["id", "kernel", "block", "grid", "ns"])
        for r in rows:
            w.writerow([r[0], r[4].replace("b200ba::", "")[:90], r[7], r[8], r[-1]])
    names = [r[4].split("(")[0].replace("b200ba::", "").replace("void ", "") for r in rows]
    ends = [i for i, n in enumerate(names) if n == "k_lm_decide"]
    summ = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-post",
            "note": "per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes"}
    if len(ends) >= 3:
        a, b = ends[1] + 1, ends[2] + 1
        it = collections.OrderedDict()
        for i in range(a, b):
            x = it.setdefault(names[i], [0, 0.0]); x[0] += 1; x[1] += float(rows[i][-1])
        tot = sum(x[1] for x in it.values())
        summ["one_lm_iteration"] = {"launches": b - a, "total_us": tot / 1e3,
                                    "kernels": {k: {"launches": x[0], "mean_us": x[1] / x[0] / 1e3, "share": x[1] / tot} for k, x in sorted(it.items(), key=lambda kv: -kv[1][1])}}
    json.dump(summ, open(os.path.join(P, f"launches_{tag}_venice_summary.json"), "w"), indent=1)
    print("launch list:", summ.get("one_lm_iteration", {}).get("launches"), "launches per LM iteration")

# ---- ncu --set full digests ----
allrows = []
for rep in (f"ncu_{tag}_pcg", f"ncu_{tag}_hot"):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    tmp = os.path.join("/tmp", rep + ".json")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_digest.py"), path, "6", "--json", tmp], stdout=subprocess.DEVNULL)
    last = collections.OrderedDict()
    for d in json.load(open(tmp)):
        last[d["Kernel Name"]] = d                                # the last captured launch of every kernel (after its warm-up launches)
    for d in last.values():
        try:
            rd, wr = d["dram__bytes_read.sum"].split(), d["dram__bytes_write.sum"].split()
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            d["dram_traffic_bytes_per_launch"] = float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]]
        except Exception:
            pass
        d["capture"] = "ncu --set full --clock-control none --import-source on (cold caches: ncu flushes them before every kernel)"
        allrows.append(d)
if allrows:
    json.dump(allrows, open(os.path.join(P, f"ncu_{tag}_summary.json"), "w"), indent=1)
    print(len(allrows), "kernels digested:", [d["Kernel Name"].split("(")[0] for d in allrows])

# ---- warm-cache DRAM traffic ----
src = os.path.join(G, f"warm_{tag}.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    acc = collections.OrderedDict()
    for r in rows:
        acc.setdefault((int(r[0]), r[4].split("(")[0].replace("b200ba::", "")), {})[r[-3]] = float(r[-1])
    out = {"command": "ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct "
                      "-k regex:'k_cg_point_pass|k_cg_camera' env PROF_FP32=0 python tools/prof_one.py 12   (one pass per kernel, caches NOT flushed: four consecutive PCG steps)",
           "launches": [{"id": i, "kernel": k, **m} for (i, k), m in acc.items()]}
    json.dump(out, open(os.path.join(P, f"warm_{tag}_traffic.json"), "w"), indent=1)

# ---- what-if timings of the by-point PCG sweep (us per launch, tools/prof_cg.py, 20 launches back to back) ----
def grab(log, key="cg_point_pass "):
    res, cur = collections.OrderedDict(), None
    if not os.path.exists(os.path.join(G, log)):
        return res
    for line in open(os.path.join(G, log)):
        if line.startswith("=="):
            cur = line[2:].strip().replace("sequencesfm_b200/variants/", "").replace("sequencesfm_b200/", "")
        elif line.startswith(key) and cur:
            res[cur] = round(1e3 * float(line.split()[-1]), 2)
    return res
whatif = {
    "unit": "us per launch of k_cg_point_pass_tabc at Venice shape (tools/prof_cg.py: 20 launches back to back, CUDA events)",
    "knob": "-DB200BA_WHATIF=<bit mask>: 1 no global loads at the end of a slice, 2 one table load per observation instead of seven, 4 (almost) no arithmetic, "
            "8 never wait for the row stream, 16 no row stream, 32 v never written, 64 no L2 prefetch of the next slice's X / V^-1 (results are WRONG with any of them)",
    "keys": "lib_w<mask>.so = library built with -DB200BA_WHATIF=<mask>; lib_whatif1..4.so = the first encoding of the same knob (1 no slice-end loads, "
            "2 one table load, 3 no arithmetic, 4 never wait for the row stream); 'in-tree' = the library of that call's commit (in the L2-prefetch set: prefetch 16 rows ahead)",
    "before_the_range_balance (contiguous slice ranges per CTA, slice end weighted 2/3 of a row)": {**grab("ab_whatif.log"), **grab("ab_w.log"), **grab("ab_w2.log")},
    "after (ranges dealt round-robin over the CTAs, slice end weighted 4/3)": grab("ab_w3.log"),
    "slice_end_weight_sweep (thirds of a row body)": grab("slicew_x.log"),
    "ranges_round_robin (XPOSE=1) vs contiguous (XPOSE=0); lib_early = slice-end loads requested before the last row body": grab("abx_x.log"),
    "L2_prefetch_of_the_row_stream (rows ahead of the ring)": grab("ab_pf.log"),
    "L2_cache_hints (1: row stream evict_first, 2: V^-1 evict_last)": grab("ab_hint.log"),
    "threads_per_CTA": grab("ab_blk.log"),
    "two_row_tail_block (slice-end loads in flight during the last two row bodies)": grab("ab_e2.log"),
}
json.dump(whatif, open(os.path.join(P, f"whatif_{tag}.json"), "w"), indent=1)
if os.path.exists(os.path.join(G, "micro_fp64.txt")):
    shutil.copy(os.path.join(G, "micro_fp64.txt"), os.path.join(P, f"micro_fp64_{tag}.txt"))
print("done")
