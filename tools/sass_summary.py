#!/usr/bin/env python
"""SASS opcode histogram of the kernels of sequencesfm_b200/libb200ba.so -> profiles/sass_<name>_summary.json.
Shows which async-copy / barrier / wide-load instructions each hot kernel really contains (UBLKCP = TMA bulk copy, LDGSTS =
cp.async, SYNCS = mbarrier, LDG.E.*.256, UBLKPF = bulk L2 prefetch, STAS = st.async into a cluster peer's shared memory,
UCGABAR = barrier.cluster, PREEXIT / ACQBULK = griddepcontrol.launch_dependents / .wait) and its FP64 instruction count.  Usage: tools/sass_summary.py <name>"""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "sequencesfm_b200", "libb200ba.so")], capture_output=True, text=True).stdout
res = collections.OrderedDict(); cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("b200ba::", "").replace("void ", "")
        cur = res.setdefault(fn, collections.Counter()); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
summ = {}
for fn, c in res.items():
    if not fn.startswith("k_"):
        continue
    grp = collections.Counter()
    for op, n in c.items():
        grp[op.split(".")[0]] += n
    summ[fn] = {"instructions": sum(c.values()), "fp64": sum(n for op, n in grp.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")),
                "tma_bulk_copy_UBLKCP": grp.get("UBLKCP", 0), "bulk_prefetch_UBLKPF": grp.get("UBLKPF", 0), "cp_async_LDGSTS": grp.get("LDGSTS", 0),
                "mbarrier_SYNCS": grp.get("SYNCS", 0), "LDG_256": sum(n for op, n in c.items() if op.startswith("LDG") and ".256" in op),
                "STG_256": sum(n for op, n in c.items() if op.startswith("STG") and ".256" in op), "LDS": grp.get("LDS", 0),
                "st_async_STAS": grp.get("STAS", 0), "cluster_barrier_UCGABAR": grp.get("UCGABAR_ARV", 0) + grp.get("UCGABAR_WAIT", 0),
                "griddepcontrol_PREEXIT_ACQBULK": grp.get("PREEXIT", 0) + grp.get("ACQBULK", 0), "atomics_ATOM_RED": grp.get("ATOM", 0) + grp.get("ATOMG", 0) + grp.get("RED", 0) + grp.get("ATOMS", 0),
                "top_opcodes": dict(grp.most_common(10))}
json.dump(summ, open(os.path.join(ROOT, "profiles", f"sass_{name}_summary.json"), "w"), indent=1)
for fn in ("k_cg_point_pass_tabc", "k_cg_camera_pass", "k_cg_camera_step", "k_lin_tab", "k_backsub_tab", "k_cost_tab", "k_linearize_cameras"):
    if fn in summ: print(fn, {k: v for k, v in summ[fn].items() if k != "top_opcodes"})
