#!/usr/bin/env python
"""Digest an .ncu-rep: key raw metrics per launch + the instructions carrying most stall samples.
Usage: tools/ncu_digest.py report.ncu-rep [top_n] [--json out.json]"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 14
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    summ = []
    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        d = {"Kernel Name": name}
        for w in WANT:
            if w in ix:
                d[w] = f"{r[ix[w]]} {units[ix[w]]}".strip()
        summ.append(d)
        if name in seen:
            continue
        seen.add(name)
        print("=====", name)
        for w in WANT:
            if w in d:
                print(f"  {w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')} = {d[w]}")
    if out_json:
        json.dump(summ, open(out_json, "w"), indent=1)
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"] + [len(src)]
    done = set()
    for a, b in zip(secs[:-1], secs[1:]):
        name = src[a][1]
        if name in done:
            continue
        done.add(name)
        h = src[a + 1]; ix = {x: i for i, x in enumerate(h)}
        body = src[a + 2:b]
        tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
        print(f"----- {name}: {tot} samples, {len(body)} SASS instructions")
        order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:top]
        keys = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        for i in sorted(order):
            r = body[i]
            st = {k[6:]: int(r[ix[k]]) for k in keys if r[ix[k]] not in ("", "0")}
            print(f"  {i:4d} {r[ix['Source']].strip()[:64]:64s} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  x{r[ix['Instructions Executed']]:>8s} {st}")


if __name__ == "__main__":
    main()
