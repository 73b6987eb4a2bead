"""Explicit reduced camera system (options.reduced_system) against the matrix-free plan on the small shapes:
per-iteration cost of both plans, their difference, and the time per LM iteration (fixed 50 PCG steps and exact-step mode).

    python tools/dense_check.py [shape ...]        shapes: small demo ladybug trafalgar (default: all)
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencesfm_b200 import binding
from sequencesfm_b200.synth import make_problem, load_problem_npz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def problem(name):
    if name == "demo":
        return load_problem_npz(os.path.join(ROOT, "tests/golden/demo_kf49.npz"))
    return make_problem(name, 1, outlier_frac=0.02)


def run(p, mode, **kw):
    opt = binding.default_options(reduced_system=mode, **kw)
    with binding.Engine(opt) as e:
        e.set_problem_from(p)
        plan = e.plan()
        r = e.solve()
    return r, plan


def timing(p, mode, steps=50):
    opt = binding.default_options(reduced_system=mode, max_iterations=12, min_iterations=10**6, gradient_threshold=0.0,
                                  cg_max_steps=steps, poll_every=64)
    with binding.Engine(opt) as e:
        e.set_problem_from(p); e.begin(); e.iterate(3); e.restart(); e.iterate(8)
        ms = e.last_iterate_ms() / 8
        plan = e.plan()
    return ms, plan


out = {}
for name in (sys.argv[1:] or ["small", "demo", "ladybug", "trafalgar"]):
    p = problem(name)
    row = {"N": p.N, "M": p.M, "K": p.K}
    for tag, kw in (("fixed50", {}), ("exact", {"cg_rel_tol": 1e-14, "cg_max_steps": 2000})):
        a, pa = run(p, 1, **kw)
        b, pb = run(p, 0, **kw)
        n = min(len(a.trace), len(b.trace))
        rel = np.abs(a.trace[:n, 1] - b.trace[:n, 1]) / a.trace[:n, 1]
        row[tag] = {"plan": pb["explicit_schur"], "pairs": pb["explicit_pairs"], "iters": (a.iterations, b.iterations),
                    "max_rel_cost_diff_first8": float(rel[:8].max()), "max_rel_cost_diff": float(rel.max()),
                    "final_px": (a.mean_px[1], b.mean_px[1]), "accept_equal": bool(np.array_equal(a.trace[:n, 5], b.trace[:n, 5])),
                    "cg_steps": (a.trace[:6, 6].tolist(), b.trace[:6, 6].tolist()), "ms_solve": (a.ms_solve, b.ms_solve)}
    m0, pl0 = timing(p, 1); m1, pl1 = timing(p, 0)
    row["ms_per_lm_iteration"] = {"matrix_free": m0, "explicit": m1, "launches": (pl0["launches_per_iteration"], pl1["launches_per_iteration"])}
    out[name] = row
    print(name, json.dumps(row), flush=True)
