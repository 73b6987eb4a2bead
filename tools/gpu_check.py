"""First-light GPU check: stage parity + full-solve parity vs the C oracle, then kernel timings at Venice shape."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem, SHAPES
from sequencesfm_b200 import binding
from oracle import oracle

out = {}
def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))

for shape, seed in [("tiny", 1), ("small", 2), ("ladybug", 1)]:
    p = make_problem(shape, seed, outlier_frac=0.02)
    eng = binding.Engine()
    eng.set_problem_from(p)
    eng.begin()
    g = eng.debug_linearize()
    o = oracle.linearize(p)
    x = np.random.default_rng(0).normal(size=(p.N, 6))
    lam = o["lambda0"]
    Sx = eng.debug_schur_matvec(lam, x)
    oS = oracle.linearize(p, x=x.reshape(-1), lam=lam)["Sx"]
    stage = {k: rel(g[k], o[k]) for k in ["w", "gc", "gp", "U", "V"]}
    stage["cost"] = abs(g["cost"] - o["cost"]) / o["cost"]; stage["lambda0"] = abs(g["lambda0"] - lam) / lam
    stage["Sx"] = rel(Sx, oS)
    t = time.time(); r = eng.solve(); tg = time.time() - t
    t = time.time(); c = oracle.solve(p); tc = time.time() - t
    n = min(len(r.trace), len(c.trace))
    same = 0
    while same < n and r.trace[same, 5] == c.trace[same, 5]: same += 1
    errrel = np.abs(r.trace[:same, 1] - c.trace[:same, 1]) / c.trace[:same, 1]
    out[shape] = {"stage": stage, "gpu": [r.status, r.iterations, r.mean_px, r.ms_solve, tg, r.gpu_launches],
                  "cpu": [c.status, c.iterations, c.mean_px, tc], "same_accept_prefix": same,
                  "err_rel_max": float(errrel.max()) if same else None,
                  "final_cost": [r.final_cost, float(c.trace[-1, 1])],
                  "cam_diff": float(np.abs(r.cam_RT - c.cam_RT).max()), "pts_diff": float(np.abs(r.pts - c.pts).max())}
    print(shape, json.dumps(out[shape]), flush=True)
    eng.close()

for shape in ["trafalgar", "venice"]:
    t = time.time(); p = make_problem(shape, 1); tgen = time.time() - t
    eng = binding.Engine(max_iterations=10)
    t = time.time(); eng.set_problem_from(p); tset = time.time() - t
    eng.begin()
    eng.debug_linearize()
    names = eng.kernel_names()
    times = {n: eng.time_kernel(i, 10) for i, n in enumerate(names)}
    t = time.time(); r = eng.solve(); tg = time.time() - t
    out[shape] = {"gen_s": tgen, "set_s": tset, "kernel_ms": times, "solve": [r.status, r.iterations, r.mean_px, r.ms_solve, tg, r.gpu_launches],
                  "ms_per_lm_iter": r.ms_solve / max(1, r.iterations), "trace_err": r.trace[:, 1].tolist(), "acc": r.trace[:, 5].tolist()}
    print(shape, json.dumps(out[shape]), flush=True)
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
