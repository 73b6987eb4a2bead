"""Generate tests/golden/*.json from the UNMODIFIED reference engines (oracle/_ref, built by oracle/Makefile
from /root/reference).  Run in the build container only (it needs oracle/_ref); the fixtures are committed.

    python tools/make_golden.py            # all cases
Each fixture records the generator arguments (the problem is re-created from the seed by
sequencesfm_b200.synth.make_problem, itself pinned by an input checksum) and the reference's own numbers:
SSBA: per-iteration |residual|^2, lambda, trial cost, rho, accept flag, status, iteration count, mean
reprojection error before/after, samples of the adjusted cameras/points.  PBA: initial/final MSE, LM/CG counts.
"""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200.synth import make_problem
from oracle import refs

CASES = [
    ("tiny", 1, {}), ("tiny", 2, {"outlier_frac": 0.05}), ("tiny", 3, {"fy_ratio": 1.02, "pp": (12.0, -7.0)}),
    ("small", 1, {}), ("small", 2, {}), ("small", 3, {"outlier_frac": 0.02}),
    ("ladybug", 1, {}), ("ladybug", 2, {}), ("ladybug", 3, {"rot_sigma": 1e-3}),
    ("trafalgar", 1, {}),
]

def checksum(p):
    h = hashlib.sha256()
    for a in (p.K5, p.cam_RT, p.pts, p.obs_xy, p.obs_cam, p.obs_pt):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()

def add_pba_poses(out_dir):
    """python tools/make_golden.py --pba-poses : add PBA's adjusted cameras to the existing fixtures (no SSBA re-run)."""
    for shape in ("ladybug", "trafalgar"):
        path = os.path.join(out_dir, f"ssba_{shape}_s1.json")
        fx = json.load(open(path))
        p = make_problem(shape, 1, **fx["kwargs"])
        assert checksum(p) == fx["input_sha256"]
        q = refs.pba_adjust(p, double=True, args=("-schur", "-v", "0"))
        assert abs(q.final_mse - fx["pba_double_schur"]["final_mse"]) < 1e-6 * q.final_mse
        fx["pba_double_schur"]["cam_RT_all"] = q.cam_RT.tolist()
        json.dump(fx, open(path, "w"), indent=1)
        print(shape, "PBA cameras stored", q.cam_RT.shape, flush=True)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if "--pba-poses" in sys.argv:
        return add_pba_poses(out_dir)
    only = sys.argv[1:] 
    for shape, seed, kw in CASES:
        name = f"ssba_{shape}_s{seed}"
        if only and name not in only and shape not in only:
            continue
        p = make_problem(shape, seed, **kw)
        t = time.time()
        r = refs.ssba_adjust(p)
        fx = {"engine": "SSBA-3.0 via oracle/ssba_ref_harness.cpp (unmodified /root/reference sources)",
              "shape": shape, "seed": seed, "kwargs": kw, "N": p.N, "M": p.M, "K": p.K, "input_sha256": checksum(p),
              "ret": r.ret, "status": r.status, "iterations": r.iterations,
              "mean_px": list(r.mean_px), "seconds_minimize": r.seconds,
              "trace_cols": ["iter", "err", "lambda", "new_err", "rho", "accepted"],
              "trace": [[None if np.isnan(v) else float(v) for v in row] for row in r.trace],
              "cam_RT_head": r.cam_RT[:3].tolist(), "pts_head": r.pts[:5].tolist(),
              "cam_RT_sum": float(np.sum(r.cam_RT)), "pts_sum": float(np.sum(r.pts))}
        if shape in ("ladybug", "trafalgar") and seed == 1:
            for dbl in (True, False):
                q = refs.pba_adjust(p, double=dbl, args=("-schur", "-v", "0"))
                fx["pba_double_schur" if dbl else "pba_float_schur"] = {
                    "initial_mse": q.initial_mse, "final_mse": q.final_mse, "lm_iterations": q.lm_iterations,
                    "cg_iterations": q.cg_iterations, "return_code": q.return_code, "seconds": q.seconds,
                    "cam_RT_head": q.cam_RT[:3].tolist(), "pts_head": q.pts[:5].tolist()}
                if dbl:      # every adjusted camera: the pose-agreement test aligns them with a 7-dof similarity
                    fx["pba_double_schur"]["cam_RT_all"] = q.cam_RT.tolist()
        json.dump(fx, open(os.path.join(out_dir, name + ".json"), "w"), indent=1)
        print(name, "status", r.status, "its", r.iterations, "px", r.mean_px, f"{time.time()-t:.1f}s", flush=True)

if __name__ == "__main__":
    main()
