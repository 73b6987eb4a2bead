"""Timing of the batched two-view triangulation (b200tri_triangulate) beside the CPU: the literal per-match loop of
src/sfm_triangulation.cpp with OpenCV's own solve / undistortPoints (tools/make_golden_tri.reference_loop, on a bounded
sample) and the vectorised numpy checker.   python tools/tri_bench.py [n]  -> one JSON object"""
import importlib.util, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200 import tri
from sequencesfm_b200.synth import make_two_view
# oracle/ is imported only inside the CPU leg below


def run(n=1_000_000, cpu=True, device=0):
    pt1, pt2, K, P, P1, X = make_two_view(5, n)
    with tri.Triangulator(device=device) as eng:
        eng.triangulate(pt1, pt2, K, P, P1)
        best = None
        for _ in range(3):
            t = time.perf_counter(); r = eng.triangulate(pt1, pt2, K, P, P1); wall = (time.perf_counter() - t) * 1e3
            if best is None or wall < best["wall_ms"]:
                best = {"wall_ms": wall, "ms_device": r.ms_device, "ms_kernels": r.ms_kernels}
    out = {"n_matches": n, **best, "matches_per_s_kernels": n / (best["ms_kernels"] * 1e-3), "mean_reproj_error": r.mean_reproj_error,
           "bytes_per_match": {"in": 16, "out": 40}, "kernel_gbs": n * 56 / (best["ms_kernels"] * 1e-3) / 1e9}
    if not cpu:
        return out
    from oracle import tri_oracle
    m = min(n, 200000)
    t = time.perf_counter(); Xo, eo, _, mo = tri_oracle.triangulate(pt1[:m], pt2[:m], K, P, P1); dt = time.perf_counter() - t
    out["cpu_numpy_checker"] = {"sample": m, "ms": dt * 1e3, "matches_per_s": m / dt, "cores": "numpy / LAPACK batched SVD",
                                "max_abs_diff_points": float(np.abs(Xo - r.pts[:m]).max())}
    try:
        spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tools", "make_golden_tri.py"))
        mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
        m = min(n, 20000)
        t = time.perf_counter(); mg.reference_loop(pt1[:m], pt2[:m], K, P, P1, None); dt = time.perf_counter() - t
        out["cpu_opencv_loop"] = {"sample": m, "ms": dt * 1e3, "matches_per_s": m / dt, "cores": 1,
                                  "what": "per-match loop of src/sfm_triangulation.cpp:158-195 with cv2.solve(DECOMP_SVD) (Python call overhead included)"}
    except Exception as ex:
        out["cpu_opencv_loop"] = {"error": str(ex)}
    return out


if __name__ == "__main__":
    print(json.dumps({"triangulation": run(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)}, indent=1))
