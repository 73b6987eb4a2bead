"""Generate tests/golden/tri_*.npz with OpenCV itself (this image's cv2 wheel): the two OpenCV calls on the reference's
triangulation path, cv::undistortPoints and cv::solve(DECOMP_SVD), are executed by cv2 inside a literal transcription of
src/sfm_triangulation.cpp:63-202 (a Python loop per match).  The fixtures are committed; tests compare the numpy oracle and
the CUDA path against them.
    python tools/make_golden_tri.py
"""
import os, sys
import cv2
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200.synth import make_two_view

CASES = [("s1_300", 1, 300, {}), ("s2_wide_500", 2, 500, {"baseline": 3.0}), ("s3_dist_400", 3, 400, {"dist": [-0.12, 0.05, 1e-3, -5e-4]}),
         ("s4_narrow_200", 4, 200, {"baseline": 0.15, "noise_px": 0.2}), ("s5_dist5_257", 5, 257, {"dist": [-0.2, 0.08, 5e-4, 3e-4, -0.01]})]


def reference_loop(pt1, pt2, K, P, P1, dist):
    Kinv = np.linalg.inv(K)
    d = np.zeros(4) if dist is None else np.asarray(dist, np.float64)
    fx, cx, fy, cy = K[0, 0], K[0, 2], K[1, 1], K[1, 2]
    und = lambda p: cv2.undistortPoints(p.reshape(-1, 1, 2).astype(np.float32), K, d).reshape(-1, 2)   # 5 iterations (MAX_ITER criterion), like 2.4.9
    u1n, u2n = und(pt1), und(pt2)
    k1 = np.stack([u1n[:, 0].astype(np.float64) * fx + cx, u1n[:, 1].astype(np.float64) * fy + cy], 1).astype(np.float32)
    k2 = np.stack([u2n[:, 0].astype(np.float64) * fx + cx, u2n[:, 1].astype(np.float64) * fy + cy], 1).astype(np.float32)
    KP1 = K @ P1
    X, err = np.zeros((len(pt1), 3)), np.zeros(len(pt1))
    for i in range(len(pt1)):
        u = Kinv @ np.array([k1[i, 0], k1[i, 1], 1.0]); v = Kinv @ np.array([k2[i, 0], k2[i, 1], 1.0])
        wi = wi1 = 1.0
        for _ in range(10):
            A = np.array([(u[0] * P[2, :3] - P[0, :3]) / wi, (u[1] * P[2, :3] - P[1, :3]) / wi,
                          (v[0] * P1[2, :3] - P1[0, :3]) / wi1, (v[1] * P1[2, :3] - P1[1, :3]) / wi1])
            B = np.array([[-(u[0] * P[2, 3] - P[0, 3]) / wi], [-(u[1] * P[2, 3] - P[1, 3]) / wi],
                          [-(v[0] * P1[2, 3] - P1[0, 3]) / wi1], [-(v[1] * P1[2, 3] - P1[1, 3]) / wi1]])
            ok, x = cv2.solve(A, B, flags=cv2.DECOMP_SVD)
            x4 = np.array([x[0, 0], x[1, 0], x[2, 0], 1.0])
            p2x, p2x1 = P[2] @ x4, P1[2] @ x4
            if abs(np.float32(wi - p2x)) <= 0.0001 and abs(np.float32(wi1 - p2x1)) <= 0.0001:
                break
            wi, wi1 = p2x, p2x1
        X[i] = x4[:3]
        q = KP1 @ x4
        pr = np.array([q[0] / q[2], q[1] / q[2]], np.float32)
        dd = pr - k2[i]
        err[i] = np.sqrt(float(dd[0]) ** 2 + float(dd[1]) ** 2)
    return X, err, k1, float(err.mean())


def main():
    out = os.path.join(ROOT, "tests", "golden")
    for name, seed, n, kw in CASES:
        pt1, pt2, K, P, P1, Xt = make_two_view(seed, n, **kw)
        X, err, k1, mean = reference_loop(pt1, pt2, K, P, P1, kw.get("dist"))
        # the gating of src/sfm_tracker.cpp:718-771, literally: std::sort, index 4 n / 5, x 2.4, status of the SECOND front test
        srt = sorted(err.tolist()); cutoff = srt[4 * len(srt) // 5] * 2.4
        front0 = [(P[2, :3] @ x + P[2, 3]) > 0 for x in X]; front1 = [(P1[2, :3] @ x + P1[2, 3]) > 0 for x in X]
        keep = np.array([1 if (front1[i] and err[i] < cutoff) else 0 for i in range(len(err))], np.uint8)
        np.savez_compressed(os.path.join(out, f"tri_{name}.npz"), gate_cutoff=cutoff, gate_keep=keep, gate_front=np.array([sum(front0), sum(front1)]), seed=seed, n=n, kwargs=repr(kw), pt1=pt1, pt2=pt2, K=K, P=P, P1=P1,
                            dist=np.asarray(kw.get("dist", []), np.float64), X=X, err=err, pt1_undistorted=k1, mean=mean,
                            cv2_version=cv2.__version__)
        print(name, "mean reprojection error", mean, "max |X - truth|", np.abs(X - Xt).max())


if __name__ == "__main__":
    main()
