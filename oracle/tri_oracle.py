"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's two-view triangulation
(``TriangulatePoints`` / ``IterativeLinearLSTriangulation``, ``src/sfm_triangulation.cpp:63-202``).

The two OpenCV calls on the path are external to the reference tree (OpenCV 2.4.9, named in ``cmake/FindOpenCV.cmake``):
``cv::undistortPoints`` (modules/imgproc/src/undistort.cpp, cvUndistortPoints: 5 fixed-point iterations, float output) is
restated in ``undistort_pixels``; ``cv::solve(A, B, X, DECOMP_SVD)`` is restated as the SVD pseudo-inverse solve
``svd_solve``.  Pinned by ``tests/test_tri_oracle.py`` against ``cv2.solve`` / ``cv2.undistortPoints`` of this image's cv2
wheel (4.13; same API, same algorithm family) and against ``tests/golden/tri_*.npz`` generated with them
(``tools/make_golden_tri.py``).  Only tests/, smoke() and bench.py's CPU-baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

EPSILON = 0.0001          # src/sfm_triangulation.h:25


def undistort_pixels(pt: np.ndarray, K: np.ndarray, dist) -> np.ndarray:
    """pixels (float32 n x 2) -> undistorted pixels as the reference stores them (float32): cvUndistortPoints without
    R / P, then x * fx + cx (src/sfm_triangulation.cpp:125-155)."""
    pt = np.asarray(pt, np.float32).reshape(-1, 2)
    fx, cx, fy, cy = K[0, 0], K[0, 2], K[1, 1], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    x = (pt[:, 0].astype(np.float64) - cx) * ifx
    y = (pt[:, 1].astype(np.float64) - cy) * ify
    x0, y0 = x.copy(), y.copy()
    if dist is not None and len(dist) and np.any(np.asarray(dist) != 0):
        k = np.zeros(8)
        k[:len(dist)] = np.asarray(dist, np.float64).ravel()
        for _ in range(5):
            r2 = x * x + y * y
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x)
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y
            x = (x0 - dx) * icdist
            y = (y0 - dy) * icdist
    u = np.stack([x, y], 1).astype(np.float32)                       # CV_32FC2 destination
    return np.stack([u[:, 0].astype(np.float64) * fx + cx, u[:, 1].astype(np.float64) * fy + cy], 1).astype(np.float32)


def svd_solve(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """cv::solve(A, B, X, DECOMP_SVD) for a batch of 4x3 systems: X = V diag(1/s) U^t B, singular values below
    DBL_EPSILON * 2 * sum(s) treated as zero (OpenCV's SVBkSb threshold)."""
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    thr = (np.finfo(np.float64).eps * 2) * s.sum(axis=-1, keepdims=True)
    inv = np.where(s > thr, 1.0 / np.where(s > thr, s, 1.0), 0.0)
    return np.einsum("nji,nj->ni", Vt, inv * np.einsum("nij,ni->nj", U, B))


def iterative_ls(u: np.ndarray, P: np.ndarray, u1: np.ndarray, P1: np.ndarray) -> np.ndarray:
    """IterativeLinearLSTriangulation (:63-100), batched; u, u1: n x 3 homogeneous normalised image points."""
    n = len(u)
    wi, wi1 = np.ones(n), np.ones(n)
    X = np.zeros((n, 3))
    active = np.ones(n, bool)
    for _ in range(10):
        idx = np.nonzero(active)[0]
        if not len(idx):
            break
        a, b, w0, w1 = u[idx], u1[idx], wi[idx][:, None], wi1[idx][:, None]
        A = np.stack([(a[:, 0:1] * P[2, :3] - P[0, :3]) / w0, (a[:, 1:2] * P[2, :3] - P[1, :3]) / w0,
                      (b[:, 0:1] * P1[2, :3] - P1[0, :3]) / w1, (b[:, 1:2] * P1[2, :3] - P1[1, :3]) / w1], 1)
        B = np.stack([-(a[:, 0] * P[2, 3] - P[0, 3]) / w0[:, 0], -(a[:, 1] * P[2, 3] - P[1, 3]) / w0[:, 0],
                      -(b[:, 0] * P1[2, 3] - P1[0, 3]) / w1[:, 0], -(b[:, 1] * P1[2, 3] - P1[1, 3]) / w1[:, 0]], 1)
        x = svd_solve(A, B)
        X[idx] = x
        p2x = x @ P[2, :3] + P[2, 3]
        p2x1 = x @ P1[2, :3] + P1[2, 3]
        done = (np.abs((wi[idx] - p2x).astype(np.float32)) <= EPSILON) & (np.abs((wi1[idx] - p2x1).astype(np.float32)) <= EPSILON)
        wi[idx] = np.where(done, wi[idx], p2x)
        wi1[idx] = np.where(done, wi1[idx], p2x1)
        active[idx[done]] = False
    return X


def triangulate(pt1, pt2, K, P, P1, dist=None, Kinv=None):
    """TriangulatePoints (:95-202): returns (pts n x 3, reproj_err n, pt1_undistorted n x 2 float32, mean error)."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    Kinv = np.linalg.inv(K) if Kinv is None else np.asarray(Kinv, np.float64).reshape(3, 3)
    P, P1 = np.asarray(P, np.float64).reshape(3, 4), np.asarray(P1, np.float64).reshape(3, 4)
    k1, k2 = undistort_pixels(pt1, K, dist), undistort_pixels(pt2, K, dist)
    h = lambda k: np.concatenate([k.astype(np.float64), np.ones((len(k), 1))], 1) @ Kinv.T
    X = iterative_ls(h(k1), P, h(k2), P1)
    q = np.concatenate([X, np.ones((len(X), 1))], 1) @ (K @ P1).T
    proj = np.stack([q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]], 1).astype(np.float32)
    d = proj - k2                                                   # Point2f arithmetic
    err = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)
    return X, err, k1, float(err.mean()) if len(err) else 0.0


def gate(pts, err, P, P1):
    """The tracker's gating of a fresh triangulation (``src/sfm_tracker.cpp:718-771`` with ``TestTriangulation``,
    ``src/sfm_cameraMatrices.cpp:180-204``): returns (status, keep uint8, cutoff, n_front_P, n_front_P1)."""
    pts, err = np.asarray(pts, np.float64), np.asarray(err, np.float64)
    P, P1 = np.asarray(P, np.float64).reshape(3, 4), np.asarray(P1, np.float64).reshape(3, 4)
    n = len(pts)
    if n < 10:
        return -1, np.zeros(n, np.uint8), 0.0, 0, 0
    f0 = pts @ P[2, :3] + P[2, 3] > 0
    f1 = pts @ P1[2, :3] + P1[2, 3] > 0
    cutoff = np.sort(err)[4 * n // 5] * 2.4
    keep = (f1 & (err < cutoff)).astype(np.uint8)          # the status vector holds the SECOND TestTriangulation's result
    status = 0
    if f0.sum() / n < 0.75 or f1.sum() / n < 0.75:
        status = -2
    elif keep.sum() <= 10:
        status = -3
    return status, keep, float(cutoff), int(f0.sum()), int(f1.sum())
