"""GPU: batched two-view triangulation through the C ABI (include/b200ba_tri.h) against the OpenCV-generated fixtures and
the numpy checker.  Tolerances (include/b200ba_tri.h): undistorted pixels bit-identical (float32); points within
1e-9 x scene scale on well-conditioned pairs (QR here, Jacobi SVD in OpenCV), 1e-7 on the narrow-baseline case."""
import os

import numpy as np
import pytest

from oracle import tri_oracle
from sequencesfm_b200 import tri
from sequencesfm_b200.binding import B200BAError
from sequencesfm_b200.synth import make_two_view

from .test_tri_oracle import TRI_FIXTURES, load_tri_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    with tri.Triangulator(device=0) as t:
        yield t


@pytest.mark.parametrize("path", TRI_FIXTURES, ids=[os.path.basename(p) for p in TRI_FIXTURES])
def test_matches_opencv_golden(eng, path):
    fx, dist = load_tri_fixture(path)
    r = eng.triangulate(fx["pt1"], fx["pt2"], fx["K"], fx["P"], fx["P1"], dist)
    assert r.gpu_launches == 2 and r.n_degenerate == 0
    assert np.array_equal(r.pt1_undistorted, fx["pt1_undistorted"])
    tol = 1e-7 if "narrow" in path else 1e-9
    assert np.abs(r.pts - fx["X"]).max() <= tol * np.abs(fx["X"]).max()
    assert np.abs(r.reproj_err - fx["err"]).max() < 1e-4
    assert abs(r.mean_reproj_error - float(fx["mean"])) < 1e-6
    g = eng.gate()                                                 # on OUR device-resident points / errors
    st, keep, cutoff, n0, n1 = tri_oracle.gate(r.pts, r.reproj_err, fx["P"], fx["P1"])
    assert (g.status, g.cutoff, g.n_front_P, g.n_front_P1) == (st, cutoff, n0, n1) and np.array_equal(g.keep, keep)
    assert g.n_kept == int(keep.sum())
    assert (g.keep != fx["gate_keep"]).sum() <= 2                  # vs the fixture's own errors: only threshold ties may differ
    assert abs(g.cutoff - float(fx["gate_cutoff"])) < 1e-4


def test_gate_return_codes_and_sizes(eng):
    pt1, pt2, K, P, P1, _ = make_two_view(41, 200000)
    r = eng.triangulate(pt1, pt2, K, P, P1)
    g = eng.gate()
    st, keep, cutoff, n0, n1 = tri_oracle.gate(r.pts, r.reproj_err, P, P1)
    assert g.status == 0 and g.cutoff == cutoff and np.array_equal(g.keep, keep) and 0.75 * len(pt1) < g.n_kept < len(pt1)
    eng.triangulate(pt1[:9], pt2[:9], K, P, P1)
    assert eng.gate().status == -1                                  # "triangulated point is too few"
    Pb = P.copy(); Pb[2] *= -1.0; P1b = P1.copy(); P1b[2] *= -1.0   # cameras looking the other way: points behind them
    eng.triangulate(pt1[:500], pt2[:500], K, Pb, P1b)
    assert eng.gate().status == -2


@pytest.mark.parametrize("seed,n,kw", [(11, 1, {}), (12, 127, {}), (13, 129, {"baseline": 2.0}), (14, 100000, {}),
                                        (15, 5000, {"dist": [-0.1, 0.03, 0.0, 0.0]}), (16, 3000, {"dist": [0.0] * 5})])
def test_matches_checker(eng, seed, n, kw):
    pt1, pt2, K, P, P1, _ = make_two_view(seed, n, **kw)
    X, err, k1, mean = tri_oracle.triangulate(pt1, pt2, K, P, P1, kw.get("dist"))
    r = eng.triangulate(pt1, pt2, K, P, P1, kw.get("dist"))
    assert np.array_equal(r.pt1_undistorted, k1)
    assert np.abs(r.pts - X).max() <= 1e-9 * np.abs(X).max()
    assert np.abs(r.reproj_err - err).max() < 1e-4 and abs(r.mean_reproj_error - mean) < 1e-6
    again = eng.triangulate(pt1, pt2, K, P, P1, kw.get("dist"))
    assert np.array_equal(again.pts, r.pts) and again.mean_reproj_error == r.mean_reproj_error   # run-to-run bits


def test_edge_cases(eng):
    pt1, pt2, K, P, P1, _ = make_two_view(21, 64)
    r = eng.triangulate(pt1[:0], pt2[:0], K, P, P1)                    # empty: cv::mean of nothing is 0
    assert len(r.pts) == 0 and r.mean_reproj_error == 0.0
    with pytest.raises(B200BAError, match="n_dist"):
        eng.triangulate(pt1, pt2, K, P, P1, dist=[0.1, 0.2, 0.3])
    r = eng.triangulate(pt1, pt1, K, P, P)                              # identical views: rank-deficient systems
    assert r.n_degenerate == 64 and np.isnan(r.pts).all()
    ok = eng.triangulate(pt1, pt2, K, P, P1)
    assert ok.n_degenerate == 0 and np.isfinite(ok.pts).all()


def test_full_size_properties(eng):
    """1 M matches: points reproject onto their measurements, sit in front of both cameras, mean error as injected."""
    pt1, pt2, K, P, P1, X = make_two_view(31, 1_000_000, noise_px=0.5)
    r = eng.triangulate(pt1, pt2, K, P, P1)
    assert r.n_degenerate == 0
    z1 = r.pts[:, 2]; z2 = (r.pts @ P1[:, :3].T + P1[:, 3])[:, 2]
    assert (z1 > 0).all() and (z2 > 0).all()
    q = np.c_[r.pts, np.ones(len(X))] @ (K @ P1).T
    e2 = np.hypot(q[:, 0] / q[:, 2] - pt2[:, 0], q[:, 1] / q[:, 2] - pt2[:, 1])
    assert np.abs(e2 - r.reproj_err).max() < 1e-3 and abs(r.mean_reproj_error - e2.mean()) < 1e-5
    assert 0.2 < r.mean_reproj_error < 0.5
    assert np.median(np.abs(r.pts - X).max(axis=1)) < 0.1
