"""CPU: the triangulation checker oracle/tri_oracle.py (numpy restatement of src/sfm_triangulation.cpp:63-202) against the
committed fixtures produced with OpenCV's own cv::undistortPoints and cv::solve(DECOMP_SVD) (tools/make_golden_tri.py) and,
live, against cv2 on fresh inputs."""
import glob
import os

import numpy as np
import pytest

from oracle import tri_oracle
from sequencesfm_b200.synth import make_two_view

from .util import GOLDEN_DIR

TRI_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "tri_*.npz")))
# the checker solves with LAPACK's SVD, the fixture with OpenCV's Jacobi SVD: both backward stable -> cond(A) * 1e-16
X_RTOL = 1e-9


def load_tri_fixture(path):
    fx = np.load(path)
    dist = fx["dist"] if len(fx["dist"]) else None
    return fx, dist


def test_fixtures_present():
    assert len(TRI_FIXTURES) >= 5


@pytest.mark.parametrize("path", TRI_FIXTURES, ids=[os.path.basename(p) for p in TRI_FIXTURES])
def test_oracle_matches_opencv_golden(path):
    fx, dist = load_tri_fixture(path)
    X, err, k1, mean = tri_oracle.triangulate(fx["pt1"], fx["pt2"], fx["K"], fx["P"], fx["P1"], dist)
    assert np.array_equal(k1, fx["pt1_undistorted"]), "undistorted pixels must be bit-identical (float32)"
    scale = np.abs(fx["X"]).max(axis=1, keepdims=True)
    assert np.abs(X - fx["X"]).max() <= X_RTOL * scale.max()
    assert np.abs(err - fx["err"]).max() < 1e-4          # the pixel error is computed from float32-rounded projections
    assert abs(mean - float(fx["mean"])) < 1e-6
    # the tracker's gating on the FIXTURE's points / errors: same cut-off bits, same mask
    status, keep, cutoff, n0, n1 = tri_oracle.gate(fx["X"], fx["err"], fx["P"], fx["P1"])
    assert status == 0 and cutoff == float(fx["gate_cutoff"]) and np.array_equal(keep, fx["gate_keep"])
    assert [n0, n1] == fx["gate_front"].tolist()


def test_gate_return_codes():
    pt1, pt2, K, P, P1, X = make_two_view(8, 50, noise_px=0.1)
    pts, err, _, _ = tri_oracle.triangulate(pt1, pt2, K, P, P1)
    assert tri_oracle.gate(pts[:9], err[:9], P, P1)[0] == -1                       # fewer than 10 points
    assert tri_oracle.gate(-pts, err, P, P1)[0] == -2                              # behind the cameras
    few = pts[:12].copy(); few[:3] *= -1.0                                         # 9 of 12 in front: exactly 75 %, not below
    st, keep, cutoff, n0, n1 = tri_oracle.gate(few, np.ones(12), P, P1)
    assert st == -3 and keep.sum() == 9 and cutoff == 2.4 and (n0, n1) == (9, 9)   # at most 10 survive


def test_oracle_pieces_against_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    A = rng.standard_normal((200, 4, 3)); B = rng.standard_normal((200, 4))
    want = np.stack([cv2.solve(A[i], B[i].reshape(4, 1), flags=cv2.DECOMP_SVD)[1][:, 0] for i in range(200)])
    assert np.abs(tri_oracle.svd_solve(A, B) - want).max() < 1e-12
    pt1, pt2, K, P, P1, _ = make_two_view(9, 500, dist=[-0.15, 0.07, 8e-4, -6e-4, 0.01])
    d = np.array([-0.15, 0.07, 8e-4, -6e-4, 0.01])
    un = cv2.undistortPoints(pt1.reshape(-1, 1, 2), K, d).reshape(-1, 2)
    want = np.stack([un[:, 0].astype(np.float64) * K[0, 0] + K[0, 2], un[:, 1].astype(np.float64) * K[1, 1] + K[1, 2]], 1).astype(np.float32)
    assert np.array_equal(tri_oracle.undistort_pixels(pt1, K, d), want)


def test_oracle_recovers_noise_free_points():
    pt1, pt2, K, P, P1, X = make_two_view(3, 400, noise_px=0.0)
    got, err, _, mean = tri_oracle.triangulate(pt1, pt2, K, P, P1)
    assert np.abs(got - X).max() < 5e-3 and mean < 1e-3   # float32 pixel storage limits it


def test_tri_header_symbols_exported():
    import re
    from sequencesfm_b200 import binding, tri
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "b200ba_tri.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(b200tri_[a-z0-9_]+)\s*\(", hdr)))
    lib = binding.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(tri.ABI_SYMBOLS) == declared


def test_triangulator_needs_a_gpu(have_cuda):
    if have_cuda:
        pytest.skip("CUDA present")
    from sequencesfm_b200 import tri
    from sequencesfm_b200.binding import B200BAError
    with pytest.raises(B200BAError) as e:
        tri.Triangulator()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
