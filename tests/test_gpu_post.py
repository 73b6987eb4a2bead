"""GPU: the post-BA ground-plane stage through the C ABI (include/b200ba_post.h) against the CPU checker
(oracle/post_oracle.c, itself bit-exact against the reference's own functions) and the committed golden fixtures.

Bar (include/b200ba_post.h): winning hypothesis, sigma^2, cpFilter and the re-axis are bit-exact; plane / plane_se3 agree to
the single precision the reference computes them in (tolerance written below)."""
import numpy as np
import pytest

from oracle import post as opost
from sequencesfm_b200 import post as ppost
from sequencesfm_b200.binding import B200BAError
from sequencesfm_b200.synth import make_ground_scene

from .test_post_oracle import POST_FIXTURES, _same_bits, _unhex, load_post_fixture

pytestmark = pytest.mark.gpu

# The reference runs its 3x3 SVD in float on a float copy of the scatter matrix, so its plane carries ~6e-8 relative
# noise; our inlier sums are tree-ordered (1e-16 relative), which can move the float copy by one ulp.
PLANE_ATOL = 2e-6


@pytest.fixture(scope="module")
def stage():
    with ppost.PostStage(device=0) as s:
        yield s


def _triples(seed, M, H):
    tri, _ = opost.oracle_draw_triples(M, H, opost.rand_stream(seed, 4 * H + 64))
    return tri


def _check_fit(g, o, pts):
    assert g.best_hypothesis == o.best_hypothesis
    assert g.sigma2 == o.sigma2, "sigma^2 must be bit-exact (exact median, reference operation order)"
    assert np.array_equal(g.keep, o.keep), "cpFilter must be bit-exact"
    assert g.n_kept == int(o.keep.sum()) and g.n_inliers == o.n_inliers
    assert abs(g.score - o.hyp_score[o.best_hypothesis]) <= 1e-12 * abs(g.score)
    np.testing.assert_allclose(g.plane, o.plane, rtol=0, atol=PLANE_ATOL * max(1.0, abs(o.plane[3])))
    scale = max(1.0, np.abs(o.plane_se3[:, 3]).max())
    np.testing.assert_allclose(g.plane_se3, o.plane_se3, rtol=0, atol=PLANE_ATOL * scale)


@pytest.mark.parametrize("name", [n for n in POST_FIXTURES if "flat" not in n])
def test_plane_fit_matches_reference_golden(stage, name):
    fx, pts, cams, tri, keep = load_post_fixture(name)
    g = stage.plane_fit(pts, tri, fx["k_threshold"])
    assert np.array_equal(g.keep, keep), "cpFilter differs from the reference's"
    assert g.n_kept == fx["n_kept"] and g.gpu_launches > 0
    ref_plane, ref_se3 = _unhex(fx["plane"]), _unhex(fx["plane_se3"]).reshape(3, 4)
    np.testing.assert_allclose(g.plane, ref_plane, rtol=0, atol=PLANE_ATOL * max(1.0, abs(ref_plane[3])))
    np.testing.assert_allclose(g.plane_se3, ref_se3, rtol=0, atol=PLANE_ATOL * max(1.0, np.abs(ref_se3[:, 3]).max()))
    # re-axis with the reference's axes: bit-exact
    axes = _unhex(fx["axes_se3"]).reshape(3, 4)
    p2, c2 = stage.transform(axes, pts[keep != 0], cams)
    assert _same_bits(c2, _unhex(fx["cameras_out"]))
    assert _same_bits(p2[:4], _unhex(fx["points_out_head"]))
    po, co = opost.oracle_transform(axes, pts[keep != 0], cams)
    assert _same_bits(p2, po) and _same_bits(c2, co)


@pytest.mark.parametrize("seed,M,H,kw", [
    (21, 10, 100, {"outlier_frac": 0.0}),            # the minimum the reference accepts
    (22, 255, 1, {}), (23, 257, 13, {}), (24, 4097, 100, {"outlier_frac": 0.35}),
    (25, 70001, 100, {}), (26, 200000, 100, {"noise": 0.02}), (27, 33333, 201, {"tilt": 2.0}),
])
def test_plane_fit_matches_oracle(stage, seed, M, H, kw):
    pts, _ = make_ground_scene(seed, M, **kw)
    tri = _triples(seed, M, H)
    o = opost.oracle_plane_fit(pts, tri, 2.0)
    _check_fit(stage.plane_fit(pts, tri, 2.0), o, pts)


@pytest.mark.parametrize("cap", [-1, 2, 64])
def test_every_selection_path_gives_the_same_median(cap, stage):
    """cap = -1: six radix passes, no candidate buffer; cap = 2: the buffer overflows and the call falls back to them;
    cap = 64: some hypotheses fill most of the buffer."""
    pts, _ = make_ground_scene(31, 50001)
    tri = _triples(31, len(pts), 40)
    base = stage.plane_fit(pts, tri)
    assert base.select_fallback == 0
    with ppost.PostStage(device=0, select_cap=cap) as s:
        g = s.plane_fit(pts, tri)
    assert g.select_fallback == {-1: 0, 2: 1}.get(cap, g.select_fallback)
    assert (g.sigma2, g.best_hypothesis, g.score) == (base.sigma2, base.best_hypothesis, base.score)
    assert np.array_equal(g.keep, base.keep) and _same_bits(g.plane_se3, base.plane_se3)
    o = opost.oracle_plane_fit(pts, tri, 2.0)
    assert g.sigma2 == o.sigma2 and np.array_equal(g.keep, o.keep)


def test_ties_duplicates_and_degenerate_hypotheses(stage):
    """Many identical distances around the median (points exactly on a plane), duplicated points (degenerate triples are
    skipped like src/sfm_utils.cpp:853) - the exact order statistic must still come out."""
    pts, _ = make_ground_scene(41, 6000, tilt=0.0, on_plane_frac=0.45, outlier_frac=0.2)
    pts[100:140] = pts[100]                                   # 40 copies of one point
    tri = _triples(41, len(pts), 60)
    tri[5] = [100, 101, 102]                                  # zero normal
    tri[17] = [110, 120, 130]
    o = opost.oracle_plane_fit(pts, tri, 2.0)
    g = stage.plane_fit(pts, tri, 2.0)
    assert g.n_degenerate == 2
    _check_fit(g, o, pts)


def test_error_behaviour(stage):
    pts, _ = make_ground_scene(51, 500)
    tri = _triples(51, 500, 20)
    with pytest.raises(B200BAError) as e:                     # reference: "too few points" -> -1
        stage.plane_fit(pts[:9], tri % 9)
    assert e.value.code == -1
    with pytest.raises(B200BAError):                          # index out of range
        stage.plane_fit(pts, tri + 490)
    with pytest.raises(B200BAError, match="degenerate"):
        stage.plane_fit(pts, np.zeros((5, 3), np.int32))
    with pytest.raises(B200BAError, match="4096 hypotheses"):
        stage.plane_fit(pts, np.tile(tri, (300, 1)))
    bad = pts.copy(); bad[77, 1] = np.nan
    with pytest.raises(B200BAError, match="non-finite"):
        stage.plane_fit(bad, tri)
    flat, _ = make_ground_scene(5, 3000, tilt=0.0, on_plane_frac=0.7)   # median distance 0: the reference yields NaN
    with pytest.raises(B200BAError, match="no inliers"):
        stage.plane_fit(flat, _triples(5, 3000, 100))
    g = stage.plane_fit(pts, tri)                             # the handle survives all of that
    assert g.n_kept > 400


def test_ground_stage_equals_the_reference_sequence(stage):
    """fit -> drop -> axes from the first camera -> re-axis (src/sfm_tracker.cpp:884-904) in one call."""
    pts, cams = make_ground_scene(61, 40000)
    tri = _triples(61, len(pts), 100)
    o = opost.oracle_plane_fit(pts, tri, 2.0)
    fit, p2, c2, axes = stage.ground_stage(pts, cams, tri, 2.0)
    assert np.array_equal(fit.keep, o.keep) and len(p2) == int(o.keep.sum())
    # same arithmetic as the checker when it starts from OUR plane (bit-exact) ...
    axes_o = opost.oracle_ground_axes(cams[0], fit.plane, fit.plane_se3)
    assert _same_bits(axes, axes_o)
    po, co = opost.oracle_transform(axes_o, pts[o.keep != 0], cams)
    assert _same_bits(p2, po) and _same_bits(c2, co)
    # ... and close to the all-reference chain (the plane carries the reference's float noise)
    axes_r = opost.oracle_ground_axes(cams[0], o.plane, o.plane_se3)
    pr, cr = opost.oracle_transform(axes_r, pts[o.keep != 0], cams)
    np.testing.assert_allclose(p2, pr, rtol=0, atol=1e-4)
    np.testing.assert_allclose(c2, cr, rtol=0, atol=1e-4)
    # properties: the kept points now lie about z = 0 and camera 0 sits above the origin looking along +x of the ground
    assert abs(np.median(p2[:, 2])) < 0.1
    R0, t0 = c2[0].reshape(3, 4)[:, :3], c2[0].reshape(3, 4)[:, 3]
    centre0 = -R0.T @ t0
    # (1e-4: the plane normal is a single-precision vector, |n| - 1 ~ 1e-7, times a 50 m camera height)
    assert abs(centre0[0]) < 1e-4 and abs(centre0[1]) < 1e-4 and abs(R0[0, 1]) < 1e-6 and centre0[2] > 10


def test_full_size_properties(stage):
    """Venice-sized map (993 923 points, 100 hypotheses): size-independent properties instead of the CPU checker."""
    M = 993923
    pts, cams = make_ground_scene(71, M)
    tri = _triples(71, M, 100)
    g = stage.plane_fit(pts, tri)
    assert g.select_fallback == 0
    b = g.best_hypothesis
    A, B, Cc = pts[tri[b, 0]], pts[tri[b, 1]], pts[tri[b, 2]]
    mean = ((A + B) + Cc) / 3.0
    n = np.cross(Cc - A, B - A); n /= np.linalg.norm(n)
    err = np.abs((pts - mean) @ n)
    med = np.partition(err, M // 2)[M // 2]
    s = 1.4826 * (1 + 5.0 / (2 * M - 6)) * np.sqrt(med) * 4.6851
    assert abs(g.sigma2 - s * s / 20.0) <= 1e-12 * g.sigma2   # numpy's dot product associates differently: not bit-exact
    keep_np = err < 2.0 * g.sigma2
    assert (keep_np != (g.keep != 0)).sum() <= 2              # only points within an ulp of the threshold may differ
    assert g.n_kept == int(g.keep.sum())
    assert abs(g.score - np.minimum(err, g.sigma2).sum()) <= 1e-9 * g.score
    inl = pts[err < g.sigma2]
    np.testing.assert_allclose(g.inlier_mean, inl.mean(0), rtol=0, atol=1e-9)
    np.testing.assert_allclose(g.inlier_scatter, (inl - inl.mean(0)).T @ (inl - inl.mean(0)), rtol=1e-9, atol=1e-6)
    # idempotence of the mask: fitting the kept points again with the same hypothesis keeps a superset-free subset
    again = stage.plane_fit(pts, tri)
    assert np.array_equal(again.keep, g.keep) and again.sigma2 == g.sigma2, "run-to-run bit reproducibility"


@pytest.mark.parametrize("mode", ["calls", "fused"])
def test_cpp_adapter_runs_the_tracker_sequence(tmp_path, mode):
    """adapter/sfm_post_b200.cpp (the reference's five function names, built against the shim headers) after srand(seed):
    same hypotheses as the reference's rand() draw, same mask, same re-axed map as the checker chain."""
    import ctypes as C
    import os
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pts, cams = make_ground_scene(81, 12345)
    exe = str(tmp_path / "test_post_adapter")
    subprocess.check_call(["g++", "-std=c++11", "-O1", f"-I{ROOT}/adapter/shim", f"-I{ROOT}/include",
                           f"{ROOT}/adapter/test_post_adapter.cpp", f"{ROOT}/adapter/sfm_post_b200.cpp",
                           f"-L{ROOT}/sequencesfm_b200", "-lb200ba", f"-Wl,-rpath,{ROOT}/sequencesfm_b200", "-o", exe])
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    with open(fin, "w") as f:
        f.write(f"{len(cams)} {len(pts)}\n")
        for row in cams: f.write(" ".join(repr(float(v)) for v in row) + "\n")
        for row in pts: f.write(" ".join(repr(float(v)) for v in row) + "\n")
    subprocess.check_call([exe, mode, "4321", str(fin), str(fout)])
    lines = open(fout).read().split("\n")
    ret, m_out = (int(v) for v in lines[0].split())
    assert ret == 0
    plane = np.array(lines[1].split(), np.float64)
    axes = np.array(lines[2].split(), np.float64).reshape(3, 4)
    cam_out = np.array([l.split() for l in lines[3:3 + len(cams)]], np.float64)
    rest = np.array([l.split() for l in lines[3 + len(cams):3 + len(cams) + m_out]], np.float64)
    ids, p_out = rest[:, 0].astype(int), rest[:, 1:]
    # the checker chain on glibc's rand() stream for the same seed
    libc = C.CDLL(None)
    libc.srand(4321)
    stream = np.array([libc.rand() for _ in range(464)], np.int32)
    tri, _ = opost.oracle_draw_triples(len(pts), 100, stream)
    o = opost.oracle_plane_fit(pts, tri, 2.0)
    assert m_out == int(o.keep.sum()) and np.array_equal(ids, np.nonzero(o.keep)[0]), "kept points / order / carried fields"
    np.testing.assert_allclose(plane, o.plane, rtol=0, atol=PLANE_ATOL * max(1.0, abs(o.plane[3])))
    po, co = opost.oracle_transform(axes, pts[o.keep != 0], cams)      # from the adapter's own axes: bit-exact
    assert _same_bits(p_out, po) and _same_bits(cam_out, co)
    np.testing.assert_allclose(axes, opost.oracle_ground_axes(cams[0], o.plane, o.plane_se3), rtol=0, atol=1e-4)


@pytest.mark.parametrize("discard", [0, 1])
def test_stage_chained_onto_a_bundle_adjustment_on_the_device(stage, discard):
    """b200post_ground_stage_ba: the adjusted points go from the BA engine to the plane fit without leaving the device
    (b200ba_get_result_device) - same bits as downloading them with b200ba_get_result and running the stage on the host
    arrays.  discard = 1: gradient-converged result discarded like src/sfm_ba_ssba.cpp:221 -> the inputs flow through."""
    from sequencesfm_b200 import binding
    from sequencesfm_b200.synth import make_problem
    p = make_problem("ladybug", 3)
    kw = dict(device=0, max_iterations=6) if not discard else dict(device=0, max_iterations=50, gradient_threshold=1e30, discard_converged=1)
    with binding.Engine(**kw) as eng:
        eng.set_problem_from(p)
        res = eng.solve()
        assert (res.ret == -1) == bool(discard)
        tri = _triples(91, p.M, 100)
        fit_d, pts_d, cams_d, axes_d = stage.ground_stage_ba(eng, tri)
        assert eng.N == p.N and eng.M == p.M
    if discard:
        assert np.array_equal(res.pts, p.pts)
    fit_h, pts_h, cams_h, axes_h = stage.ground_stage(res.pts, res.cam_RT, tri)
    assert np.array_equal(fit_d.keep, fit_h.keep) and fit_d.sigma2 == fit_h.sigma2 and fit_d.best_hypothesis == fit_h.best_hypothesis
    assert _same_bits(fit_d.plane_se3, fit_h.plane_se3) and _same_bits(axes_d, axes_h)
    assert _same_bits(pts_d, pts_h) and _same_bits(cams_d, cams_h)
    o = opost.oracle_plane_fit(res.pts, tri, 2.0)                     # and the checker on the downloaded BA result
    assert np.array_equal(fit_d.keep, o.keep) and fit_d.sigma2 == o.sigma2
    po, co = opost.oracle_transform(axes_d, res.pts[o.keep != 0], res.cam_RT)
    assert _same_bits(pts_d, po) and _same_bits(cams_d, co)
