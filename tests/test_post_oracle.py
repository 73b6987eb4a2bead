"""CPU: the post-BA stage checkers.  oracle/post_oracle.c (plain-C restatement) against the committed golden fixtures made
from the reference's own functions (tools/make_golden_post.py) and, when oracle/_ref/libpost_ref.so is present, against
the reference itself on fresh seeds - bit for bit.  Also the host-only entry points of the product ABI
(b200post_draw_triples, b200post_ground_axes: O(1) scalar work the reference also does on the host)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import post as opost
from sequencesfm_b200 import post as ppost
from sequencesfm_b200.synth import make_ground_scene

from .util import GOLDEN_DIR, golden_names


def _unhex(v):
    return np.array([float.fromhex(s) for s in v], np.float64)


def _sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _same_bits(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def load_post_fixture(name):
    fx = json.load(open(os.path.join(GOLDEN_DIR, name + ".json")))
    pts, cams = make_ground_scene(fx["seed"], fx["M"], **fx["kwargs"])
    assert _sha(pts, cams) == fx["input_sha256"], "synthetic scene generator changed"
    stream = opost.rand_stream(fx["rand_stream_seed"], fx["rand_stream_len"])
    tri, used = opost.oracle_draw_triples(fx["M"], fx["n_ransac"], stream)
    assert used == fx["rand_used"], "hypothesis draw consumed a different number of rand() values than the reference"
    keep = np.unpackbits(np.frombuffer(bytes.fromhex(fx["keep_packed_hex"]), np.uint8))[:fx["M"]].astype(np.int32)
    return fx, pts, cams, tri, keep


POST_FIXTURES = golden_names("post_")


def test_fixtures_present():
    assert len(POST_FIXTURES) >= 7


@pytest.mark.parametrize("name", POST_FIXTURES)
def test_oracle_matches_reference_golden_bit_for_bit(name):
    fx, pts, cams, tri, keep = load_post_fixture(name)
    o = opost.oracle_plane_fit(pts, tri, fx["k_threshold"])
    assert np.array_equal(o.keep, keep)
    assert int(o.keep.sum()) == fx["n_kept"]
    assert _same_bits(o.plane, _unhex(fx["plane"]))
    assert _same_bits(o.plane_se3, _unhex(fx["plane_se3"]))
    if np.isnan(o.plane).any():
        return          # the all-ties scene: sigma^2 = 0, no inliers, the reference itself produces NaN
    axes = opost.oracle_ground_axes(cams[0], o.plane, o.plane_se3)
    assert _same_bits(axes, _unhex(fx["axes_se3"]))
    p2, c2 = opost.oracle_transform(axes, pts[keep != 0], cams)
    assert _sha(p2) == fx["points_out_sha256"] and _sha(c2) == fx["cameras_out_sha256"]
    assert _same_bits(c2, _unhex(fx["cameras_out"]))


@pytest.mark.skipif(not opost.have_ref(), reason="oracle/_ref/libpost_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed,M,kw", [(11, 64, {}), (12, 999, {"outlier_frac": 0.4}), (13, 30000, {"noise": 0.05}),
                                        (14, 2048, {"tilt": 1.5}), (15, 100, {"outlier_frac": 0.0})])
def test_oracle_matches_live_reference(seed, M, kw):
    pts, cams = make_ground_scene(seed, M, **kw)
    stream = opost.rand_stream(seed, 600)
    r, used = opost.ref_plane_fit(pts, stream, 100, 2.0)
    tri, used2 = opost.oracle_draw_triples(M, 100, stream)
    assert used == used2
    o = opost.oracle_plane_fit(pts, tri, 2.0)
    assert np.array_equal(o.keep, r.keep) and _same_bits(o.plane, r.plane) and _same_bits(o.plane_se3, r.plane_se3)
    a_r = opost.ref_ground_axes(cams[0], r.plane, r.plane_se3)
    assert _same_bits(a_r, opost.oracle_ground_axes(cams[0], r.plane, r.plane_se3))
    pr, cr = opost.ref_transform(a_r, pts, cams)
    po, co = opost.oracle_transform(a_r, pts, cams)
    assert _same_bits(pr, po) and _same_bits(cr, co)


@pytest.mark.skipif(not opost.have_ref(), reason="oracle/_ref/libpost_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("n_ransac,k_thr", [(1, 2.0), (7, 3.0), (250, 0.5)])
def test_oracle_matches_live_reference_other_settings(n_ransac, k_thr):
    """nRansacs and kThreshold away from their defaults (src/sfm_utils.h:90-91)."""
    pts, cams = make_ground_scene(40 + n_ransac, 3001, outlier_frac=0.25)
    stream = opost.rand_stream(n_ransac, 4 * n_ransac + 64)
    r, used = opost.ref_plane_fit(pts, stream, n_ransac, k_thr)
    tri, used2 = opost.oracle_draw_triples(len(pts), n_ransac, stream)
    o = opost.oracle_plane_fit(pts, tri, k_thr)
    assert used == used2 and np.array_equal(o.keep, r.keep) and _same_bits(o.plane, r.plane) and _same_bits(o.plane_se3, r.plane_se3)


def test_oracle_rejects_what_the_reference_rejects():
    pts, _ = make_ground_scene(1, 9)
    with pytest.raises(RuntimeError, match="rc=-1"):      # nPoints < 10 -> -1, src/sfm_utils.cpp:826-829
        opost.oracle_plane_fit(pts, np.zeros((4, 3), np.int32))


# ---- host-only entry points of the product ABI (no GPU needed) --------------------------------------------------------
def test_post_header_symbols_exported():
    from sequencesfm_b200 import binding
    lib = binding.load()
    for sym in ppost.ABI_SYMBOLS:
        assert hasattr(lib, sym), f"libb200ba.so does not export {sym}"


@pytest.mark.parametrize("name", [n for n in POST_FIXTURES if "flat" not in n])
def test_product_ground_axes_bit_exact(name):
    fx, pts, cams, tri, keep = load_post_fixture(name)
    axes = ppost.ground_axes(cams[0], _unhex(fx["plane"]), _unhex(fx["plane_se3"]))
    assert _same_bits(axes, _unhex(fx["axes_se3"]))


def test_product_draw_triples_follows_rand():
    """b200post_draw_triples consumes the C library's rand() exactly like src/sfm_utils.cpp:837-845."""
    libc = C.CDLL(None)
    libc.srand(1234)
    stream = np.array([libc.rand() for _ in range(700)], np.int32)
    want, used = opost.oracle_draw_triples(57, 200, stream)
    libc.srand(1234)
    got = ppost.draw_triples(57, 200)
    assert np.array_equal(got, want)
    assert libc.rand() == stream[used]                  # and leaves the stream where the reference would
    assert (got[:, 0] != got[:, 1]).all() and (got[:, 1] != got[:, 2]).all() and (got[:, 0] != got[:, 2]).all()


def test_post_stage_needs_a_gpu(have_cuda):
    if have_cuda:
        pytest.skip("CUDA present")
    from sequencesfm_b200.binding import B200BAError
    with pytest.raises(B200BAError) as e:
        ppost.PostStage()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
