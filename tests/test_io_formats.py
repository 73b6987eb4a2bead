"""CPU: on-disk model formats (include/b200ba_io.h) against the reference's own readers / writers
(Thirdparty/pba/pba/pba_util.h through oracle/io_ref_harness.cpp) in both directions, committed golden files written by the
reference, lossless round trips, and the error behaviour."""
import os

import numpy as np
import pytest

from oracle import refs
from sequencesfm_b200 import io as bio
from sequencesfm_b200.binding import B200BAError
from sequencesfm_b200.synth import make_problem

from .util import GOLDEN_DIR

FLOAT = 4e-7          # the reference stores float: relative agreement of its values with the file's doubles
EXT = {0: "txt", 1: "nvm", 2: "out"}


def _model(seed=1, shape="tiny", rgb=True):
    p = make_problem(shape, seed)
    m = bio.from_problem(p)
    rng = np.random.default_rng(seed)
    m.focal = rng.uniform(500, 700, m.N)                      # BAL has one focal per camera
    m.radial = rng.uniform(-1e-2, 1e-2, m.N)
    if rgb:
        m.pt_rgb = rng.integers(0, 256, (m.M, 3)).astype(np.int32)
    m.names = [f"img{i:03d}.jpg" for i in range(m.N)]
    return m


def _close(a, b, tol=FLOAT):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))))


def _same_model(a, b, tol, rgb=True, dtype=None):
    assert np.array_equal(a.obs_cam, b["obs_cam"] if isinstance(b, dict) else b.obs_cam)
    g = (lambda k: b[k]) if isinstance(b, dict) else (lambda k: getattr(b, k))
    assert np.array_equal(a.obs_pt, g("obs_pt"))
    for k in ("cam_RT", "focal", "pts", "obs_xy"):
        assert _close(getattr(a, k), g(k), tol), k
    if rgb:
        assert np.array_equal(a.pt_rgb, g("pt_rgb"))


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_round_trip_is_lossless(tmp_path, kind):
    m = _model(kind + 1)
    if kind == 1:
        m.distortion_type[:] = -1
    path = tmp_path / f"model.{EXT[kind]}"
    bio.save(path, m)
    r = bio.load(path)
    assert np.array_equal(r.obs_cam, m.obs_cam) and np.array_equal(r.obs_pt, m.obs_pt)
    assert np.array_equal(r.pts, m.pts) and np.array_equal(r.focal, m.focal) and np.array_equal(r.obs_xy, m.obs_xy)
    if kind == 0:     # BAL stores the rotation as a Rodrigues vector: exp(log(R)) is exact to rounding, not bit for bit
        assert np.abs(r.cam_RT - m.cam_RT).max() < 1e-13
        assert (r.pt_rgb == 0).all() and r.names == []
    else:
        assert np.array_equal(r.cam_RT, m.cam_RT) and np.array_equal(r.pt_rgb, m.pt_rgb)
    if kind == 1:
        assert r.names == m.names and (r.distortion_type == -1).all()
        assert np.allclose(r.radial, m.radial, rtol=1e-15)       # stored as d = radial f^2, read back as d / f^2
    else:
        assert np.array_equal(r.radial, m.radial) and (r.distortion_type == 1).all()


@pytest.mark.skipif(not refs.have_io(), reason="oracle/_ref/libio_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_reads_what_the_reference_writes_and_vice_versa(tmp_path, kind):
    m = _model(10 + kind, "small")
    if kind == 1:
        m.distortion_type[:] = -1
    # the reference writes (setprecision(12) of its float copy), both read: same model to float precision
    p_ref = tmp_path / f"ref.{EXT[kind]}"
    refs.io_save(p_ref, kind, m)
    ours, theirs = bio.load(p_ref), refs.io_load(p_ref)
    _same_model(ours, theirs, FLOAT, rgb=kind != 0)
    assert _close(ours.radial, theirs["radial"], 1e-6) and np.array_equal(ours.distortion_type, theirs["distortion_type"])
    # (BAL: the reference's writer takes the Rodrigues vector of its float matrix next to the pi singularity - the flipped
    #  rotation of a camera looking along +z is a half turn - and loses digits there: measured up to 1.2e-3)
    if kind == 2:     # SaveBundlerOut bounds its per-point scan by point_data.size() (pba_util.h:293): it drops observations
        assert ours.K < m.K and _close(ours.cam_RT, m.cam_RT) and _close(ours.pts, m.pts) and _close(ours.focal, m.focal)
    else:
        _same_model(ours, m, 1e-2 if kind == 0 else FLOAT, rgb=kind != 0)
    # we write, the reference reads: it sees the model (in float)
    p_our = tmp_path / f"our.{EXT[kind]}"
    bio.save(p_our, m)
    back = refs.io_load(p_our)
    _same_model(m, back, FLOAT, rgb=kind != 0)
    assert _close(m.radial, back["radial"], 1e-6)


@pytest.mark.parametrize("name", ["io_ref_tiny.txt", "io_ref_tiny.nvm", "io_ref_tiny.out"])
def test_golden_files_written_by_the_reference(name):
    """tests/golden/io_ref_tiny.*: written by the reference's SaveBundlerModel / SaveNVM / SaveBundlerOut
    (tools/make_golden_io.py); io_ref_tiny.npz holds what the reference's own loaders read back from them."""
    want = np.load(os.path.join(GOLDEN_DIR, "io_ref_tiny.npz"))
    key = name.rsplit(".", 1)[1]
    got = bio.load(os.path.join(GOLDEN_DIR, name))
    assert np.array_equal(got.obs_cam, want[f"{key}_obs_cam"]) and np.array_equal(got.obs_pt, want[f"{key}_obs_pt"])
    for k in ("cam_RT", "focal", "pts", "obs_xy"):
        assert _close(getattr(got, k), want[f"{key}_{k}"]), (name, k)
    assert np.array_equal(got.distortion_type, want[f"{key}_distortion_type"])
    if key != "txt":
        assert np.array_equal(got.pt_rgb, want[f"{key}_pt_rgb"])


def test_loaded_bal_problem_feeds_the_engine_arrays(tmp_path):
    """BAL file -> Problem: observations sorted by point, one shared focal, reprojection error preserved by the rescale."""
    from sequencesfm_b200.synth import mean_reprojection_error_px
    p = make_problem("small", 3)
    m = bio.from_problem(p)
    rng = np.random.default_rng(0)
    perm = rng.permutation(m.K)                                  # BAL files need not be sorted by point
    m.obs_xy, m.obs_cam, m.obs_pt = m.obs_xy[perm], m.obs_cam[perm], m.obs_pt[perm]
    path = tmp_path / "problem-12-800-pre.txt"
    bio.save(path, m)
    q = bio.load(path).to_problem()
    assert np.all(np.diff(q.obs_pt) >= 0)
    key = q.obs_pt.astype(np.int64) * q.N + q.obs_cam
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(q.obs_cam, p.obs_cam) and np.array_equal(q.obs_pt, p.obs_pt)
    assert np.allclose(q.K5, [p.K5[0], 0, 0, p.K5[0], 0])
    assert abs(mean_reprojection_error_px(q) - mean_reprojection_error_px(p)) < 1e-9


def test_error_behaviour(tmp_path):
    with pytest.raises(B200BAError, match="cannot open"):
        bio.load(tmp_path / "missing.txt")
    bad = tmp_path / "bad.txt"
    bad.write_text("2 3 1\n0 7 1.0 2.0\n" + "0 0 0 0 0 0 1 0 0\n" * 2 + "0 0 1\n" * 3)    # point index 7 of 3
    with pytest.raises(B200BAError, match="point index out of range"):
        bio.load(bad)
    trunc = tmp_path / "trunc.txt"
    trunc.write_text("2 3 2\n0 0 1.0 2.0\n")
    with pytest.raises(B200BAError, match="truncated"):
        bio.load(trunc)
    one = tmp_path / "one.nvm"
    one.write_text("NVM_V3\n1\nimg.jpg 500 1 0 0 0 0 0 0 0 0\n0\n")
    with pytest.raises(B200BAError, match="more than one camera"):
        bio.load(one)
    m = _model(5)
    m.obs_pt = m.obs_pt[::-1].copy()                             # not grouped by point: NVM / .out cannot express it
    with pytest.raises(B200BAError, match="not grouped by point"):
        bio.save(tmp_path / "x.nvm", m)


def test_hostile_inputs_are_errors_not_crashes(tmp_path):
    """The C entry points never throw and never cast a non-finite / out-of-range token to an index."""
    with pytest.raises(B200BAError, match="cannot open"):
        bio.load(tmp_path)                                       # a directory
    fifo = tmp_path / "pipe.txt"
    os.mkfifo(fifo)
    with pytest.raises(B200BAError, match="cannot open"):
        bio.load(fifo)                                           # unseekable: refused before any read blocks
    for i, head in enumerate(["nan 3 1", "2 inf 1", "2 3 1e30", "0x2 3 1"]):
        f = tmp_path / f"h{i}.txt"
        f.write_text(head + "\n0 0 1.0 2.0\n" + "0 0 0 0 0 0 1 0 0\n" * 2 + "0 0 1\n" * 3)
        with pytest.raises(B200BAError):
            bio.load(f)
    f = tmp_path / "idx.txt"
    f.write_text("2 3 1\n0 4294967296 1.0 2.0\n" + "0 0 0 0 0 0 1 0 0\n" * 2 + "0 0 1\n" * 3)   # would wrap to 0 as int32
    with pytest.raises(B200BAError):
        bio.load(f)
    m = _model(5)
    m.focal = m.focal.copy(); m.focal[1] = 0.0
    with pytest.raises(B200BAError, match="focal"):
        m.to_problem()


def test_io_header_symbols_exported():
    import re
    from sequencesfm_b200 import binding
    root = os.path.dirname(GOLDEN_DIR.rstrip("/")).rsplit("/tests", 1)[0]
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "b200ba_io.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(b200io_[a-z0-9_]+)\s*\(", hdr)))
    lib = binding.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(bio.ABI_SYMBOLS) == declared
