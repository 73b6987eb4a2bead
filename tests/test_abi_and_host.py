"""CPU: the C-ABI library loads and exports every symbol include/b200ba.h declares; host-side logic (synthetic
generator, marshalling mirror of src/sfm_ba_ssba.cpp:88-182, point partition)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sequencesfm_b200 import ba, binding
from sequencesfm_b200.synth import SHAPES, make_problem, mean_reprojection_error_px, partition_points

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="b200ba.h", prefix="b200ba_"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = binding.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), f"libb200ba.so does not export {sym}"
    assert sorted(binding.ABI_SYMBOLS) == declared
    # the post-BA stage header
    from sequencesfm_b200 import post
    declared_post = _declared_symbols("b200ba_post.h", "b200post_")
    assert len(declared_post) >= 9
    for sym in declared_post:
        assert hasattr(lib, sym), f"libb200ba.so does not export {sym}"
    assert sorted(post.ABI_SYMBOLS) == declared_post
    from sequencesfm_b200 import io as bio
    declared_io = _declared_symbols("b200ba_io.h", "b200io_")
    for sym in declared_io:
        assert hasattr(lib, sym), f"libb200ba.so does not export {sym}"
    assert sorted(bio.ABI_SYMBOLS) == declared_io
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["b200ba.h", "b200ba_io.h", "b200ba_post.h", "b200ba_tri.h"]


def test_options_struct_layout_and_defaults():
    o = binding.default_options()
    assert o.struct_size == C.sizeof(binding.Options)
    # the constants hard-coded at the seam, src/sfm_ba_ssba.cpp:193-217
    assert (o.max_iterations, o.min_iterations, o.tau, o.inlier_px) == (50, 10, 1e-3, 2.0)
    assert (o.round_pixels, o.normalize, o.fix_first_camera, o.discard_converged) == (1, 1, 0, 1)
    assert (o.cg_max_steps, o.precond) == (50, 1)
    p = binding.default_options("pba")
    assert (p.fix_first_camera, p.inlier_px, p.round_pixels, p.discard_converged) == (1, 0.0, 0, 0)


def test_no_cuda_means_loud_failure(have_cuda):
    """No CPU fallback: without a device, create() must fail with B200BA_E_CUDA, not compute on the host."""
    if have_cuda:
        pytest.skip("CUDA present")
    with pytest.raises(binding.B200BAError) as e:
        binding.Engine()
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sequencesfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").replace("the CPU oracle", "") or f == "synth.py", f
    for f in ("adapter/sfm_ba_b200.cpp", "include/b200ba.h"):
        assert "ba_oracle" not in open(os.path.join(ROOT, f)).read()


@pytest.mark.parametrize("shape", ["tiny", "small", "ladybug"])
def test_synthetic_problem_properties(shape):
    p = make_problem(shape, 5)
    N, M, K = SHAPES[shape]
    assert (p.N, p.M, p.K) == (N, M, K)
    assert np.all(np.diff(p.obs_pt) >= 0), "observations must be sorted by point (src/sfm_ba_ssba.cpp:161-182)"
    key = p.obs_pt.astype(np.int64) * N + p.obs_cam
    assert np.all(np.diff(key) > 0), "cameras distinct and ascending inside a point (std::map order)"
    deg = np.bincount(p.obs_pt, minlength=M)
    assert deg.min() >= 2
    assert p.obs_xy.dtype == np.float64 and np.array_equal(p.obs_xy, p.obs_xy.astype(np.float32).astype(np.float64))
    assert 0.3 < mean_reprojection_error_px(p) < 20
    q = make_problem(shape, 5)
    assert np.array_equal(p.obs_xy, q.obs_xy) and np.array_equal(p.cam_RT, q.cam_RT)


def test_partition_points_balanced_and_covering():
    p = make_problem("ladybug", 2)
    for world in (1, 2, 3, 4, 8):
        parts = partition_points(p.obs_pt, p.M, world)
        assert parts[0][0] == 0 and parts[-1][1] == p.M
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        counts = [int(np.searchsorted(p.obs_pt, b) - np.searchsorted(p.obs_pt, a)) for a, b in parts]
        assert sum(counts) == p.K
        assert max(counts) - min(counts) <= 32
        sub = p.slice_points(*parts[-1])
        assert sub.K == counts[-1] and sub.obs_pt.min() == 0 and sub.obs_pt.max() == sub.M - 1


def _tracker_containers(p, frame_of=lambda i: 7 * i + 3):
    """Flat problem -> the containers the tracker hands to adjustBundle_* (sparse frame ids exercise map order)."""
    Kmat = np.array([[p.K5[0], p.K5[1], p.K5[2]], [0, p.K5[3], p.K5[4]], [0, 0, 1.0]])
    Pmats = {frame_of(i): p.cam_RT[i].reshape(3, 4).copy() for i in range(p.N)}
    imgpts = {frame_of(i): [] for i in range(p.N)}
    cloud = [ba.CloudPoint(pt=p.pts[j].copy()) for j in range(p.M)]
    for k in range(p.K):
        f = frame_of(int(p.obs_cam[k]))
        cloud[int(p.obs_pt[k])].img_kp[f] = len(imgpts[f])
        imgpts[f].append((float(p.obs_xy[k, 0]), float(p.obs_xy[k, 1])))
    return cloud, Kmat, imgpts, Pmats


def test_marshal_reproduces_reference_order():
    p = make_problem("tiny", 4)
    cloud, Kmat, imgpts, Pmats = _tracker_containers(p)
    # shuffle dict insertion order: marshalling must follow key order like std::map, not insertion order
    Pm = dict(reversed(list(Pmats.items())))
    K5, cam_RT, pts, xy, oc, op, ids = ba.marshal(cloud, Kmat, imgpts, Pm)
    assert ids == sorted(Pmats.keys())
    assert np.array_equal(K5, p.K5) and np.array_equal(cam_RT, p.cam_RT) and np.array_equal(pts, p.pts)
    assert np.array_equal(oc, p.obs_cam) and np.array_equal(op, p.obs_pt) and np.array_equal(xy, p.obs_xy)


def test_python_mirror_with_oracle_as_engine(monkeypatch):
    """The host mirror's write-back / error convention, exercised on CPU by standing the oracle in for the
    engine (tests may do that; the product path never does)."""
    from oracle import oracle
    from sequencesfm_b200.binding import SolveResult

    class FakeEngine:
        def __init__(self, opt): self.opt = opt
        def __enter__(self): return self
        def __exit__(self, *a): pass
        def set_problem(self, K5, cam, pts, xy, oc, op):
            from sequencesfm_b200.synth import Problem
            self.p = Problem(K5, cam, pts, xy, oc, op)
        def solve(self):
            r = oracle.solve(self.p, max_iterations=self.opt.max_iterations, inlier_px=self.opt.inlier_px,
                             round_pixels=self.opt.round_pixels, fix_first_camera=self.opt.fix_first_camera,
                             gradient_threshold=self.opt.gradient_threshold, solver=1)
            ret = -1 if (self.opt.discard_converged and r.status == 2) else 0
            out = SolveResult(ret, r.status, r.iterations, r.trace, r.mean_px, r.cam_RT, r.pts, 0, 0, 0, 0, 0, 0, 0)
            out.opt_seen = self.opt
            return out

    monkeypatch.setattr(binding, "Engine", FakeEngine)
    p = make_problem("tiny", 6)
    cloud, Kmat, imgpts, Pmats = _tracker_containers(p)
    before = {k: v.copy() for k, v in Pmats.items()}
    rep = {}
    assert ba.adjustBundle_ssba(cloud, Kmat, imgpts, Pmats, report=rep) == 0
    res = rep["result"]
    assert res.mean_px[1] < res.mean_px[0]
    assert any(not np.array_equal(before[k], Pmats[k]) for k in Pmats)
    assert np.array_equal(cloud[0].pt, res.pts[0])
    # a gradient-converged outcome is discarded (src/sfm_ba_ssba.cpp:221): inputs untouched, -1 returned
    cloud2, Kmat2, imgpts2, Pmats2 = _tracker_containers(p)
    snap = {k: v.copy() for k, v in Pmats2.items()}
    assert ba.adjustBundle_ssba(cloud2, Kmat2, imgpts2, Pmats2, options={"gradient_threshold": 1e30}) == -1
    assert all(np.array_equal(snap[k], Pmats2[k]) for k in snap)
    assert np.array_equal(cloud2[0].pt, p.pts[0])
    # Shipped build (USE_PBA undefined): adjustBundle_pba forwards to adjustBundle_ssba and returns ITS value
    # (src/sfm_ba_pba.cpp:167-171): same results, and -1 with untouched inputs on a gradient-converged outcome
    cloud3, Kmat3, imgpts3, Pmats3 = _tracker_containers(p)
    rep3 = {}
    assert ba.adjustBundle_pba(cloud3, Kmat3, np.zeros(4), imgpts3, Pmats3, report=rep3) == 0
    assert rep3["result"].opt_seen.round_pixels == 1 and rep3["result"].opt_seen.fix_first_camera == 0
    assert all(np.array_equal(Pmats[k], Pmats3[k]) for k in Pmats) and np.array_equal(cloud[5].pt, cloud3[5].pt)
    cloud4, Kmat4, imgpts4, Pmats4 = _tracker_containers(p)
    assert ba.adjustBundle_pba(cloud4, Kmat4, np.zeros(4), imgpts4, Pmats4, options={"gradient_threshold": 1e30}) == -1
    assert all(np.array_equal(snap[k], Pmats4[k]) for k in snap)
    # a -DUSE_PBA build: PBA personality (camera 0 constant, sub-pixel, no robustifier), always 0 (src/sfm_ba_pba.cpp:173)
    cloud5, Kmat5, imgpts5, Pmats5 = _tracker_containers(p)
    rep5 = {}
    assert ba.adjustBundle_pba(cloud5, Kmat5, np.zeros(4), imgpts5, Pmats5, options={"gradient_threshold": 1e30}, report=rep5, use_pba=True) == 0
    assert rep5["result"].opt_seen.round_pixels == 0 and rep5["result"].opt_seen.fix_first_camera == 1


def test_bench_byte_model_is_the_surveys():
    """bench.py's algorithmic bytes are SURVEY.md section 8(d)'s model: fp64 Venice shape 216 MB linearisation, 335.5 MB per
    PCG step, 17.34 GB per LM iteration; the per-kernel split of a PCG step adds up to the step."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, M, K = 1778, 993923, 5001946
    ab = bench.algorithmic_bytes(N, M, K)
    assert abs(ab["lin"] - 216.06e6) < 0.1e6
    assert abs(ab["cg_step"] - 335.51e6) < 0.1e6
    assert abs(ab["iteration"] - 17.342e9) < 0.01e9
    km = bench.kernel_models(N, M, K)
    assert km["cg_point_pass"] + km["cg_camera_pass"] == ab["cg_step"]
    assert km["linearize"] == ab["lin"]


def test_layout_conflict_factor():
    """Host side of b200ba_set_problem, no device needed (b200ba_debug_layout_stats): grouping points into quarter-warps and
    colouring their observations over the slots brings the shared-memory camera-table reads of the by-point sweeps from ~2.5
    wavefronts per quarter-warp phase (random cameras) to nearly conflict-free; the slot assignment is a permutation of every
    point's observations (checked inside the call) and the SELL padding stays negligible."""
    p = make_problem("trafalgar", 1)
    st = binding.layout_stats(p, 64)
    assert st["conflict_factor_plain"] > 2.2 and st["conflict_factor"] < 1.2, st
    assert st["padding"] < 1.01
    off = binding.layout_stats(p, 0)                                     # lane_window = 0: input order kept
    assert abs(off["conflict_factor"] - off["conflict_factor_plain"]) < 1e-12
    # tiny and degenerate graphs go through the same code: every point seen by the same two cameras, long tracks
    q = make_problem((3, 40, 120), 2)
    assert binding.layout_stats(q, 64)["conflict_factor"] <= binding.layout_stats(q, 0)["conflict_factor"] + 1e-12
    r = make_problem((400, 64, 64 * 50), 3)                              # track length ~50 > COLOUR_DMAX: input order kept, still valid
    assert binding.layout_stats(r, 64)["padding"] < 1.5


def test_explicit_reduced_system_pair_lists():
    """Host side of the explicit reduced camera system (csrc/dense_host.h, no device needed): every ordered pair of observations
    of one point with camera(a) <= camera(b) appears exactly once, filed under its destination block, pairs of a block in point
    order; the kernel shape follows the camera count (cluster with S in registers, cluster with S in shared memory, whole grid,
    matrix-free)."""
    rng = np.random.default_rng(5)
    p = make_problem((9, 60, 260), 3)
    oc, op = p.obs_cam.copy(), p.obs_pt.copy()
    # a point seen twice by the same camera: both orders of the two observations go to the diagonal block
    k0 = int(np.flatnonzero(op == op[10])[0]); oc[k0 + 1] = oc[k0]
    q = type(p)(p.K5, p.cam_RT, p.pts, p.obs_xy, oc, op, "dup", 0)
    dl = binding.dense_lists(q)
    want = []
    for j in range(q.M):
        ks = np.flatnonzero(op == j)
        for a in ks:
            for b in ks:
                if oc[a] <= oc[b]:
                    i, k = int(oc[a]), int(oc[b])
                    want.append((i * q.N - i * (i - 1) // 2 + (k - i), int(a), int(b)))
    want.sort(key=lambda t: t[0])                                        # stable: point order inside a block
    got = list(zip(dl["block"].tolist(), dl["a"].tolist(), dl["b"].tolist()))
    assert got == want
    assert dl["pairs"] == len(want) and dl["blocks"] == q.N * (q.N + 1) // 2
    assert dl["chunks"] >= dl["pairs"] // 32
    dup = [(a, b) for _, a, b in want if oc[a] == oc[b] and a != b]
    assert (k0, k0 + 1) in dup and (k0 + 1, k0) in dup
    # shapes by camera count
    kinds = {n: binding.dense_lists(make_problem((n, 40, 120), 1))["shape"] for n in (5, 49, 53, 54, 90, 120, 257, 296, 400)}
    assert [kinds[n]["kind"] for n in (5, 49, 53)] == [0, 0, 0] and kinds[49]["ctas"] == 8 and kinds[49]["rows"] == 42
    assert kinds[54]["kind"] == 1 and kinds[54]["ctas"] == 8 and kinds[90]["kind"] == 1 and kinds[90]["ctas"] == 16
    assert kinds[120]["kind"] == 2 and kinds[120]["ctas"] == 60 and kinds[257]["kind"] == 2 and kinds[257]["ctas"] == 129 and kinds[257]["rows"] == 12
    assert kinds[296]["kind"] == 2 and kinds[400]["kind"] == -1         # more than two cameras per multiprocessor: matrix-free
    for s in kinds.values():
        assert s["smem"] <= 227 * 1024 and (s["kind"] < 0 or s["rows"] % 6 == 0)


def test_committed_bench_line_follows_the_contract():
    """The bench line committed as closing evidence (profiles/bench_r2g_venice.json = one `python bench.py` run on a B200) carries
    every key the measurement contract names - metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better /
    scaling / vs_baseline / dtype / data / config.workload / clocks / e2e with its byte counts / gpu_launches / roofline / cpu_baseline -
    and its numbers are consistent with each other and with the byte model."""
    import json
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r2g_venice.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
        assert k in d, k
    assert d["metric"] == "lm_iterations_per_s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-9                      # value = LM iterations per second of the timed steps
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == d["unit"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_launch"] == 5001946 * 48 + 993923 * 96            # K (16 + 4 s) + 12 s M, s = 8
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0.9 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["parity"]["pass"] is True and d["parity"]["max_rel_cost_diff"] < d["parity"]["tolerance"] == 1e-9
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
