"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the post-BA stage checkers.

``ref_*``  -> ``oracle/_ref/libpost_ref.so``: the reference's own ``PointsPlaneFit_RANSAC`` / ``CalcCameraGround_axes`` /
              ``TransCloudPoint`` / ``TransCamera`` (``src/sfm_utils.cpp:792-1259``) compiled in place by ``oracle/Makefile``
              through ``oracle/post_ref_harness.cpp``.
``oracle_*`` -> ``oracle/_build/libpost_oracle.so``: the plain-C restatement ``oracle/post_oracle.c``.
Only ``tests/``, the golden generator, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref", "libpost_ref.so")
_ORC = os.path.join(_HERE, "_build", "libpost_oracle.so")
_vp = C.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp)


def have_ref() -> bool:
    return os.path.exists(_REF)


def _orc():
    if not os.path.exists(_ORC):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    lib = C.CDLL(_ORC)
    lib.post_oracle_plane_fit.argtypes = [C.c_int, _vp, C.c_int, _vp, C.c_double] + [_vp] * 8
    return lib


@dataclass
class PlaneFit:
    plane: np.ndarray          # [nx, ny, nz, -n.mean]
    plane_se3: np.ndarray      # 3x4
    keep: np.ndarray           # cpFilter, int32 0/1
    best_hypothesis: int = -1
    sigma2: float = float("nan")
    n_inliers: int = -1
    hyp_sigma2: np.ndarray | None = None
    hyp_score: np.ndarray | None = None


def rand_stream(seed: int, n: int) -> np.ndarray:
    """A replayable stand-in for glibc rand(): n values in [0, 2^31-1)."""
    return np.random.default_rng(seed).integers(0, 2**31 - 1, n).astype(np.int32)


def oracle_draw_triples(M: int, n_ransac: int, stream: np.ndarray):
    lib = _orc()
    tri = np.zeros((n_ransac, 3), np.int32)
    used = C.c_int(0)
    rc = lib.post_oracle_draw_triples(M, n_ransac, _p(stream), len(stream), _p(tri), C.byref(used))
    if rc != 0:
        raise RuntimeError(f"post_oracle_draw_triples rc={rc}")
    return tri, used.value


def oracle_plane_fit(pts: np.ndarray, triples: np.ndarray, k_threshold: float = 2.0) -> PlaneFit:
    lib = _orc()
    pts = np.ascontiguousarray(pts, np.float64)
    tri = np.ascontiguousarray(triples, np.int32)
    M, H = len(pts), len(tri)
    plane, se3, keep = np.zeros(4), np.zeros(12), np.zeros(M, np.int32)
    best, ninl, s2 = C.c_int(-1), C.c_int(0), C.c_double(0)
    hs, hc = np.zeros(H), np.zeros(H)
    rc = lib.post_oracle_plane_fit(M, _p(pts), H, _p(tri), k_threshold, _p(plane), _p(se3), _p(keep),
                                   C.addressof(best), C.addressof(s2), C.addressof(ninl), _p(hs), _p(hc))
    if rc != 0:
        raise RuntimeError(f"post_oracle_plane_fit rc={rc}")
    return PlaneFit(plane, se3.reshape(3, 4), keep, best.value, s2.value, ninl.value, hs, hc)


def oracle_ground_axes(cam: np.ndarray, plane: np.ndarray, plane_se3: np.ndarray) -> np.ndarray:
    lib = _orc()
    cam = np.ascontiguousarray(cam, np.float64).reshape(12)
    out = np.zeros(12)
    lib.post_oracle_ground_axes(_p(cam), _p(np.ascontiguousarray(plane, np.float64)),
                                _p(np.ascontiguousarray(plane_se3, np.float64).reshape(12)), _p(out))
    return out.reshape(3, 4)


def oracle_transform(axes: np.ndarray, pts: np.ndarray, cams: np.ndarray):
    lib = _orc()
    pts = np.ascontiguousarray(pts, np.float64).copy()
    cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 12).copy()
    lib.post_oracle_transform(_p(np.ascontiguousarray(axes, np.float64).reshape(12)), len(pts), _p(pts), len(cams), _p(cams))
    return pts, cams


# ---- the reference itself -------------------------------------------------------------------------------------------
def ref_plane_fit(pts: np.ndarray, stream: np.ndarray, n_ransac: int = 100, k_threshold: float = 2.0):
    lib = C.CDLL(_REF)
    pts = np.ascontiguousarray(pts, np.float64)
    M = len(pts)
    plane, se3, keep = np.zeros(4), np.zeros(12), np.zeros(M, np.int32)
    used = C.c_int(0)
    rc = lib.post_ref_plane_fit(M, _p(pts), n_ransac, C.c_double(k_threshold), _p(stream), len(stream), C.byref(used),
                                _p(plane), _p(se3), _p(keep))
    if rc != 0:
        raise RuntimeError(f"reference PointsPlaneFit_RANSAC rc={rc}")
    return PlaneFit(plane, se3.reshape(3, 4), keep), used.value


def ref_ground_axes(cam, plane, plane_se3) -> np.ndarray:
    lib = C.CDLL(_REF)
    out = np.zeros(12)
    lib.post_ref_ground_axes(_p(np.ascontiguousarray(cam, np.float64).reshape(12)), _p(np.ascontiguousarray(plane, np.float64)),
                             _p(np.ascontiguousarray(plane_se3, np.float64).reshape(12)), _p(out))
    return out.reshape(3, 4)


def ref_transform(axes, pts, cams):
    lib = C.CDLL(_REF)
    pts = np.ascontiguousarray(pts, np.float64).copy()
    cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 12).copy()
    lib.post_ref_transform(_p(np.ascontiguousarray(axes, np.float64).reshape(12)), len(pts), _p(pts), len(cams), _p(cams))
    return pts, cams
