"""ctypes binding of ``include/b200ba_post.h``: the post-BA ground-plane stage of ``libb200ba.so``.

Python mirror of the four reference functions the tracker runs after every bundle adjustment
(``src/sfm_tracker.cpp:876-908``): ``PointsPlaneFit_RANSAC``, ``CalcCameraGround_axes``, ``TransCloudPoint``,
``TransCamera`` (``src/sfm_utils.cpp:807-1259``), same argument meaning and error behaviour.  Everything per-point runs
in the CUDA kernels of ``csrc/post_stage.cu``; there is no CPU fallback (``PostStage()`` raises without a device).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .binding import B200BAError, load

# every symbol include/b200ba_post.h declares
ABI_SYMBOLS = [
    "b200post_create", "b200post_destroy", "b200post_last_error", "b200post_draw_triples", "b200post_plane_fit",
    "b200post_ground_axes", "b200post_transform", "b200post_ground_stage", "b200post_ground_stage_ba", "b200post_debug_time_kernel",
]
KERNELS = {0: "k_post_hist (exponent pass)", 1: "k_post_gather", 2: "k_post_score", 3: "k_post_mask"}


class Plane(C.Structure):
    _fields_ = [("plane", C.c_double * 4), ("plane_se3", C.c_double * 12), ("sigma2", C.c_double), ("score", C.c_double),
                ("inlier_mean", C.c_double * 3), ("inlier_scatter", C.c_double * 9), ("best_hypothesis", C.c_int32),
                ("n_inliers", C.c_int32), ("n_kept", C.c_int32), ("n_degenerate", C.c_int32), ("select_fallback", C.c_int32),
                ("gpu_launches", C.c_int32), ("ms_device", C.c_double), ("ms_kernels", C.c_double), ("reserved", C.c_int32 * 4)]


@dataclass
class PlaneFit:
    plane: np.ndarray           # Vec4d plane
    plane_se3: np.ndarray       # Matx34d plane_se3 (3x4)
    keep: np.ndarray            # cpFilter (int32 0/1)
    sigma2: float
    score: float
    best_hypothesis: int
    n_inliers: int
    n_kept: int
    n_degenerate: int
    select_fallback: int
    gpu_launches: int
    ms_device: float
    ms_kernels: float
    inlier_mean: np.ndarray
    inlier_scatter: np.ndarray


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _lib():
    lib = load()
    if not getattr(lib, "_post_ready", False):
        lib.b200post_last_error.restype = C.c_char_p
        lib.b200post_last_error.argtypes = [C.c_void_p]
        lib.b200post_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        lib.b200post_destroy.argtypes = [C.c_void_p]
        lib.b200post_destroy.restype = None
        lib.b200post_draw_triples.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        lib.b200post_plane_fit.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_double,
                                           C.POINTER(Plane), C.c_void_p]
        lib.b200post_ground_axes.argtypes = [C.c_void_p] * 4
        lib.b200post_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        lib.b200post_ground_stage.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                              C.c_double, C.POINTER(Plane), C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
        lib.b200post_ground_stage_ba.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.POINTER(Plane), C.c_void_p,
                                                 C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p]
        lib.b200post_debug_time_kernel.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
        lib._post_ready = True
    return lib


def draw_triples(M: int, n_ransac: int = 100) -> np.ndarray:
    """The reference's hypothesis draw, consuming the C library's rand() stream (``src/sfm_utils.cpp:837-845``)."""
    tri = np.zeros((n_ransac, 3), np.int32)
    rc = _lib().b200post_draw_triples(M, n_ransac, _vp(tri))
    if rc != 0:
        raise B200BAError(rc, "b200post_draw_triples: needs M >= 3")
    return tri


def ground_axes(cam_RT: np.ndarray, plane: np.ndarray, plane_se3: np.ndarray) -> np.ndarray:
    """``CalcCameraGround_axes`` (``src/sfm_utils.cpp:1098-1146``): host arithmetic on 28 scalars."""
    cam = np.ascontiguousarray(cam_RT, np.float64).reshape(12)
    out = np.zeros(12)
    rc = _lib().b200post_ground_axes(_vp(cam), _vp(np.ascontiguousarray(plane, np.float64)),
                                     _vp(np.ascontiguousarray(plane_se3, np.float64).reshape(12)), _vp(out))
    if rc != 0:
        raise B200BAError(rc, "b200post_ground_axes")
    return out.reshape(3, 4)


class PostStage:
    """One handle = one CUDA device + grow-only scratch; reuse it across keyframes."""

    def __init__(self, device: int = -1, select_cap: int = 0):
        self._lib = _lib()
        self._h = C.c_void_p()
        rc = self._lib.b200post_create(device, select_cap, C.byref(self._h))
        if rc != 0:
            raise B200BAError(rc, (self._lib.b200post_last_error(None) or b"").decode())

    def close(self):
        if self._h:
            self._lib.b200post_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B200BAError(rc, (self._lib.b200post_last_error(self._h) or b"").decode())

    @staticmethod
    def _result(pl: Plane, keep) -> PlaneFit:
        return PlaneFit(np.array(pl.plane), np.array(pl.plane_se3).reshape(3, 4), keep, pl.sigma2, pl.score, pl.best_hypothesis,
                        pl.n_inliers, pl.n_kept, pl.n_degenerate, pl.select_fallback, pl.gpu_launches, pl.ms_device,
                        pl.ms_kernels, np.array(pl.inlier_mean), np.array(pl.inlier_scatter).reshape(3, 3))

    def plane_fit(self, pts: np.ndarray, triples: np.ndarray, k_threshold: float = 2.0) -> PlaneFit:
        """``PointsPlaneFit_RANSAC`` with the hypotheses handed in (``draw_triples``)."""
        pts = np.ascontiguousarray(pts, np.float64)
        tri = np.ascontiguousarray(triples, np.int32)
        keep = np.zeros(len(pts), np.int32)
        pl = Plane()
        self._check(self._lib.b200post_plane_fit(self._h, len(pts), _vp(pts), len(tri), _vp(tri), k_threshold, C.byref(pl), _vp(keep)))
        return self._result(pl, keep)

    def transform(self, axes_se3: np.ndarray, pts: np.ndarray, cam_RT: np.ndarray):
        """``TransCloudPoint`` + ``TransCamera``; returns new arrays."""
        axes = np.ascontiguousarray(axes_se3, np.float64).reshape(12)
        p = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
        c = np.ascontiguousarray(cam_RT, np.float64).reshape(-1, 12).copy()
        self._check(self._lib.b200post_transform(self._h, _vp(axes), len(p), _vp(p) if len(p) else None, len(c), _vp(c) if len(c) else None))
        return p, c

    def ground_stage(self, pts: np.ndarray, cam_RT: np.ndarray, triples: np.ndarray, k_threshold: float = 2.0):
        """The whole stage of ``src/sfm_tracker.cpp:884-904``; returns (fit, kept transformed points, cameras, axes)."""
        p = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
        c = np.ascontiguousarray(cam_RT, np.float64).reshape(-1, 12).copy()
        tri = np.ascontiguousarray(triples, np.int32)
        keep = np.zeros(len(p), np.int32)
        axes = np.zeros(12)
        m_out = C.c_int32(0)
        pl = Plane()
        self._check(self._lib.b200post_ground_stage(self._h, len(p), _vp(p), len(c), _vp(c), len(tri), _vp(tri), k_threshold,
                                                    C.byref(pl), _vp(keep), C.byref(m_out), _vp(axes)))
        return self._result(pl, keep), p[:m_out.value].copy(), c, axes.reshape(3, 4)

    def ground_stage_ba(self, engine, triples: np.ndarray, k_threshold: float = 2.0):
        """The stage chained onto a finished ``binding.Engine`` without the points leaving the device
        (``b200post_ground_stage_ba``); returns (fit, kept transformed points, adjusted + re-axed cameras, axes)."""
        tri = np.ascontiguousarray(triples, np.int32)
        p = np.zeros((engine.M, 3))
        c = np.zeros((engine.N, 12))
        keep = np.zeros(engine.M, np.int32)
        axes = np.zeros(12)
        m_out = C.c_int32(0)
        pl = Plane()
        self._check(self._lib.b200post_ground_stage_ba(self._h, engine.h, len(tri), _vp(tri), k_threshold, C.byref(pl), _vp(keep),
                                                       C.byref(m_out), _vp(p), _vp(c), _vp(axes)))
        return self._result(pl, keep), p[:m_out.value].copy(), c, axes.reshape(3, 4)

    def time_kernel(self, which: int, reps: int = 10) -> float:
        ms = C.c_double(0)
        self._check(self._lib.b200post_debug_time_kernel(self._h, which, reps, C.byref(ms)))
        return ms.value
