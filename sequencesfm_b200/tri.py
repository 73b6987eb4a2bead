"""ctypes binding of ``include/b200ba_tri.h``: batched two-view triangulation (``TriangulatePoints``,
``src/sfm_triangulation.cpp:95-202``) on the GPU.  No CPU fallback (``Triangulator()`` raises without a device)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .binding import B200BAError, load

ABI_SYMBOLS = ["b200tri_create", "b200tri_destroy", "b200tri_last_error", "b200tri_triangulate", "b200tri_gate"]


class Report(C.Structure):
    _fields_ = [("mean_reproj_error", C.c_double), ("n_degenerate", C.c_int32), ("gpu_launches", C.c_int32),
                ("ms_device", C.c_double), ("ms_kernels", C.c_double), ("reserved", C.c_int32 * 4)]


class GateReport(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_front_P", C.c_int32), ("n_front_P1", C.c_int32), ("n_kept", C.c_int32),
                ("cutoff", C.c_double), ("gpu_launches", C.c_int32), ("ms_device", C.c_double), ("reserved", C.c_int32 * 4)]


@dataclass
class Gate:
    status: int            # 0 ok, -1 / -2 / -3: the return codes of SfM_Tracker::triangulateTwoFrames
    keep: np.ndarray       # uint8 n
    cutoff: float
    n_front_P: int
    n_front_P1: int
    n_kept: int
    ms_device: float


@dataclass
class Triangulation:
    pts: np.ndarray               # n x 3   CloudPoint::pt
    reproj_err: np.ndarray        # n       CloudPoint::reprojection_error
    pt1_undistorted: np.ndarray   # n x 2   correspImg1Pt (float32)
    mean_reproj_error: float      # the function's return value
    n_degenerate: int
    gpu_launches: int
    ms_device: float
    ms_kernels: float


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Triangulator:
    def __init__(self, device: int = -1):
        self._lib = load()
        self._lib.b200tri_last_error.restype = C.c_char_p
        self._lib.b200tri_last_error.argtypes = [C.c_void_p]
        self._lib.b200tri_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        self._lib.b200tri_destroy.argtypes = [C.c_void_p]
        self._lib.b200tri_destroy.restype = None
        self._lib.b200tri_triangulate.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32] + [C.c_void_p] * 5 + [C.POINTER(Report)]
        self._lib.b200tri_gate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(GateReport)]
        self._n_last = 0
        self._h = C.c_void_p()
        rc = self._lib.b200tri_create(device, C.byref(self._h))
        if rc != 0:
            raise B200BAError(rc, (self._lib.b200tri_last_error(None) or b"").decode())

    def close(self):
        if self._h:
            self._lib.b200tri_destroy(self._h)
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

    def triangulate(self, pt1, pt2, K, P, P1, dist=None, Kinv=None) -> Triangulation:
        """``TriangulatePoints(pt_set1, pt_set2, K, Kinv, distcoeff, P, P1, pointcloud, correspImg1Pt)``."""
        pt1 = np.ascontiguousarray(pt1, np.float32).reshape(-1, 2)
        pt2 = np.ascontiguousarray(pt2, np.float32).reshape(-1, 2)
        K = np.ascontiguousarray(K, np.float64).reshape(3, 3)
        Kinv = np.linalg.inv(K) if Kinv is None else np.ascontiguousarray(Kinv, np.float64).reshape(3, 3)
        Kinv = np.ascontiguousarray(Kinv)
        P = np.ascontiguousarray(P, np.float64).reshape(3, 4)
        P1 = np.ascontiguousarray(P1, np.float64).reshape(3, 4)
        d = None if dist is None else np.ascontiguousarray(dist, np.float64).ravel()
        n = len(pt1)
        pts, err, out = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 2), np.float32)
        rep = Report()
        rc = self._lib.b200tri_triangulate(self._h, n, _vp(pt1), _vp(pt2), _vp(K), _vp(Kinv), _vp(d), 0 if d is None else len(d),
                                           _vp(P), _vp(P1), _vp(pts), _vp(err), _vp(out), C.byref(rep))
        if rc != 0:
            raise B200BAError(rc, (self._lib.b200tri_last_error(self._h) or b"").decode())
        self._n_last = n
        return Triangulation(pts, err, out, rep.mean_reproj_error, rep.n_degenerate, rep.gpu_launches, rep.ms_device, rep.ms_kernels)

    def gate(self) -> Gate:
        """The tracker's gating of the LAST triangulation (``src/sfm_tracker.cpp:718-771``), on the device-resident results."""
        keep = np.zeros(self._n_last, np.uint8)
        rep = GateReport()
        rc = self._lib.b200tri_gate(self._h, _vp(keep) if self._n_last else None, C.byref(rep))
        if rc != 0:
            raise B200BAError(rc, (self._lib.b200tri_last_error(self._h) or b"").decode())
        return Gate(rep.status, keep, rep.cutoff, rep.n_front_P, rep.n_front_P1, rep.n_kept, rep.ms_device)
