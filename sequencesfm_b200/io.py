"""ctypes binding of ``include/b200ba_io.h``: BAL / bundler-model text, VisualSFM ``.nvm`` and Bundler ``.out`` files
(reference: ``Thirdparty/pba/pba/pba_util.h:52-411``).  Host code of ``libb200ba.so``; no GPU needed.

``load(path)`` returns a :class:`Model` whose arrays are exactly what ``b200ba_set_problem`` takes; ``Model.to_problem()``
turns it into a ``synth.Problem`` (observations sorted by point, measurements rescaled to one shared focal length).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .binding import B200BAError, load as _load_lib

ABI_SYMBOLS = ["b200io_alloc", "b200io_free", "b200io_load", "b200io_save", "b200io_last_error", "b200io_sort_by_point",
               "b200io_shared_focal"]


class _CModel(C.Structure):
    _fields_ = [("N", C.c_int32), ("M", C.c_int32), ("K", C.c_int32), ("cam_RT", C.POINTER(C.c_double)),
                ("focal", C.POINTER(C.c_double)), ("radial", C.POINTER(C.c_double)), ("distortion_type", C.POINTER(C.c_int32)),
                ("pts", C.POINTER(C.c_double)), ("pt_rgb", C.POINTER(C.c_int32)), ("obs_xy", C.POINTER(C.c_double)),
                ("obs_cam", C.POINTER(C.c_int32)), ("obs_pt", C.POINTER(C.c_int32)), ("names", C.POINTER(C.c_char_p)),
                ("reserved", C.c_int32 * 4)]


@dataclass
class Model:
    cam_RT: np.ndarray            # N x 12 [R|T], x_cam = R X + T, +z forward
    focal: np.ndarray             # N
    radial: np.ndarray            # N
    distortion_type: np.ndarray   # N: 0 none, 1 projection (BAL / .out k1), -1 measurement (NVM)
    pts: np.ndarray               # M x 3
    pt_rgb: np.ndarray            # M x 3
    obs_xy: np.ndarray            # K x 2
    obs_cam: np.ndarray           # K
    obs_pt: np.ndarray            # K
    names: list = field(default_factory=list)

    @property
    def N(self): return len(self.cam_RT)
    @property
    def M(self): return len(self.pts)
    @property
    def K(self): return len(self.obs_cam)

    def to_problem(self, name: str = "file"):
        """Observations sorted by (point, camera), one shared focal length (median; measurements rescaled)."""
        from .synth import Problem
        lib = _lib()
        cm = _to_c(self)
        try:
            K5 = np.zeros(5)
            _check(lib.b200io_sort_by_point(cm))
            _check(lib.b200io_shared_focal(cm, K5.ctypes.data_as(C.c_void_p)))
            m = _from_c(cm.contents)
        finally:
            lib.b200io_free(cm)
        return Problem(K5, m.cam_RT, m.pts, m.obs_xy, m.obs_cam, m.obs_pt, name, 0)


def _lib():
    lib = _load_lib()
    if not getattr(lib, "_io_ready", False):
        lib.b200io_last_error.restype = C.c_char_p
        lib.b200io_alloc.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.POINTER(_CModel))]
        lib.b200io_free.argtypes = [C.POINTER(_CModel)]
        lib.b200io_free.restype = None
        lib.b200io_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(_CModel))]
        lib.b200io_save.argtypes = [C.c_char_p, C.POINTER(_CModel)]
        lib.b200io_sort_by_point.argtypes = [C.POINTER(_CModel)]
        lib.b200io_shared_focal.argtypes = [C.POINTER(_CModel), C.c_void_p]
        lib._io_ready = True
    return lib


def _check(rc):
    if rc != 0:
        raise B200BAError(rc, (_lib().b200io_last_error() or b"").decode())


def _from_c(c: _CModel) -> Model:
    N, M, K = c.N, c.M, c.K
    a = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].astype(dt).copy()
    names = [c.names[i].decode() if c.names[i] else "" for i in range(N)] if c.names else []
    return Model(a(c.cam_RT, 12 * N, np.float64).reshape(N, 12), a(c.focal, N, np.float64), a(c.radial, N, np.float64),
                 a(c.distortion_type, N, np.int32), a(c.pts, 3 * M, np.float64).reshape(M, 3), a(c.pt_rgb, 3 * M, np.int32).reshape(M, 3),
                 a(c.obs_xy, 2 * K, np.float64).reshape(K, 2), a(c.obs_cam, K, np.int32), a(c.obs_pt, K, np.int32), names)


def _to_c(m: Model):
    lib = _lib()
    cm = C.POINTER(_CModel)()
    _check(lib.b200io_alloc(m.N, m.M, m.K, 1 if m.names else 0, C.byref(cm)))
    c = cm.contents
    def put(p, arr, dt):
        arr = np.ascontiguousarray(arr, dt).ravel()
        if arr.size:
            C.memmove(p, arr.ctypes.data, arr.nbytes)
    put(c.cam_RT, m.cam_RT, np.float64); put(c.focal, m.focal, np.float64); put(c.radial, m.radial, np.float64)
    put(c.distortion_type, m.distortion_type, np.int32); put(c.pts, m.pts, np.float64); put(c.pt_rgb, m.pt_rgb, np.int32)
    put(c.obs_xy, m.obs_xy, np.float64); put(c.obs_cam, m.obs_cam, np.int32); put(c.obs_pt, m.obs_pt, np.int32)
    if m.names:
        libc = C.CDLL(None)
        libc.strdup.restype = C.c_void_p
        libc.strdup.argtypes = [C.c_char_p]
        arr = C.cast(c.names, C.POINTER(C.c_void_p))
        for i, s in enumerate(m.names):
            arr[i] = libc.strdup(s.encode())
    return cm


def load(path: str) -> Model:
    """``LoadModelFile`` (pba_util.h:392-411): ``*.nvm`` -> NVM, ``*.out`` -> Bundler, anything else -> BAL text."""
    lib = _lib()
    cm = C.POINTER(_CModel)()
    _check(lib.b200io_load(str(path).encode(), C.byref(cm)))
    try:
        return _from_c(cm.contents)
    finally:
        lib.b200io_free(cm)


def save(path: str, m: Model) -> None:
    lib = _lib()
    cm = _to_c(m)
    try:
        _check(lib.b200io_save(str(path).encode(), cm))
    finally:
        lib.b200io_free(cm)


def from_problem(p, focal: float | None = None) -> Model:
    """A ``synth.Problem`` (shared intrinsics, principal point already subtracted by the caller or zero) as a Model."""
    f = float(p.K5[0]) if focal is None else focal
    N, M = len(p.cam_RT), len(p.pts)
    xy = np.asarray(p.obs_xy, np.float64) - np.array([p.K5[2], p.K5[4]])
    return Model(np.asarray(p.cam_RT, np.float64).reshape(N, 12).copy(), np.full(N, f), np.zeros(N), np.ones(N, np.int32),
                 np.asarray(p.pts, np.float64).copy(), np.zeros((M, 3), np.int32), xy, np.asarray(p.obs_cam, np.int32).copy(),
                 np.asarray(p.obs_pt, np.int32).copy(), [])
