"""TEST INFRASTRUCTURE ONLY: ctypes binding of ``oracle/ba_oracle.c`` (the plain-C CPU restatement).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline legs may import this module.
The shared object is built by ``oracle/Makefile`` (``make -C oracle oracle``) into ``oracle/_build``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libba_oracle.so")
TRACE_COLS = 8

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Options(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("min_iterations", C.c_int), ("tau", C.c_double),
                ("inlier_px", C.c_double), ("round_pixels", C.c_int), ("normalize", C.c_int),
                ("fix_first_camera", C.c_int), ("solver", C.c_int), ("cg_max_steps", C.c_int),
                ("cg_rel_tol", C.c_double), ("precond", C.c_int), ("gradient_threshold", C.c_double),
                ("update_threshold", C.c_double)]


class Report(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("ret", C.c_int), ("n_trace", C.c_int),
                ("mean_px", C.c_double * 2), ("scale", C.c_double), ("mean", C.c_double * 3),
                ("param_length", C.c_double), ("total_cg_steps", C.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ba_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ba_oracle_solve.restype = C.c_int
        _lib.ba_oracle_linearize.restype = C.c_int
    return _lib


def default_options(**kw) -> Options:
    o = Options()
    lib().ba_oracle_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


@dataclass
class OracleResult:
    ret: int
    status: int
    iterations: int
    trace: np.ndarray        # rows {iter, err, lambda, new_err, rho, accepted, cg_steps, cg_rel}
    mean_px: tuple[float, float]
    cam_RT: np.ndarray
    pts: np.ndarray
    total_cg_steps: int = 0
    extra: dict = field(default_factory=dict)


def solve(p, **opts) -> OracleResult:
    o = default_options(**opts)
    cam = np.ascontiguousarray(p.cam_RT, dtype=np.float64).copy()
    pts = np.ascontiguousarray(p.pts, dtype=np.float64).copy()
    K5 = np.ascontiguousarray(p.K5, dtype=np.float64)
    xy = np.ascontiguousarray(p.obs_xy, dtype=np.float64)
    oc = np.ascontiguousarray(p.obs_cam, dtype=np.int32)
    op = np.ascontiguousarray(p.obs_pt, dtype=np.int32)
    cap = o.max_iterations + 2
    tr = np.full((cap, TRACE_COLS), np.nan)
    rep = Report()
    ret = lib().ba_oracle_solve(C.byref(o), C.c_int(p.N), C.c_int(p.M), C.c_int(p.K), _d(K5), _d(cam), _d(pts),
                                _d(xy), _i(oc), _i(op), _d(tr), C.c_int(cap), C.byref(rep))
    return OracleResult(ret, rep.status, rep.iterations, tr[: rep.n_trace].copy(), (rep.mean_px[0], rep.mean_px[1]),
                        cam, pts, rep.total_cg_steps,
                        {"scale": rep.scale, "mean": tuple(rep.mean), "param_length": rep.param_length})


def linearize(p, x: np.ndarray | None = None, lam: float = 0.0, **opts) -> dict:
    """One linearisation at the problem's state (normalised like the solve); optionally S(lam) @ x."""
    o = default_options(**opts)
    N, M, K = p.N, p.M, p.K
    K5 = np.ascontiguousarray(p.K5, dtype=np.float64)
    cam = np.ascontiguousarray(p.cam_RT, dtype=np.float64)
    pts = np.ascontiguousarray(p.pts, dtype=np.float64)
    xy = np.ascontiguousarray(p.obs_xy, dtype=np.float64)
    oc = np.ascontiguousarray(p.obs_cam, dtype=np.int32)
    op = np.ascontiguousarray(p.obs_pt, dtype=np.int32)
    out = {"w": np.zeros(K), "gc": np.zeros((N, 6)), "gp": np.zeros((M, 3)), "U": np.zeros((N, 6, 6)),
           "V": np.zeros((M, 3, 3)), "Xn": np.zeros((M, 3)), "RTn": np.zeros((N, 12))}
    cost = C.c_double(0); lam0 = C.c_double(0)
    Sx = None
    if x is not None:
        x = np.ascontiguousarray(x, dtype=np.float64)
        Sx = np.zeros(6 * N)
    lib().ba_oracle_linearize(C.byref(o), C.c_int(N), C.c_int(M), C.c_int(K), _d(K5), _d(cam), _d(pts), _d(xy), _i(oc), _i(op),
                              C.byref(cost), _d(out["w"]), _d(out["gc"]), _d(out["gp"]), _d(out["U"]), _d(out["V"]),
                              C.byref(lam0), _d(x), C.c_double(lam), _d(Sx), _d(out["Xn"]), _d(out["RTn"]))
    out["cost"] = cost.value
    out["lambda0"] = lam0.value
    if Sx is not None:
        out["Sx"] = Sx.reshape(N, 6)
    return out
