"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the reference engines built into ``oracle/_ref``.

``oracle/_ref/libssba_ref.so`` and ``libpba_ref.so`` are the UNMODIFIED SSBA-3.0 / PBA sources of
``/root/reference`` compiled by ``oracle/Makefile`` together with the small harnesses
``oracle/ssba_ref_harness.cpp`` / ``oracle/pba_ref_harness.cpp`` that replay SequenceSFM's own marshalling
(``src/sfm_ba_ssba.cpp:85-233``, ``src/sfm_ba_pba.cpp:56-166``).  Only ``tests/``, the golden generator,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def have_ssba() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libssba_ref.so"))


def have_pba() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libpba_ref.so"))


@dataclass
class SsbaResult:
    ret: int                 # what adjustBundle_ssba returns: 0 good / -1 discarded
    status: int              # 0 TIMEOUT, 1 SMALL_UPDATE, 2 CONVERGED (Math/v3d_optimization.h:16-21)
    iterations: int
    trace: np.ndarray        # rows {iter, err, lambda, new_err, rho, accepted}
    mean_px: tuple[float, float]
    seconds: float
    cam_RT: np.ndarray
    pts: np.ndarray


def ssba_adjust(p, *, round_pixels: bool = True, max_iterations: int = 50, trace: bool = True) -> SsbaResult:
    lib = C.CDLL(os.path.join(REF_DIR, "libssba_ref.so"))
    cam = np.ascontiguousarray(p.cam_RT, dtype=np.float64).copy()
    pts = np.ascontiguousarray(p.pts, dtype=np.float64).copy()
    K5 = np.ascontiguousarray(p.K5, dtype=np.float64)
    xy = np.ascontiguousarray(p.obs_xy, dtype=np.float64)
    oc = np.ascontiguousarray(p.obs_cam, dtype=np.int32)
    op = np.ascontiguousarray(p.obs_pt, dtype=np.int32)
    cap = max(4, max_iterations + 2)
    tr = np.full((cap, 6), np.nan)
    ntr = C.c_int(0); st = C.c_int(0); it = C.c_int(0); secs = C.c_double(0)
    mean = np.zeros(2)
    lib.ssba_ref_adjust.restype = C.c_int
    ret = lib.ssba_ref_adjust(C.c_int(p.N), C.c_int(p.M), C.c_int(p.K), _d(K5), _d(cam), _d(pts), _d(xy), _i(oc), _i(op),
                              C.c_int(int(round_pixels)), C.c_int(max_iterations),
                              _d(tr) if trace else None, C.c_int(cap), C.byref(ntr),
                              C.byref(st), C.byref(it), _d(mean), C.byref(secs))
    return SsbaResult(ret, st.value, it.value, tr[: ntr.value].copy(), (float(mean[0]), float(mean[1])),
                      secs.value, cam, pts)


class _noop:
    def __enter__(self): return self
    def __exit__(self, *a): pass


@dataclass
class PbaResult:
    initial_mse: float
    final_mse: float
    lm_iterations: int
    cg_iterations: int
    seconds: float
    return_code: str
    successful: int
    cam_RT: np.ndarray
    pts: np.ndarray


class _quiet_stdout:
    """PBA and SSBA print banners with std::cout / printf even at verbosity 0; keep them off the caller's stdout
    (bench.py must print exactly one JSON line)."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        C.CDLL(None).fflush(None)          # drain the C / C++ stdio buffers into /dev/null before restoring fd 1
        os.dup2(self._saved, 1)
        os.close(self._null); os.close(self._saved)


class _captured_stdout:
    """fd-level capture of what the reference prints (PBA reports its LM iterations only on std::cout)."""

    def __enter__(self):
        import sys, tempfile
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        os.dup2(self._tmp.fileno(), 1)
        self.text = ""
        return self

    def __exit__(self, *a):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode(errors="replace")
        self._tmp.close()


def pba_trace(p, *, double: bool = True, args: tuple[str, ...] = ("-schur",), threads: int = 1):
    """The reference PBA's own per-iteration report (verbosity 2, SparseBundleCPU.cpp:3880-3925) parsed into rows
    {iteration, cg steps, e = mean squared error px^2 AFTER the step (6 digits), u = damping used, r = gain ratio (nan when the
    step was rejected), accepted}; plus the PbaResult.  Inputs are rounded to float32 exactly like the harness does."""
    import re
    with _captured_stdout() as cap:
        res = pba_adjust(p, double=double, args=tuple(args) + ("-v", "2"), threads=threads, quiet=False)
    rows = []
    for line in cap.text.splitlines():
        m = re.match(r"#\s*(\d+)\s+(\d+)\s*(\S)?\s*e=\s*(\S+)\s+u=\s*(\S+)\s+r=\s*(\S+)", line)
        if not m:
            continue
        r = m.group(6)
        acc = not r.startswith("-")
        rows.append([float(m.group(1)), float(m.group(2)), float(m.group(4)), float(m.group(5)), float(r) if acc else float("nan"), 1.0 if acc else 0.0])
    return np.array(rows, dtype=np.float64).reshape(-1, 6), res, cap.text


def pba_adjust(p, *, double: bool = True, args: tuple[str, ...] = (), threads: int = 0, quiet: bool = True) -> PbaResult:
    lib = C.CDLL(os.path.join(REF_DIR, "libpba_ref.so"))
    cam = np.ascontiguousarray(p.cam_RT, dtype=np.float64).copy()
    pts = np.ascontiguousarray(p.pts, dtype=np.float64).copy()
    K5 = np.ascontiguousarray(p.K5, dtype=np.float64)
    xy = np.ascontiguousarray(p.obs_xy, dtype=np.float64)
    oc = np.ascontiguousarray(p.obs_cam, dtype=np.int32)
    op = np.ascontiguousarray(p.obs_pt, dtype=np.int32)
    argv = (C.c_char_p * max(1, len(args)))(*[a.encode() for a in args])
    stats = np.zeros(8)
    lib.pba_ref_adjust.restype = C.c_int
    ctx = _quiet_stdout() if quiet else _noop()
    with ctx:
        lib.pba_ref_adjust(C.c_int(p.N), C.c_int(p.M), C.c_int(p.K), _d(K5), None, _d(cam), _d(pts), _d(xy), _i(oc), _i(op),
                           C.c_int(-3 if double else -2), C.c_int(len(args)), argv, C.c_int(threads), _d(stats))
    code = int(stats[5])
    return PbaResult(float(stats[0]), float(stats[1]), int(stats[2]), int(stats[3]), float(stats[4]),
                     chr(code) if code > 0 else "", int(stats[6]), cam, pts)


# ---- the reference's model-file readers / writers (oracle/io_ref_harness.cpp -> oracle/_ref/libio_ref.so) -------------
def have_io() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libio_ref.so"))


def io_load(path: str) -> dict:
    """LoadModelFile of Thirdparty/pba/pba/pba_util.h:392-411 (float storage widened to float64 arrays)."""
    lib = C.CDLL(os.path.join(REF_DIR, "libio_ref.so"))
    N, M, K = C.c_int(0), C.c_int(0), C.c_int(0)
    if lib.io_ref_load(str(path).encode(), C.byref(N), C.byref(M), C.byref(K)) != 0:
        raise RuntimeError(f"reference LoadModelFile failed on {path}")
    N, M, K = N.value, M.value, K.value
    out = {"cam_RT": np.zeros((N, 12)), "focal": np.zeros(N), "radial": np.zeros(N), "distortion_type": np.zeros(N, np.int32),
           "pts": np.zeros((M, 3)), "pt_rgb": np.zeros((M, 3), np.int32), "obs_xy": np.zeros((K, 2)),
           "obs_cam": np.zeros(K, np.int32), "obs_pt": np.zeros(K, np.int32)}
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.io_ref_copy(vp(out["cam_RT"]), vp(out["focal"]), vp(out["radial"]), vp(out["distortion_type"]), vp(out["pts"]), vp(out["pt_rgb"]),
                    vp(out["obs_xy"]), vp(out["obs_cam"]), vp(out["obs_pt"]))
    return out


def io_save(path: str, kind: int, m) -> None:
    """SaveBundlerModel (kind 0) / SaveNVM (1) / SaveBundlerOut (2) of pba_util.h on the arrays of a sequencesfm_b200.io.Model."""
    lib = C.CDLL(os.path.join(REF_DIR, "libio_ref.so"))
    vp = lambda a, dt: np.ascontiguousarray(a, dt).ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(m.cam_RT, np.float64), np.ascontiguousarray(m.focal, np.float64), np.ascontiguousarray(m.radial, np.float64),
            np.ascontiguousarray(m.distortion_type, np.int32), np.ascontiguousarray(m.pts, np.float64), np.ascontiguousarray(m.pt_rgb, np.int32),
            np.ascontiguousarray(m.obs_xy, np.float64), np.ascontiguousarray(m.obs_cam, np.int32), np.ascontiguousarray(m.obs_pt, np.int32)]
    lib.io_ref_save(str(path).encode(), kind, len(m.cam_RT), len(m.pts), len(m.obs_cam), *[a.ctypes.data_as(C.c_void_p) for a in keep])
