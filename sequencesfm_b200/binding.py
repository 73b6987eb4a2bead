"""ctypes binding of the C ABI in ``include/b200ba.h`` (``sequencesfm_b200/libb200ba.so``).

This is the Python face of the drop-in boundary: plain pointers and sizes cross it, nothing else.  The
library is the hand-written sm_100a CUDA engine; there is no fallback of any kind -- if the shared object
is missing or no CUDA device is usable the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200BA_LIB", os.path.join(_HERE, "libb200ba.so"))   # override: kernel-variant experiments
TRACE_COLS = 8
COMM_ID_BYTES = 128
PEER_HANDLE_BYTES = 64

STATUS_NAMES = {0: "TIMEOUT", 1: "SMALL_UPDATE", 2: "CONVERGED", 3: "MSE_THRESHOLD"}

# every symbol include/b200ba.h declares (tests check that the library exports each of them)
ABI_SYMBOLS = [
    "b200ba_default_options", "b200ba_pba_preset", "b200ba_pba_lm_preset", "b200ba_create", "b200ba_destroy", "b200ba_last_error",
    "b200ba_set_problem", "b200ba_solve", "b200ba_begin", "b200ba_iterate", "b200ba_restart", "b200ba_last_iterate_ms", "b200ba_last_restart_ms", "b200ba_finish", "b200ba_get_result",
    "b200ba_get_result_device", "b200ba_problem_size", "b200ba_get_plan", "b200ba_append", "b200ba_set_state",
    "b200ba_comm_unique_id", "b200ba_comm_init", "b200ba_comm_peer_handle", "b200ba_comm_open_peers", "b200ba_debug_linearize", "b200ba_debug_schur_matvec",
    "b200ba_debug_time_kernel", "b200ba_kernel_name", "b200ba_version", "b200ba_debug_layout_stats", "b200ba_debug_dense_lists",
]


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("max_iterations", C.c_int32),
                ("min_iterations", C.c_int32), ("tau", C.c_double), ("inlier_px", C.c_double),
                ("gradient_threshold", C.c_double), ("update_threshold", C.c_double),
                ("round_pixels", C.c_int32), ("normalize", C.c_int32), ("fix_first_camera", C.c_int32),
                ("discard_converged", C.c_int32), ("cg_max_steps", C.c_int32), ("cg_rel_tol", C.c_double),
                ("precond", C.c_int32), ("poll_every", C.c_int32), ("lane_window", C.c_int32), ("pcg_fp32", C.c_int32),
                ("keep_problem", C.c_int32), ("lm_variant", C.c_int32), ("cg_min_steps", C.c_int32), ("reduced_system", C.c_int32),
                ("reserved", C.c_int32 * 1)]


class Report(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("ret", C.c_int32), ("n_trace", C.c_int32),
                ("total_cg_steps", C.c_int32), ("gpu_launches", C.c_int32), ("mean_px", C.c_double * 2),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("ms_solve", C.c_double),
                ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("trace", C.POINTER(C.c_double)),
                ("trace_cap", C.c_int32), ("iterations_total", C.c_int32), ("accepted_total", C.c_int32), ("reserved", C.c_int32 * 3)]


class Plan(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("camera_table", C.c_int32), ("cluster_size", C.c_int32),
                ("cooperative_step", C.c_int32), ("cuda_graph", C.c_int32), ("peer_windows", C.c_int32),
                ("world", C.c_int32), ("rank", C.c_int32), ("launches_per_iteration", C.c_int32),
                ("n_slices", C.c_int32), ("padded_entries", C.c_int32), ("n_rows", C.c_int32), ("n_segments", C.c_int32),
                ("fp32_pcg", C.c_int32), ("device_bytes", C.c_int64), ("explicit_schur", C.c_int32), ("explicit_pairs", C.c_int32),
                ("reserved", C.c_int32 * 6)]


TABLE_NAMES = {0: "global", 1: "shared", 2: "cluster"}


class B200BAError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200ba error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libb200ba.so (fails loudly when it has not been built: ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built; run `make -C sequencesfm_b200/csrc` (no CPU fallback exists)")
        lib = C.CDLL(LIB_PATH)
        lib.b200ba_last_error.restype = C.c_char_p
        lib.b200ba_last_error.argtypes = [C.c_void_p]
        lib.b200ba_kernel_name.restype = C.c_char_p
        lib.b200ba_kernel_name.argtypes = [C.c_int32]
        lib.b200ba_create.argtypes = [C.POINTER(Options), C.POINTER(C.c_void_p)]
        lib.b200ba_destroy.argtypes = [C.c_void_p]
        lib.b200ba_destroy.restype = None
        _lib = lib
    return _lib


def default_options(preset: str = "ssba", **kw) -> Options:
    o = Options()
    lib = load()
    rc = (lib.b200ba_pba_lm_preset(C.byref(o)) if preset == "pba_lm" else lib.b200ba_pba_preset(C.byref(o)) if preset == "pba"
          else lib.b200ba_default_options(C.byref(o)))
    if rc:
        raise B200BAError(rc, "default options")
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def dense_lists(p, n_sm: int = 148) -> dict:
    """``b200ba_debug_dense_lists`` (host only, no device needed): the pair lists and the PCG-kernel shape of the explicit
    reduced camera system for a problem's observation graph."""
    oc = np.ascontiguousarray(p.obs_cam, dtype=np.int32); op = np.ascontiguousarray(p.obs_pt, dtype=np.int32)
    n = np.zeros(4, dtype=np.int64); shape = np.zeros(6, dtype=np.int64)
    lib = load()
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    rc = lib.b200ba_debug_dense_lists(C.c_int32(p.N), C.c_int32(p.M), C.c_int32(p.K), _ip(oc), _ip(op), lp(n), None, None, None,
                                      C.c_int64(0), C.c_int32(n_sm), lp(shape))
    if rc:
        raise B200BAError(rc, "b200ba_debug_dense_lists")
    a = np.zeros(int(n[0]), dtype=np.int32); b = np.zeros_like(a); blk = np.zeros_like(a)
    rc = lib.b200ba_debug_dense_lists(C.c_int32(p.N), C.c_int32(p.M), C.c_int32(p.K), _ip(oc), _ip(op), lp(n), _ip(a), _ip(b), _ip(blk),
                                      C.c_int64(len(a)), C.c_int32(n_sm), None)
    if rc:
        raise B200BAError(rc, "b200ba_debug_dense_lists")
    return {"pairs": int(n[0]), "chunks": int(n[1]), "blocks": int(n[2]), "a": a, "b": b, "block": blk,
            "shape": dict(zip(("kind", "ctas", "rows", "ld", "threads", "smem"), (int(v) for v in shape)))}


@dataclass
class SolveResult:
    ret: int
    status: int
    iterations: int
    trace: np.ndarray
    mean_px: tuple[float, float]
    cam_RT: np.ndarray
    pts: np.ndarray
    total_cg_steps: int
    gpu_launches: int
    ms_solve: float
    ms_h2d: float
    ms_d2h: float
    initial_cost: float
    final_cost: float
    extra: dict = field(default_factory=dict)


def layout_stats(p, lane_window: int = 64) -> dict:
    """``b200ba_debug_layout_stats`` (host only, no device needed): conflict factor of the point-order layout."""
    out = np.zeros(4)
    oc = np.ascontiguousarray(p.obs_cam, dtype=np.int32); op = np.ascontiguousarray(p.obs_pt, dtype=np.int32)
    rc = load().b200ba_debug_layout_stats(C.c_int32(p.N), C.c_int32(p.M), C.c_int32(p.K), _ip(oc), _ip(op), C.c_int32(lane_window), _dp(out))
    if rc:
        raise B200BAError(rc, "b200ba_debug_layout_stats")
    return {"conflict_factor": out[0], "conflict_factor_plain": out[1], "ms": out[2], "padding": out[3]}


class Engine:
    """Thin RAII wrapper over one ``b200ba_handle``."""

    def __init__(self, options: Options | None = None, **kw):
        self.lib = load()
        self.opt = options if options is not None else default_options(**kw)
        self.h = C.c_void_p()
        rc = self.lib.b200ba_create(C.byref(self.opt), C.byref(self.h))
        if rc:
            raise B200BAError(rc, self.lib.b200ba_last_error(None).decode())
        self.N = self.M = self.K = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.b200ba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise B200BAError(rc, self.lib.b200ba_last_error(self.h).decode())

    # -- multi-GPU ------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load().b200ba_comm_unique_id(buf)
        if rc:
            raise B200BAError(rc, load().b200ba_last_error(None).decode())
        return buf.raw

    def comm_init(self, id_bytes: bytes, rank: int, world: int):
        self._check(self.lib.b200ba_comm_init(self.h, C.c_char_p(id_bytes), C.c_int32(rank), C.c_int32(world)))

    def comm_peer_handle(self, max_cameras: int) -> bytes:
        """Allocate this rank's NVLink exchange window; returns its CUDA IPC handle (all-gather it over the launcher)."""
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._check(self.lib.b200ba_comm_peer_handle(self.h, C.c_int32(max_cameras), buf))
        return buf.raw

    def comm_open_peers(self, all_handles: bytes):
        self._check(self.lib.b200ba_comm_open_peers(self.h, C.c_char_p(all_handles)))

    def comm_setup_torch(self, rank: int, world: int, max_cameras: int, peer_windows: bool = True):
        """Launcher plumbing over an initialised torch.distributed NCCL group: broadcast the NCCL id, create the
        communicator and (optionally) exchange the NVLink peer-window handles."""
        import torch
        import torch.distributed as dist
        idt = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        self.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        if peer_windows and world > 1 and os.environ.get("B200BA_NO_P2P") is None:
            mine = torch.frombuffer(bytearray(self.comm_peer_handle(max_cameras)), dtype=torch.uint8).cuda()
            allh = [torch.zeros(PEER_HANDLE_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            self.comm_open_peers(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))

    # -- problem / solve ------------------------------------------------------------------------------
    def set_problem(self, K5, cam_RT, pts, obs_xy, obs_cam, obs_pt):
        self._K5 = np.ascontiguousarray(K5, dtype=np.float64)
        self._cam = np.ascontiguousarray(cam_RT, dtype=np.float64).reshape(-1, 12)
        self._pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(-1, 2)
        oc = np.ascontiguousarray(obs_cam, dtype=np.int32)
        op = np.ascontiguousarray(obs_pt, dtype=np.int32)
        self.N, self.M, self.K = self._cam.shape[0], self._pts.shape[0], oc.shape[0]
        self._check(self.lib.b200ba_set_problem(self.h, C.c_int32(self.N), C.c_int32(self.M), C.c_int32(self.K),
                                                _dp(self._cam), _dp(self._K5), _dp(self._pts), _dp(xy), _ip(oc), _ip(op)))

    def set_problem_from(self, p):
        self.set_problem(p.K5, p.cam_RT, p.pts, p.obs_xy, p.obs_cam, p.obs_pt)

    def append(self, cam_RT, pts, obs_xy, obs_cam, obs_pt):
        """``b200ba_append``: extend the uploaded problem (needs ``keep_problem=1``); indices in the extended numbering."""
        cam = np.ascontiguousarray(cam_RT, dtype=np.float64).reshape(-1, 12)
        pt = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(-1, 2)
        oc = np.ascontiguousarray(obs_cam, dtype=np.int32); op = np.ascontiguousarray(obs_pt, dtype=np.int32)
        self._check(self.lib.b200ba_append(self.h, C.c_int32(cam.shape[0]), _dp(cam), C.c_int32(pt.shape[0]), _dp(pt),
                                           C.c_int32(oc.shape[0]), _dp(xy), _ip(oc), _ip(op)))
        self.N += cam.shape[0]; self.M += pt.shape[0]; self.K += oc.shape[0]

    def set_state(self, cam_RT=None, pts=None):
        """``b200ba_set_state``: new parameter values, same observation layout."""
        cam = None if cam_RT is None else np.ascontiguousarray(cam_RT, dtype=np.float64).reshape(-1, 12)
        pt = None if pts is None else np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self._check(self.lib.b200ba_set_state(self.h, None if cam is None else _dp(cam), None if pt is None else _dp(pt)))

    def _report(self, with_trace=True):
        rep = Report()
        cap = self.opt.max_iterations + 2
        tr = np.full((cap, TRACE_COLS), np.nan)
        if with_trace:
            rep.trace = _dp(tr)
            rep.trace_cap = cap
        return rep, tr

    def _result(self, rep, tr) -> SolveResult:
        cam = np.empty((self.N, 12)); pts = np.empty((self.M, 3))
        self._check(self.lib.b200ba_get_result(self.h, _dp(cam), _dp(pts)))
        rep2 = Report()
        return SolveResult(rep.ret, rep.status, rep.iterations, tr[: rep.n_trace].copy(), (rep.mean_px[0], rep.mean_px[1]),
                           cam, pts, rep.total_cg_steps, rep.gpu_launches, rep.ms_solve, rep.ms_h2d, rep.ms_d2h,
                           rep.initial_cost, rep.final_cost, {"iterations_total": rep.iterations_total, "accepted_total": rep.accepted_total})

    def get_result(self):
        """``b200ba_get_result``: (cam_RT N x 12, pts M x 3) as new host arrays."""
        cam = np.empty((self.N, 12)); pts = np.empty((self.M, 3))
        self._check(self.lib.b200ba_get_result(self.h, _dp(cam), _dp(pts)))
        return cam, pts

    def solve(self) -> SolveResult:
        rep, tr = self._report()
        self._check(self.lib.b200ba_solve(self.h, C.byref(rep)))
        return self._result(rep, tr)

    def begin(self):
        self._check(self.lib.b200ba_begin(self.h))

    def iterate(self, n: int) -> bool:
        done = C.c_int32(0)
        self._check(self.lib.b200ba_iterate(self.h, C.c_int32(n), C.byref(done)))
        return bool(done.value)

    def restart(self):
        self._check(self.lib.b200ba_restart(self.h))

    def last_iterate_ms(self) -> float:
        ms = C.c_double(0)
        self._check(self.lib.b200ba_last_iterate_ms(self.h, C.byref(ms)))
        return ms.value

    def last_restart_ms(self) -> float:
        ms = C.c_double(0)
        self._check(self.lib.b200ba_last_restart_ms(self.h, C.byref(ms)))
        return ms.value

    def finish(self) -> SolveResult:
        rep, tr = self._report()
        self._check(self.lib.b200ba_finish(self.h, C.byref(rep)))
        return self._result(rep, tr)

    def plan(self) -> dict:
        """``b200ba_get_plan``: the execution plan chosen by set_problem, as a dict (camera_table as a name)."""
        pl = Plan()
        self._check(self.lib.b200ba_get_plan(self.h, C.byref(pl)))
        out = {k: getattr(pl, k) for k, _ in Plan._fields_ if k not in ("reserved", "struct_size")}
        out["camera_table"] = TABLE_NAMES.get(pl.camera_table, str(pl.camera_table))
        return out

    # -- stage-level ----------------------------------------------------------------------------------
    def debug_linearize(self) -> dict:
        N, M, K = self.N, self.M, self.K
        out = {"w": np.zeros(K), "gc": np.zeros((N, 6)), "gp": np.zeros((M, 3)), "U": np.zeros((N, 6, 6)), "V": np.zeros((M, 3, 3))}
        cost = C.c_double(0); lam0 = C.c_double(0)
        self._check(self.lib.b200ba_debug_linearize(self.h, C.byref(cost), _dp(out["w"]), _dp(out["gc"]), _dp(out["gp"]),
                                                    _dp(out["U"]), _dp(out["V"]), C.byref(lam0)))
        out["cost"] = cost.value; out["lambda0"] = lam0.value
        return out

    def debug_schur_matvec(self, lam: float, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        Sx = np.zeros_like(x)
        self._check(self.lib.b200ba_debug_schur_matvec(self.h, C.c_double(lam), _dp(x), _dp(Sx)))
        return Sx.reshape(self.N, 6)

    def time_kernel(self, which: int, reps: int = 20) -> float:
        ms = C.c_double(0)
        self._check(self.lib.b200ba_debug_time_kernel(self.h, C.c_int32(which), C.c_int32(reps), C.byref(ms)))
        return ms.value

    def kernel_names(self) -> list[str]:
        out = []
        i = 0
        while True:
            n = self.lib.b200ba_kernel_name(i)
            if not n:
                return out
            out.append(n.decode()); i += 1
