"""Seeded synthetic BAL-shaped bundle-adjustment problems.

The reference tree ships no BAL files, so every benchmark/parity problem is produced here from a seed:
cameras in a slab looking down +z at a cloud of points, a random observation graph with the named
camera/point/observation counts, exact projections plus Gaussian pixel noise, and a perturbed initial
state.  Observations are emitted sorted by point and, inside a point, by camera id -- the order
``adjustBundle_ssba`` produces when it walks ``pointcloud[j].img_kp`` (a ``std::map<frame,kp>``),
reference ``src/sfm_ba_ssba.cpp:161-182``, and the order PBA requires
(``Thirdparty/pba/pba/SparseBundleCPU.cpp:2730-2736``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# N cameras / M points / K observations of the BASELINE.json configs (published BAL sizes).
SHAPES = {
    "tiny": (6, 120, 600),
    "small": (12, 800, 4000),
    "ladybug": (49, 7776, 31843),
    "trafalgar": (257, 65132, 225911),
    "venice": (1778, 993923, 5001946),
    "final": (13682, 4456117, 28987644),
}

# Big cases with a committed oracle trace (tests/golden/trace50_<case>_s<seed>.json, tools/make_golden_big.py):
# case -> (shape, seed, LM iterations stored).  "camheavy" has Final's camera count (the execution plan for camera tables that
# do not fit one SM's shared memory) with the point count cut down so that the CPU oracle finishes in seconds.
BIG_CASES = {
    "venice": ("venice", 1, 6),
    "final": ("final", 1, 3),
    "camheavy": ((13682, 250000, 1250000), 1, 4),
    "ladybug": ("ladybug", 1, 8),        # BASELINE configs 2 and 3: the shapes the explicit reduced-system plan runs
    "trafalgar": ("trafalgar", 1, 8),
}


def input_checksum(p: "Problem") -> str:
    """sha256 over the arrays a problem hands to b200ba_set_problem (pins the generator for the committed fixtures)."""
    import hashlib
    h = hashlib.sha256()
    for a in (p.K5, p.cam_RT, p.pts, p.obs_xy, p.obs_cam, p.obs_pt):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@dataclass
class Problem:
    """One BA problem at the drop-in boundary (everything the reference marshals, as flat arrays)."""

    K5: np.ndarray        # fx, skew, cx, fy, cy   (camIntrinsic, src/sfm_ba_ssba.cpp:90-94)
    cam_RT: np.ndarray    # N x 12 row-major [R|T] world->camera (Pmats, src/sfm_ba_ssba.cpp:130-146)
    pts: np.ndarray       # M x 3                  (pointcloud[j].pt)
    obs_xy: np.ndarray    # K x 2 float pixels     (imgpts[frame][kp].pt)
    obs_cam: np.ndarray   # K int32, index into cam_RT
    obs_pt: np.ndarray    # K int32, non-decreasing
    name: str = ""
    seed: int = 0

    @property
    def N(self) -> int:
        return int(self.cam_RT.shape[0])

    @property
    def M(self) -> int:
        return int(self.pts.shape[0])

    @property
    def K(self) -> int:
        return int(self.obs_cam.shape[0])

    def copy(self) -> "Problem":
        return Problem(self.K5.copy(), self.cam_RT.copy(), self.pts.copy(), self.obs_xy.copy(),
                       self.obs_cam.copy(), self.obs_pt.copy(), self.name, self.seed)

    def slice_points(self, j0: int, j1: int) -> "Problem":
        """Sub-problem owning points [j0, j1) and their observations (point-partitioned rank shard)."""
        k0 = int(np.searchsorted(self.obs_pt, j0, side="left"))
        k1 = int(np.searchsorted(self.obs_pt, j1, side="left"))
        return Problem(self.K5.copy(), self.cam_RT.copy(), self.pts[j0:j1].copy(), self.obs_xy[k0:k1].copy(),
                       self.obs_cam[k0:k1].copy(), (self.obs_pt[k0:k1] - j0).astype(np.int32), self.name, self.seed)


def load_problem_npz(path: str, name: str = "") -> Problem:
    """A captured BA problem stored as .npz (K5, cam_RT, pts, obs_xy float32 pixels, obs_cam, obs_pt [, frame_ids]) -
    the demo-sequence fixtures written by tools/make_demo_problem.py (BASELINE.json configs[0])."""
    z = np.load(path)
    return Problem(z["K5"].astype(np.float64), np.ascontiguousarray(z["cam_RT"], dtype=np.float64), np.ascontiguousarray(z["pts"], dtype=np.float64),
                   np.ascontiguousarray(z["obs_xy"].astype(np.float32).astype(np.float64)), z["obs_cam"].astype(np.int32), z["obs_pt"].astype(np.int32),
                   name or "fixture", 0)


def pba_inputs(p: Problem) -> Problem:
    """The numbers the reference's PBA path hands to ParallelBA (src/sfm_ba_pba.cpp:56-130): cameras, points and measurements as
    float32 (CameraT / Point3D / Point2D are float), f = (fx + fy) / 2, principal point subtracted in float."""
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32)
    f = np.float32((p.K5[0] + p.K5[3]) / 2.0)
    xy = np.stack([f32(p.obs_xy[:, 0]) - np.float32(p.K5[2]), f32(p.obs_xy[:, 1]) - np.float32(p.K5[4])], axis=1).astype(np.float64)
    K5 = np.array([float(f), 0.0, 0.0, float(f), 0.0])
    return Problem(K5, f32(p.cam_RT).astype(np.float64), f32(p.pts).astype(np.float64), np.ascontiguousarray(xy), p.obs_cam.copy(), p.obs_pt.copy(), p.name, p.seed)


def rodrigues(w: np.ndarray) -> np.ndarray:
    """Batch Rodrigues: (n,3) rotation vectors -> (n,3,3) rotation matrices."""
    w = np.asarray(w, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(w, axis=1)
    safe = np.where(th > 1e-12, th, 1.0)
    c1 = np.where(th > 1e-12, np.sin(th) / safe, 1.0)
    c2 = np.where(th > 1e-12, (1.0 - np.cos(th)) / (safe * safe), 0.5)
    J = np.zeros((w.shape[0], 3, 3))
    J[:, 0, 1] = -w[:, 2]; J[:, 0, 2] = w[:, 1]
    J[:, 1, 0] = w[:, 2]; J[:, 1, 2] = -w[:, 0]
    J[:, 2, 0] = -w[:, 1]; J[:, 2, 1] = w[:, 0]
    return np.eye(3)[None] + c1[:, None, None] * J + c2[:, None, None] * (J @ J)


def project(K5: np.ndarray, cam_RT: np.ndarray, pts: np.ndarray, obs_cam: np.ndarray, obs_pt: np.ndarray) -> np.ndarray:
    """Pixel projection of every observation, K[R|T] pinhole with skew (numpy, for generation/metrics)."""
    P = cam_RT.reshape(-1, 3, 4)
    X = pts[obs_pt]
    Pc = P[obs_cam]
    XX = np.einsum("kij,kj->ki", Pc[:, :, :3], X) + Pc[:, :, 3]
    u = XX[:, 0] / XX[:, 2]
    v = XX[:, 1] / XX[:, 2]
    fx, skew, cx, fy, cy = K5
    return np.stack([fx * u + skew * v + cx, fy * v + cy], axis=1)


def observation_graph(rng: np.random.Generator, N: int, M: int, K: int) -> tuple[np.ndarray, np.ndarray]:
    """Random bipartite graph: every point seen by >= 2 distinct cameras, K observations in total."""
    if K < 2 * M:
        raise ValueError("need K >= 2*M")
    deg = np.full(M, 2, dtype=np.int64)
    extra = rng.integers(0, M, size=K - 2 * M)
    deg += np.bincount(extra, minlength=M)
    np.minimum(deg, N, out=deg)                 # a point cannot be seen by more cameras than exist
    short = K - int(deg.sum())
    while short > 0:                            # only reachable for tiny N
        room = np.flatnonzero(deg < N)
        add = rng.choice(room, size=min(short, room.size), replace=False)
        deg[add] += 1
        short = K - int(deg.sum())
    obs_pt = np.repeat(np.arange(M, dtype=np.int64), deg)
    obs_cam = rng.integers(0, N, size=obs_pt.size)
    # enforce "cameras without replacement per point": redraw duplicates until none are left
    for _ in range(200):
        key = obs_pt * N + obs_cam
        order = np.argsort(key, kind="stable")
        sk = key[order]
        dup = np.zeros(sk.size, dtype=bool)
        dup[1:] = sk[1:] == sk[:-1]
        if not dup.any():
            break
        idx = order[dup]
        obs_cam[idx] = rng.integers(0, N, size=idx.size)
    else:  # pragma: no cover
        raise RuntimeError("could not de-duplicate observation graph")
    order = np.argsort(obs_pt * N + obs_cam, kind="stable")
    return obs_cam[order].astype(np.int32), obs_pt[order].astype(np.int32)


def make_problem(shape: str | tuple[int, int, int], seed: int = 1, *, f: float = 600.0, fy_ratio: float = 1.0,
                 pp: tuple[float, float] = (0.0, 0.0), noise_px: float = 0.5, pt_sigma: float = 0.03,
                 t_sigma: float = 0.01, rot_sigma: float = 0.0, outlier_frac: float = 0.0) -> Problem:
    """Build the seeded problem described in SURVEY.md section 8(d).

    ``outlier_frac`` > 0 adds gross (20 px) errors to that fraction of measurements so that the Huber
    branch of the reference's ``fillWeights`` (``Geometry/v3d_metricbundle.h:49-57``) is exercised.
    """
    name = shape if isinstance(shape, str) else "custom"
    N, M, K = SHAPES[shape] if isinstance(shape, str) else shape
    rng = np.random.default_rng(seed)
    centres = np.stack([rng.uniform(-3, 3, N), rng.uniform(-3, 3, N), rng.uniform(-1, 1, N)], axis=1)
    R = rodrigues(rng.normal(0.0, 0.08, (N, 3)))
    T = -np.einsum("nij,nj->ni", R, centres)
    pts = np.stack([rng.uniform(-3, 3, M), rng.uniform(-3, 3, M), rng.uniform(8, 14, M)], axis=1)
    obs_cam, obs_pt = observation_graph(rng, N, M, K)
    K5 = np.array([f, 0.0, pp[0], f * fy_ratio, pp[1]], dtype=np.float64)
    cam_true = np.concatenate([R, T[:, :, None]], axis=2).reshape(N, 12)
    xy = project(K5, cam_true, pts, obs_cam, obs_pt)
    xy += rng.normal(0.0, noise_px, xy.shape)
    if outlier_frac > 0:
        bad = rng.random(xy.shape[0]) < outlier_frac
        xy[bad] += rng.normal(0.0, 20.0, (int(bad.sum()), 2))
    # SequenceSFM stores keypoints as cv::Point2f: the boundary sees float32 pixels
    xy = xy.astype(np.float32).astype(np.float64)
    # perturbed initial state
    pts0 = pts + rng.normal(0.0, pt_sigma, pts.shape)
    T0 = T + rng.normal(0.0, t_sigma, T.shape)
    R0 = R
    if rot_sigma > 0:
        R0 = np.einsum("nij,njk->nik", rodrigues(rng.normal(0.0, rot_sigma, (N, 3))), R)
    cam0 = np.concatenate([R0, T0[:, :, None]], axis=2).reshape(N, 12)
    return Problem(K5, np.ascontiguousarray(cam0), np.ascontiguousarray(pts0), np.ascontiguousarray(xy),
                   obs_cam, obs_pt, name, seed)


def mean_reprojection_error_px(p: Problem, round_pixels: bool = True) -> float:
    """The reference's only runtime self-check (``showErrorStatistics_ssba``, src/sfm_ba_ssba.cpp:26-51)."""
    xy = np.rint(p.obs_xy.astype(np.float32)).astype(np.float64) if round_pixels else p.obs_xy
    e = project(p.K5, p.cam_RT, p.pts, p.obs_cam, p.obs_pt) - xy
    # the reference scales the normalised residual by f0 = fx in both axes
    return float(np.mean(np.linalg.norm(e, axis=1)))


def partition_points(obs_pt: np.ndarray, M: int, world: int) -> list[tuple[int, int]]:
    """Contiguous point ranges balanced by observation count."""
    K = obs_pt.shape[0]
    starts = np.searchsorted(obs_pt, np.arange(M + 1), side="left")  # CSR offsets by point
    cuts = [0]
    for r in range(1, world):
        target = (K * r) // world
        j = int(np.searchsorted(starts, target, side="left"))
        cuts.append(min(max(j, cuts[-1]), M))
    cuts.append(M)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


# ---- post-BA ground-plane stage -------------------------------------------------------
def make_ground_scene(seed: int, M: int, *, extent: float = 30.0, noise: float = 0.3, outlier_frac: float = 0.1,
                      tilt: float = 0.4, n_cams: int = 8, on_plane_frac: float = 0.0):
    """A map as the tracker hands it to ``PointsPlaneFit_RANSAC`` (``src/sfm_tracker.cpp:884``): M points scattered about a
    tilted, shifted ground plane with Gaussian height noise, a fraction of outliers well above it (buildings, trees), and
    cameras flying over it looking down.  ``on_plane_frac`` puts that fraction of points EXACTLY on a coordinate-aligned
    plane before the tilt is applied with tilt = 0 (ties in the distance median).  Returns (pts M x 3, cam_RT n_cams x 12).
    """
    rng = np.random.default_rng(seed)
    pts = np.c_[rng.uniform(-extent, extent, M), rng.uniform(-extent, extent, M), noise * rng.standard_normal(M)]
    n_out = int(outlier_frac * M)
    pts[:n_out, 2] += rng.uniform(2.0, 15.0, n_out)
    n_on = int(on_plane_frac * M)
    if n_on:
        pts[n_out:n_out + n_on, 2] = 0.0
    cams = np.zeros((n_cams, 3, 4))
    centres = np.c_[rng.uniform(-extent, extent, n_cams), rng.uniform(-extent, extent, n_cams), rng.uniform(40, 60, n_cams)]
    for i in range(n_cams):
        yaw = rng.uniform(-np.pi, np.pi)
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
        Rc = np.diag([1.0, -1.0, -1.0]) @ Rz            # looking down -z of the world
        cams[i, :, :3] = Rc
        cams[i, :, 3] = -Rc @ centres[i]
    if tilt > 0:
        w = rng.normal(0, tilt, 3)
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = rng.normal(0, 5, 3)
        pts = pts @ R.T + t
        for i in range(n_cams):                          # x_cam = Rc (R^t (X' - t)) + tc
            Rc, tc = cams[i, :, :3].copy(), cams[i, :, 3].copy()
            cams[i, :, :3] = Rc @ R.T
            cams[i, :, 3] = tc - Rc @ R.T @ t
    pts = np.ascontiguousarray(pts[rng.permutation(M)])
    return pts, np.ascontiguousarray(cams.reshape(n_cams, 12))


# ---- two-view triangulation -----------------------------------------------------------
def make_two_view(seed: int, n: int, *, baseline: float = 1.0, noise_px: float = 0.5, dist=None, f: float = 800.0,
                  pp=(640.0, 360.0)):
    """Two cameras `baseline` apart looking at n points 5-15 units ahead, pixel measurements with Gaussian noise (stored
    as float32 like cv::KeyPoint::pt) and, optionally, lens distortion applied to the ideal pixels.
    Returns (pt1, pt2, K, P, P1, X_true)."""
    rng = np.random.default_rng(seed)
    K = np.array([[f, 0, pp[0]], [0, f * 1.01, pp[1]], [0, 0, 1.0]])
    X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(5, 15, n)]
    P = np.hstack([np.eye(3), np.zeros((3, 1))])
    w = rng.normal(0, 0.05, 3); th = np.linalg.norm(w); k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    t = np.array([-baseline, 0.05 * baseline, 0.1 * baseline])
    P1 = np.hstack([R, t[:, None]])

    def project(Pm):
        xc = X @ Pm[:, :3].T + Pm[:, 3]
        x, y = xc[:, 0] / xc[:, 2], xc[:, 1] / xc[:, 2]
        if dist is not None:
            kk = np.zeros(8); kk[:len(dist)] = dist
            r2 = x * x + y * y
            rad = (1 + ((kk[4] * r2 + kk[1]) * r2 + kk[0]) * r2) / (1 + ((kk[7] * r2 + kk[6]) * r2 + kk[5]) * r2)
            xd = x * rad + 2 * kk[2] * x * y + kk[3] * (r2 + 2 * x * x)
            yd = y * rad + kk[2] * (r2 + 2 * y * y) + 2 * kk[3] * x * y
            x, y = xd, yd
        px = np.c_[x * K[0, 0] + K[0, 2], y * K[1, 1] + K[1, 2]] + rng.normal(0, noise_px, (n, 2))
        return px.astype(np.float32)

    return project(P), project(P1), K, P, P1, X
