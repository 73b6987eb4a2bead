"""Host-side mirror of SequenceSFM's BA entry points on top of the C ABI (Python face).

The reference declares, in ``src/sfm_ba.h:29-43``::

    int adjustBundle_ssba(vector<CloudPoint>& pointcloud, Mat& camIntrinsic,
                          map<int, vector<KeyPoint>>& imgpts, map<int, Matx34d>& Pmats);
    int adjustBundle_pba (vector<CloudPoint>& pointcloud, Mat& camIntrinsic, Mat& camDistortion,
                          map<int, vector<KeyPoint>>& imgpts, map<int, Matx34d>& Pmats);

``adjustBundle_ssba`` / ``adjustBundle_pba`` below take the same containers expressed in Python (lists /
dicts / numpy), marshal them exactly like ``src/sfm_ba_ssba.cpp:88-182`` and ``src/sfm_ba_pba.cpp:56-130``,
call the CUDA engine through ``libb200ba.so`` and write the results back in place with the reference's
error convention (0 = results written, -1 = discarded).  The C++ twin is ``adapter/sfm_ba_b200.cpp``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np

from . import binding


@dataclass
class CloudPoint:
    """``struct CloudPoint`` of the reference (``src/sfm_data.h:33-43``)."""

    pt: np.ndarray                                  # cv::Point3d
    img_kp: Dict[int, int] = field(default_factory=dict)   # std::map<int,int>: frame id -> keypoint index
    rgb: tuple = (0, 0, 0)
    match_id: int = 0
    reprojection_error: float = 0.0


def marshal(pointcloud: Sequence[CloudPoint], camIntrinsic: np.ndarray,
            imgpts: Dict[int, Sequence[Sequence[float]]], Pmats: Dict[int, np.ndarray]):
    """Tracker containers -> flat arrays, in the reference's order.

    Cameras in ``std::map`` key order (``src/sfm_ba_ssba.cpp:130``), points in vector order, observations by
    point then by frame id (``:161-182``: outer loop over points, inner over the ``std::map<frame,kp>``).
    """
    cam_ids = sorted(Pmats.keys())
    cam_index = {cid: i for i, cid in enumerate(cam_ids)}
    N, M = len(cam_ids), len(pointcloud)
    Kmat = np.asarray(camIntrinsic, dtype=np.float64)
    K5 = np.array([Kmat[0, 0], Kmat[0, 1], Kmat[0, 2], Kmat[1, 1], Kmat[1, 2]], dtype=np.float64)
    cam_RT = np.empty((N, 12))
    for cid, i in cam_index.items():
        cam_RT[i] = np.asarray(Pmats[cid], dtype=np.float64).reshape(12)
    pts = np.empty((M, 3))
    obs_xy, obs_cam, obs_pt = [], [], []
    for j, cp in enumerate(pointcloud):
        pts[j] = np.asarray(cp.pt, dtype=np.float64)
        for frame in sorted(cp.img_kp.keys()):
            kp = cp.img_kp[frame]
            xy = imgpts[frame][kp]
            # KeyPoint::pt is a cv::Point2f
            obs_xy.append((np.float32(xy[0]), np.float32(xy[1])))
            obs_cam.append(cam_index[frame])
            obs_pt.append(j)
    return (K5, cam_RT, pts, np.asarray(obs_xy, dtype=np.float64).reshape(-1, 2),
            np.asarray(obs_cam, dtype=np.int32), np.asarray(obs_pt, dtype=np.int32), cam_ids)


def _run(preset: str, pointcloud, camIntrinsic, imgpts, Pmats, options: dict | None, report: dict | None) -> int:
    K5, cam_RT, pts, xy, oc, op, cam_ids = marshal(pointcloud, camIntrinsic, imgpts, Pmats)
    print(f"\nadjustBundle_{preset.split('_')[0]}: \n  N (cams) = {len(cam_ids)}, M (points) = {len(pointcloud)}, K (measurements) = {len(oc)}")
    opt = binding.default_options(preset, **(options or {}))
    with binding.Engine(opt) as eng:
        eng.set_problem(K5, cam_RT, pts, xy, oc, op)
        res = eng.solve()
    print(f"mean reprojection error (in pixels): {res.mean_px[0]}")
    print(f"optimizer status = {res.status}")
    print(f"mean reprojection error (in pixels): {res.mean_px[1]}")
    if report is not None:
        report["result"] = res
    if res.ret != 0:
        return -1                                           # src/sfm_ba_ssba.cpp:285 (results discarded)
    for j, cp in enumerate(pointcloud):                     # src/sfm_ba_ssba.cpp:250-257
        cp.pt = res.pts[j].copy()
    for i, cid in enumerate(cam_ids):                       # src/sfm_ba_ssba.cpp:260-271
        Pmats[cid] = res.cam_RT[i].reshape(3, 4).copy()
    return 0


def adjustBundle_ssba(pointcloud: List[CloudPoint], camIntrinsic: np.ndarray,
                      imgpts: Dict[int, Sequence[Sequence[float]]], Pmats: Dict[int, np.ndarray],
                      *, options: dict | None = None, report: dict | None = None) -> int:
    """Drop-in for ``adjustBundle_ssba`` (``src/sfm_ba_ssba.cpp:66-286``): returns 0 and updates ``pointcloud[j].pt``
    and ``Pmats[id]`` in place unless the optimizer ended CONVERGED, in which case it returns -1 and leaves the
    inputs untouched (the reference's ``good_adjustment = (status != 2)``, ``:221``)."""
    return _run("ssba", pointcloud, camIntrinsic, imgpts, Pmats, options, report)


# Build-time switch of the reference (``#ifdef USE_PBA``, ``src/sfm_ba_pba.cpp:14,49``).  The shipped build leaves it
# undefined, so ``adjustBundle_pba`` is a forwarder to ``adjustBundle_ssba``; set it to True (or pass ``use_pba=True``)
# for the behaviour of a ``-DUSE_PBA`` build.
USE_PBA = False


def adjustBundle_pba(pointcloud: List[CloudPoint], camIntrinsic: np.ndarray, camDistortion: np.ndarray,
                     imgpts: Dict[int, Sequence[Sequence[float]]], Pmats: Dict[int, np.ndarray],
                     *, options: dict | None = None, report: dict | None = None, use_pba: bool | None = None) -> int:
    """Drop-in for ``adjustBundle_pba`` (``src/sfm_ba_pba.cpp:33-174``).

    Shipped build (``USE_PBA`` undefined, the default here): prints its banner and **forwards to adjustBundle_ssba,
    returning its value** (``src/sfm_ba_pba.cpp:167-171``) - integer-rounded pixels, Huber weights, all cameras free,
    -1 when the optimizer ended CONVERGED.  This is what the tracker calls for every BA with >= 5 keyframes
    (``src/sfm_tracker.cpp:863-866``).

    With ``USE_PBA`` the reference pins camera 0, keeps sub-pixel float measurements, uses f = (fx+fy)/2 with the
    principal point subtracted and no robustifier, and always returns 0.  ``camDistortion`` is read by the
    reference only to fill a field that ``SetFixedIntrinsics(true)`` + no ``EnableRadialDistortion`` leaves unused
    (``src/sfm_ba_pba.cpp:64-65,87,136``); it is accepted and ignored here as well.
    """
    if not (USE_PBA if use_pba is None else use_pba):
        K = sum(len(cp.img_kp) for cp in pointcloud)        # Count2DMeasurements_pba, src/sfm_ba_pba.cpp:21-30
        print(f"\nadjustBundle_pba: \n  N (cams) = {len(Pmats)}, M (points) = {len(pointcloud)}, K (measurements) = {K}")
        return adjustBundle_ssba(pointcloud, camIntrinsic, imgpts, Pmats, options=options, report=report)
    Kmat = np.array(camIntrinsic, dtype=np.float64)
    f = (Kmat[0, 0] + Kmat[1, 1]) / 2.0                      # src/sfm_ba_pba.cpp:60
    Kp = np.array([[f, 0.0, Kmat[0, 2]], [0.0, f, Kmat[1, 2]], [0.0, 0.0, 1.0]])
    _run("pba_lm", pointcloud, Kp, imgpts, Pmats, options, report)   # PBA personality AND PBA's own LM / PCG / normalisation
    return 0                                                 # src/sfm_ba_pba.cpp:173
