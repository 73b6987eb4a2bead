"""BASELINE.json configs[0]: the repo demo sequence `data/img_sfm_1s` through the reference's live BA path (SSBA on CPU).

The SequenceSFM application itself cannot be built in this image (Qt4/5, OpenGL, OpenCV 2.4 C++ libs with the nonfree SURF
module are absent), so the BA inputs are produced by a cv2 restatement of the tracker FRONT END whose only
job is to emit what `SfM_Tracker::bundleAdjustFrames` (src/sfm_tracker.cpp:858-866) hands to `adjustBundle_ssba` /
`adjustBundle_pba` -- (pointcloud, imgpts, Pmats, K) -- on real images:

  keypoints          cv2.SIFT  **instead of SURF (minHessian 200)**: cv2.xfeatures2d is absent     src/sfm_frame.cpp:100-109
  matching           FLANN 1-NN, displacement filter, fundamental-matrix RANSAC                    src/sfm_frame.cpp:122-229
  baseline           two-view reconstruction from the essential matrix                             src/sfm_tracker.cpp:441-557,246-320
  new keyframe       every >= 5 frames: 2D-3D correspondences, solvePnPRansac(EPNP), two-view
                     triangulation against the best previous keyframe, 80th-percentile gating     src/sfm_tracker.cpp:117-195,559-857
  calibration        data/img_sfm_1s.ini:35-48 (cam_138s); BA sees RAW keypoint pixels and K only, like the reference

Between keyframes the state is advanced with the REFERENCE's own BA (oracle/_ref/libssba_ref.so, unmodified SSBA) so that the
captured problems are the inputs the reference would see next.  The substitution (SIFT for SURF, cv2 4.13 for OpenCV 2.4.9)
changes WHICH keypoints exist, not what the BA does with them: parity is judged at the BA boundary - same captured inputs ->
reference SSBA vs the CUDA engine (tests/test_gpu_parity.py::test_demo_sequence_matches_reference_ssba).

    python tools/make_demo_problem.py            # -> tests/golden/demo_kf*.npz + demo_kf*.json  (build container only)
"""
import json, os, sys, time
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200.synth import Problem, input_checksum
from oracle import refs

SEQ = "/root/reference/data/img_sfm_1s"
K = np.array([[309.6069470329391, 0.0, 353.04952639076436], [0.0, 309.7358151264549, 344.4001844409269], [0.0, 0.0, 1.0]])
DIST = np.array([0.092048412922401, -0.042851770292329, -0.005008840639079, 0.000891060323127, 0.0])
KF_MIN_INTERVAL, KF_SEARCH, DISP_THRESHOLD, BASE_MAX = 5, 3, 30.0, 10
CAPTURE_AT = (5, 12, 30)          # keyframe counts whose BA inputs become fixtures (+ the BA after the sequence's last keyframe)


class Frame:
    def __init__(self, idx):
        self.idx = idx
        img = cv2.imread(os.path.join(SEQ, f"{idx:04d}.jpg"), cv2.IMREAD_GRAYSCALE)
        kp, self.desc = cv2.SIFT_create(nfeatures=3000).detectAndCompute(img, None)
        self.kp = np.array([k.pt for k in kp], dtype=np.float32)          # KeyPoint::pt is a Point2f
        self.P = np.hstack([np.eye(3), np.zeros((3, 1))])


_flann = cv2.FlannBasedMatcher(dict(algorithm=1, trees=4), dict(checks=64))


def match(fp, f):
    """matches (query = previous frame fp, train = new frame f) after the displacement filter and F-matrix RANSAC; None when
    the mean displacement is below tracker_dispThreshold (src/sfm_frame.cpp:178-180)."""
    m = _flann.match(fp.desc, f.desc)
    q = np.array([x.queryIdx for x in m]); t = np.array([x.trainIdx for x in m])
    d = np.linalg.norm(fp.kp[q] - f.kp[t], axis=1)
    med = np.median(d); sig = 1.4826 * np.median(np.abs(d - med)) + 1e-9
    keep = np.abs(d - med) < 3.0 * sig                                     # stands in for FilterData_NormDist (histogram mode +- 0.7 sigma)
    if d[keep].mean() < DISP_THRESHOLD:
        return None
    q, t = q[keep], t[keep]
    # one-to-one in the new frame (the reference resolves this implicitly through its kp -> cloud-point maps)
    _, first = np.unique(t, return_index=True)
    q, t = q[np.sort(first)], t[np.sort(first)]
    F, mask = cv2.findFundamentalMat(fp.kp[q], f.kp[t], cv2.FM_RANSAC, 0.006 * max(fp.kp[q].max(), f.kp[t].max()), 0.99)
    if F is None:
        return None
    ok = mask.ravel() == 1
    return q[ok], t[ok]


def undist(pts):
    return cv2.undistortPoints(pts.reshape(-1, 1, 2).astype(np.float64), K, DIST).reshape(-1, 2)


def triangulate(P0, P1, x0, x1):
    X = cv2.triangulatePoints(P0, P1, x0.T, x1.T)
    return (X[:3] / X[3]).T


def reproj_err(P, X, px):
    rv, _ = cv2.Rodrigues(P[:, :3])
    pr, _ = cv2.projectPoints(X.reshape(-1, 1, 3), rv, P[:, 3], K, DIST)
    return np.linalg.norm(pr.reshape(-1, 2) - px, axis=1)


class Tracker:
    def __init__(self):
        self.frames = {}
        self.kf = []                       # keyframe ids, ascending
        self.pts = []                      # cloud: list of [xyz]
        self.img_kp = []                   # per cloud point: {frame id: keypoint index}
        self.captures = {}

    def frame(self, i):
        if i not in self.frames:
            self.frames[i] = Frame(i)
        return self.frames[i]

    # ---- baseline (src/sfm_tracker.cpp:246-320,441-557) ----
    def baseline(self):
        f2 = self.frame(BASE_MAX - 1)
        best = None
        for i1 in (4, 2, 0):
            f1 = self.frame(i1)
            mt = match(f1, f2)
            if mt is None:
                continue
            q, t = mt
            x1, x2 = undist(f1.kp[q]), undist(f2.kp[t])
            E, mask = cv2.findEssentialMat(x1, x2, np.eye(3), cv2.RANSAC, 0.999, 1e-3)
            n, R, tv, mask = cv2.recoverPose(E, x1, x2, np.eye(3), mask=mask)
            ok = mask.ravel() > 0
            if best is None or ok.sum() > best[0]:
                best = (ok.sum(), f1, q[ok], t[ok], R, tv, x1[ok], x2[ok])
        _, f1, q, t, R, tv, x1, x2 = best
        f1.P = np.hstack([np.eye(3), np.zeros((3, 1))]); f2.P = np.hstack([R, tv])
        X = triangulate(f1.P, f2.P, x1, x2)
        z = X[:, 2]; lo, hi = np.quantile(z, [0.125, 0.875])                # CloudPoint_FilterZ: tracker_baselineZFilter_r1/r2
        zm, zs = z[(z >= lo) & (z <= hi)].mean(), z[(z >= lo) & (z <= hi)].std()
        good = (z > 0) & (np.abs(z - zm) < 8 * zs)
        for k in np.flatnonzero(good):
            self.pts.append(X[k]); self.img_kp.append({f1.idx: int(q[k]), f2.idx: int(t[k])})
        self.kf = [f1.idx, f2.idx]
        print(f"baseline {f1.idx}-{f2.idx}: {good.sum()} points", flush=True)

    # ---- BA boundary: exactly what adjustBundle_ssba marshals (src/sfm_ba_ssba.cpp:88-182) ----
    def problem(self):
        ids = sorted(self.kf); index = {f: i for i, f in enumerate(ids)}
        cam = np.array([self.frames[f].P.reshape(12) for f in ids])
        xy, oc, op = [], [], []
        for j, m in enumerate(self.img_kp):
            for f in sorted(m):
                xy.append(self.frames[f].kp[m[f]]); oc.append(index[f]); op.append(j)
        K5 = np.array([K[0, 0], K[0, 1], K[0, 2], K[1, 1], K[1, 2]])
        return Problem(K5, cam, np.array(self.pts, dtype=np.float64), np.array(xy, dtype=np.float32).astype(np.float64),
                       np.array(oc, dtype=np.int32), np.array(op, dtype=np.int32), "demo", 0), ids

    def bundle_adjust(self):
        p, ids = self.problem()
        nkf = len(ids)
        if nkf in CAPTURE_AT and nkf not in self.captures:
            self.captures[nkf] = (p.copy(), list(ids))
        self.last_capture = (nkf, p.copy(), list(ids))
        r = refs.ssba_adjust(p, trace=False)                                # the reference's own BA advances the state
        print(f"  BA kf={nkf} N={p.N} M={p.M} K={p.K}: status {r.status} its {r.iterations} px {r.mean_px[0]:.3f} -> {r.mean_px[1]:.3f}", flush=True)
        if r.ret == 0:
            for i, f in enumerate(ids):
                self.frames[f].P = r.cam_RT[i].reshape(3, 4).copy()
            self.pts = [x for x in r.pts]
        # gross-outlier drop standing in for the post-BA plane fit (src/sfm_tracker.cpp:876-908): points whose worst
        # reprojection error exceeds 8 px leave the cloud
        P = {f: self.frames[f].P for f in ids}
        keep = []
        for j, m in enumerate(self.img_kp):
            X = np.asarray(self.pts[j])
            worst = max(float(reproj_err(P[f], X[None], self.frames[f].kp[m[f]][None])[0]) for f in m)
            keep.append(worst < 8.0)
        self.pts = [x for x, k in zip(self.pts, keep) if k]
        self.img_kp = [m for m, k in zip(self.img_kp, keep) if k]

    # ---- new keyframe (src/sfm_tracker.cpp:117-195) ----
    def try_keyframe(self, idx):
        f = self.frame(idx)
        best = None
        for k in self.kf[-KF_SEARCH:][::-1]:
            fk = self.frames[k]
            mt = match(fk, f)
            if mt is None:
                return False
            q, t = mt
            kp2pt = {m[k]: j for j, m in enumerate(self.img_kp) if k in m}    # get2D3DCorrespondings
            sel = [(kp2pt[a], b) for a, b in zip(q.tolist(), t.tolist()) if a in kp2pt]
            if best is None or len(sel) > len(best[0]):
                best = (sel, fk, q, t)
        sel, fk, q, t = best
        if len(sel) <= 7:
            return False
        X3 = np.array([self.pts[j] for j, _ in sel], dtype=np.float32)
        x2 = np.array([f.kp[b] for _, b in sel], dtype=np.float32)
        Pl = self.frames[self.kf[-1]].P
        rv0, _ = cv2.Rodrigues(Pl[:, :3])
        ok, rv, tv, inl = cv2.solvePnPRansac(X3, x2, K, DIST, rvec=rv0.copy(), tvec=Pl[:, 3:4].copy(), useExtrinsicGuess=True,
                                              iterationsCount=1000, reprojectionError=float(0.006 * x2.max()), confidence=0.99, flags=cv2.SOLVEPNP_EPNP)
        if not ok or inl is None or len(inl) < len(sel) / 5.0:
            return False
        R, _ = cv2.Rodrigues(rv)
        f.P = np.hstack([R, tv.reshape(3, 1)])
        # triangulateTwoFrames (src/sfm_tracker.cpp:693-857)
        X = triangulate(fk.P, f.P, undist(fk.kp[q]), undist(f.kp[t]))
        front = ((X @ fk.P[:, :3].T + fk.P[:, 3])[:, 2] > 0) & ((X @ f.P[:, :3].T + f.P[:, 3])[:, 2] > 0)
        if front.sum() < 10:
            return False
        err = reproj_err(f.P, X, f.kp[t])
        cut = np.sort(err)[4 * len(err) // 5] * 2.4
        good = front & (err < cut)
        kp2pt = {m[fk.idx]: j for j, m in enumerate(self.img_kp) if fk.idx in m}
        added = 0
        for k in np.flatnonzero(good):
            a, b = int(q[k]), int(t[k])
            if a in kp2pt:
                self.img_kp[kp2pt[a]][f.idx] = b
            else:
                self.pts.append(X[k]); self.img_kp.append({fk.idx: a, f.idx: b}); added += 1
        self.kf.append(f.idx)
        print(f"keyframe {f.idx} (#{len(self.kf)}): PnP inliers {len(inl)}/{len(sel)}, +{added} points, cloud {len(self.pts)}", flush=True)
        return True


def main():
    out = os.path.join(ROOT, "tests", "golden")
    tr = Tracker()
    tr.baseline()
    tr.bundle_adjust()
    last = 10 ** 6                                                          # to the end of the sequence
    n_img = len([x for x in os.listdir(SEQ) if x.endswith(".jpg")])
    idx = BASE_MAX
    while len(tr.kf) < last and idx < n_img:
        if idx - tr.kf[-1] >= KF_MIN_INTERVAL and tr.try_keyframe(idx):
            tr.bundle_adjust()
        idx += 1
    nkf, p, ids = tr.last_capture
    tr.captures.setdefault(nkf, (p, ids))
    for nkf, (p, ids) in sorted(tr.captures.items()):
        t0 = time.time()
        r = refs.ssba_adjust(p)                                             # the golden: the reference's run on the captured inputs
        name = f"demo_kf{nkf:02d}"
        np.savez_compressed(os.path.join(out, name + ".npz"), K5=p.K5, cam_RT=p.cam_RT, pts=p.pts, obs_xy=p.obs_xy.astype(np.float32),
                            obs_cam=p.obs_cam.astype(np.int16), obs_pt=p.obs_pt, frame_ids=np.array(ids, dtype=np.int32))
        fx = {"engine": "SSBA-3.0 via oracle/ssba_ref_harness.cpp (unmodified /root/reference sources)",
              "source": "data/img_sfm_1s through tools/make_demo_problem.py (cv2 SIFT front end standing in for the tracker's SURF; see its docstring)",
              "keyframes": nkf, "frame_ids": ids, "N": p.N, "M": p.M, "K": p.K, "input_sha256": input_checksum(p),
              "ret": r.ret, "status": r.status, "iterations": r.iterations, "mean_px": list(r.mean_px), "seconds_minimize": r.seconds,
              "trace_cols": ["iter", "err", "lambda", "new_err", "rho", "accepted"],
              "trace": [[None if np.isnan(v) else float(v) for v in row] for row in r.trace],
              "cam_RT_sum": float(np.sum(r.cam_RT)), "pts_sum": float(np.sum(r.pts)),
              "cam_RT_head": r.cam_RT[:3].tolist(), "pts_head": r.pts[:5].tolist()}
        json.dump(fx, open(os.path.join(out, name + ".json"), "w"), indent=1)
        print(name, p.N, p.M, p.K, "SSBA status", r.status, "its", r.iterations, "px", r.mean_px, f"{time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
