"""Generate tests/golden/post_*.json from the reference's OWN post-BA functions (oracle/_ref/libpost_ref.so, built by
oracle/Makefile from /root/reference/src/sfm_utils.cpp).  Run in the build container only; the fixtures are committed.

    python tools/make_golden_post.py
Each fixture records the generator arguments (the scene is re-created from the seed by
sequencesfm_b200.synth.make_ground_scene and pinned by an input checksum), the rand() values the reference consumed, and
the reference's outputs: plane, plane_se3, cpFilter (bit-packed, hex), the ground axes for camera 0, and samples plus a
checksum of the re-axed points / cameras.  Doubles are stored as C99 hex strings: the comparisons are bit-exact.
"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200.synth import make_ground_scene
from oracle import post

CASES = [
    ("s1_500", 1, 500, {}), ("s2_5000", 2, 5000, {}), ("s3_20001", 3, 20001, {"outlier_frac": 0.3}),
    ("s4_37", 4, 37, {}), ("s5_flat_3000", 5, 3000, {"tilt": 0.0, "on_plane_frac": 0.7}),
    ("s6_k3_4096", 6, 4096, {"noise": 1.0}), ("s7_10", 7, 10, {"outlier_frac": 0.0}),
]
N_RANSAC = {"s6_k3_4096": 37}
K_THR = {"s6_k3_4096": 3.0}


def hexs(a):
    return [float(v).hex() for v in np.asarray(a, np.float64).ravel()]


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, seed, M, kw in CASES:
        pts, cams = make_ground_scene(seed, M, **kw)
        n_ransac, k_thr = N_RANSAC.get(name, 100), K_THR.get(name, 2.0)
        stream = post.rand_stream(seed, 4 * n_ransac + 64)
        fit, used = post.ref_plane_fit(pts, stream, n_ransac, k_thr)
        axes = post.ref_ground_axes(cams[0], fit.plane, fit.plane_se3)
        kept = pts[fit.keep != 0]
        p2, c2 = post.ref_transform(axes, kept, cams)
        fx = {"engine": "PointsPlaneFit_RANSAC / CalcCameraGround_axes / TransCloudPoint / TransCamera of /root/reference/src/sfm_utils.cpp "
                        "via oracle/post_ref_harness.cpp",
              "seed": seed, "M": M, "kwargs": kw, "n_ransac": n_ransac, "k_threshold": k_thr, "input_sha256": sha(pts, cams),
              "rand_stream_seed": seed, "rand_stream_len": int(len(stream)), "rand_used": int(used),
              "plane": hexs(fit.plane), "plane_se3": hexs(fit.plane_se3), "n_kept": int(fit.keep.sum()),
              "keep_packed_hex": np.packbits(fit.keep.astype(np.uint8)).tobytes().hex(),
              "axes_se3": hexs(axes), "points_out_sha256": sha(p2), "cameras_out_sha256": sha(c2),
              "points_out_head": hexs(p2[:4]), "cameras_out": hexs(c2)}
        with open(os.path.join(out_dir, f"post_{name}.json"), "w") as f:
            json.dump(fx, f, indent=1)
        print(name, "kept", fx["n_kept"], "/", M, "rand used", used, flush=True)


if __name__ == "__main__":
    main()
