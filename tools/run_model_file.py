"""Bundle-adjust a model file (BAL / bundler-model text, VisualSFM .nvm, Bundler .out) on the GPU and write the result.

    python tools/run_model_file.py problem-49-7776-pre.txt refined.txt [--pba] [--iterations 50] [--ground]

Load (include/b200ba_io.h) -> observations sorted by point, one shared focal -> b200ba_solve -> cameras / points written
back into the model -> save in the format the output name asks for.  --ground additionally runs the post-BA ground-plane
stage chained on the device (b200post_ground_stage_ba) and reports how many points it keeps; the saved model then holds
the re-axed cameras and the kept, re-axed points with their surviving observations.  Needs a CUDA device (no CPU fallback).
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200 import binding, io as bio, post


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("input"); ap.add_argument("output")
    ap.add_argument("--pba", action="store_true", help="PBA personality (camera 0 fixed, no robustifier, sub-pixel measurements)")
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--ground", action="store_true", help="post-BA ground-plane fit, outlier drop and re-axis")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--force", action="store_true", help="adjust files with radial distortion / per-camera focals anyway (distortion ignored)")
    ap.add_argument("--round-pixels", action="store_true", help="round measurements to whole pixels like the tracker's SSBA call (off for files)")
    a = ap.parse_args()
    m = bio.load(a.input)
    src_xy, src_focal, src_radial = m.obs_xy.copy(), m.focal.copy(), m.radial.copy()
    if np.any(src_radial != 0) or np.ptp(src_focal) > 1e-9 * abs(np.median(src_focal)):
        if not a.force:
            sys.exit(f"{a.input}: cameras carry radial distortion ({int(np.count_nonzero(src_radial))} non-zero) or different focal lengths "
                     f"({src_focal.min():.3f} .. {src_focal.max():.3f}); the BA boundary (src/sfm_ba.h) has ONE pinhole calibration, so the "
                     f"refinement would ignore the distortion and rescale every measurement to the median focal. Re-run with --force to do "
                     f"that anyway (the output keeps the source measurements, focals and distortion; only poses and points change).")
        print("WARNING: radial distortion ignored and measurements rescaled to the median focal for the adjustment (--force)", file=sys.stderr)
    p = m.to_problem(os.path.basename(a.input))
    order = np.lexsort((m.obs_cam, m.obs_pt))          # the stable (point, camera) order b200io_sort_by_point produces
    assert np.array_equal(m.obs_cam[order], p.obs_cam) and np.array_equal(m.obs_pt[order], p.obs_pt)
    src_xy = src_xy[order]
    print(f"{a.input}: {p.N} cameras, {p.M} points, {p.K} observations, shared focal {p.K5[0]:.3f}")
    opt = binding.default_options("pba" if a.pba else "ssba", device=a.device, max_iterations=a.iterations, discard_converged=0,
                                 **({} if (a.pba or a.round_pixels) else {"round_pixels": 0}))
    with binding.Engine(opt) as eng:
        eng.set_problem_from(p)
        r = eng.solve()
        print(f"status {binding.STATUS_NAMES[r.status]} after {r.iterations} iterations, mean reprojection error "
              f"{r.mean_px[0]:.4f} -> {r.mean_px[1]:.4f} px, {r.ms_solve:.2f} ms on the device")
        cams, pts, keep = r.cam_RT, r.pts, None
        if a.ground:
            with post.PostStage(device=a.device) as st:
                fit, pts, cams, axes = st.ground_stage_ba(eng, post.draw_triples(p.M, 100))
            keep = fit.keep != 0
            print(f"ground plane {fit.plane}, kept {fit.n_kept} of {p.M} points, {fit.ms_device:.2f} ms on the device")
    # the source file's own measurements, focals and distortion go back out; only cameras and points changed
    out = bio.Model(cams, src_focal, src_radial, np.zeros(p.N, np.int32), pts, m.pt_rgb if keep is None else m.pt_rgb[keep],
                    src_xy, p.obs_cam, p.obs_pt, m.names)
    if keep is not None:                       # drop the observations of dropped points, renumber the rest
        new_id = np.cumsum(keep) - 1
        ok = keep[p.obs_pt]
        out.obs_xy, out.obs_cam, out.obs_pt = src_xy[ok], p.obs_cam[ok], new_id[p.obs_pt[ok]].astype(np.int32)
    bio.save(a.output, out)
    print(f"wrote {a.output}: {out.N} cameras, {out.M} points, {out.K} observations")


if __name__ == "__main__":
    main()
