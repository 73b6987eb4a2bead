/* b200ba_io.h -- C ABI of the on-disk model formats (libb200ba.so, host code), SURVEY.md section 8(f) rank 3.
 *
 * Loads / saves the three text formats the reference reads and writes through PBA's utility header, producing exactly
 * the flat arrays b200ba_set_problem takes (so public BAL files and the tracker's NVM exports run through the same ABI):
 *
 *   BAL / "bundler model" text   LoadBundlerModel / SaveBundlerModel   Thirdparty/pba/pba/pba_util.h:309-390
 *   VisualSFM .nvm (V3, V3_R9T)  LoadNVM / SaveNVM                     Thirdparty/pba/pba/pba_util.h:52-171
 *   Bundler .out (v0.3)          LoadBundlerOut / SaveBundlerOut       Thirdparty/pba/pba/pba_util.h:173-307
 *   dispatch on the file name    LoadModelFile / SaveModelFile         Thirdparty/pba/pba/pba_util.h:392-411,728
 *
 * Conventions on load are the reference's: every camera comes out as [R|T] with x_cam = R X + T, +z forward,
 * pixel = f * x_cam.xy / x_cam.z (BAL and .out store the opposite handedness: rows 2-3 of R and T[1..2] are negated and
 * measurements become (x, -y), pba_util.h:341, DataInterface.h:287-316); NVM quaternions (w, x, y, z) + camera centres
 * become R and T = -R C (DataInterface.h:183-206,276-280); the NVM distortion d is stored like PBA does, d / f^2 with
 * distortion_type = -1 (DataInterface.h:81-86).  The reference keeps everything in float; this loader keeps the file's
 * doubles (its cameras therefore agree with the reference's to float precision, indices exactly).
 * Saving writes the same layouts with 17 significant digits instead of the reference's setprecision(12), so a
 * save -> load round trip is lossless in double.
 *
 * Deviations, on purpose: an observation whose camera or point index is out of range is an error (B200BA_E_INVALID);
 * the reference's BAL loader silently keeps a zero observation in its place (pba_util.h:336-339).  The reference's NVM
 * export in the tracker assigns imx twice and never imy (src/sfm_tracker.cpp:1109-1110); b200io_save writes both.
 *
 * Same conventions as b200ba.h: 0 or a negative B200BA_E_* code, never throws.  Pure host code: no CUDA device needed.
 */
#ifndef B200BA_IO_H
#define B200BA_IO_H

#include <stdint.h>
#include "b200ba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200io_model {
    int32_t  N, M, K;
    double*  cam_RT;            /* N x 12 row-major [R|T]                                                           */
    double*  focal;             /* N                                                                                */
    double*  radial;            /* N   PBA's `radial` field                                                         */
    int32_t* distortion_type;   /* N   0 none, 1 projection distortion (BAL k1 / .out k1), -1 measurement (NVM)     */
    double*  pts;               /* M x 3                                                                            */
    int32_t* pt_rgb;            /* M x 3 (NVM / .out colours; zeros for BAL)                                        */
    double*  obs_xy;            /* K x 2, relative to the principal point                                           */
    int32_t* obs_cam;           /* K                                                                                */
    int32_t* obs_pt;            /* K                                                                                */
    char**   names;             /* N image names (NVM) or NULL                                                      */
    int32_t  reserved[4];
} b200io_model;

/* Allocate a model with zeroed arrays of the given sizes (for callers that fill it and save). */
int b200io_alloc(int32_t N, int32_t M, int32_t K, int32_t with_names, b200io_model** out);
void b200io_free(b200io_model* m);
/* By file name like LoadModelFile: "*.nvm" -> NVM, "*.out" -> Bundler, anything else -> BAL text. */
int b200io_load(const char* path, b200io_model** out);
int b200io_save(const char* path, const b200io_model* m);
const char* b200io_last_error(void);
/* Stable sort of the observations by (point, camera): the order b200ba_set_problem requires and PBA asserts
 * (Thirdparty/pba/pba/SparseBundleCPU.cpp:2730-2736). */
int b200io_sort_by_point(b200io_model* m);
/* b200ba_set_problem takes ONE intrinsic matrix (SequenceSFM has one camera).  Rescales every measurement by
 * f_shared / f_i (exact for a pinhole without distortion) with f_shared = median focal, and writes K5 = {f, 0, 0, f, 0}. */
int b200io_shared_focal(b200io_model* m, double* K5);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_IO_H */
