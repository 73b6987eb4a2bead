/* b200ba_post.h -- C ABI of the post-BA ground-plane stage (libb200ba.so), SURVEY.md section 8(f) rank 1.
 *
 * What SfM_Tracker::bundleAdjustFrames runs right after every bundle adjustment (src/sfm_tracker.cpp:876-908) on ALL
 * points of the map.  Reference interface replaced (paths relative to the reference tree):
 *
 *   int PointsPlaneFit_RANSAC(vector<CloudPoint>&, Vec4d& plane, Matx34d& plane_se3, vector<int>& cpFilter,
 *                             int nRansacs = 100, double kThreshold = 2.0)        src/sfm_utils.h:87-91, .cpp:807-977
 *   int CalcCameraGround_axes(Matx34d& cam, Vec4d& p, Matx34d& p_se3, Matx34d& axes_se3)   src/sfm_utils.cpp:1098-1146
 *   int TransCloudPoint(Matx34d& t_se3, vector<CloudPoint>&)                       src/sfm_utils.cpp:1148-1179
 *   int TransCamera(Matx34d& t_se3, map<int, Matx34d>&)                            src/sfm_utils.cpp:1215-1259
 *
 * adapter/sfm_post_b200.cpp holds replacement bodies with exactly these C++ signatures on top of this ABI.
 *
 * Division of work: everything that touches all M points runs in CUDA kernels (distance of every point to every
 * hypothesis plane, the exact median of each hypothesis' M distances by radix selection, the truncated score, the
 * outlier mask, the inlier moments, the re-axis and the order-preserving compaction).  What is O(n_ransac) or O(1)
 * scalars - drawing the triples with rand(), turning 3 points into a plane, the 3x3 SVD of the scatter matrix, the
 * ground axes - stays in host C++ like in the reference.  There is NO CPU fallback for the per-point work: without a
 * CUDA device b200post_create fails with B200BA_E_CUDA.
 *
 * Parity with the reference on the same points and the same rand() values: best hypothesis, sigma^2 and the mask
 * cpFilter are bit-exact (the distances are evaluated in the reference's association order without FMA contraction and
 * the median is an exact order statistic); plane / plane_se3 agree to the single precision the reference computes them
 * in (it runs Eigen's JacobiSVD on a float copy of the scatter matrix; our inlier sums are tree-ordered, not sequential).
 * The re-axis of points and cameras is bit-exact.
 *
 * Same conventions as b200ba.h: 0 or a negative B200BA_E_* code, never throws, caller-owned host buffers.
 */
#ifndef B200BA_POST_H
#define B200BA_POST_H

#include <stdint.h>
#include "b200ba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200post_handle b200post_handle;

typedef struct b200post_plane {
    double  plane[4];           /* [n, -n.mean_of_inliers]   (Vec4d plane)                                    */
    double  plane_se3[12];      /* 3x4 row-major [ax; ay; az | -R mean]   (Matx34d plane_se3)                 */
    double  sigma2;             /* dSigmaSquared_best: inlier threshold on |distance|                         */
    double  score;              /* dBestDistSquared: truncated sum of the winning hypothesis                  */
    double  inlier_mean[3];
    double  inlier_scatter[9];  /* sum (p - mean)(p - mean)^t over the inliers                                */
    int32_t best_hypothesis;    /* row of `triples` that won                                                  */
    int32_t n_inliers;          /* |distance| < sigma2                                                        */
    int32_t n_kept;             /* |distance| < k_threshold * sigma2  (= number of ones in cpFilter)          */
    int32_t n_degenerate;       /* hypotheses skipped by the |n.n| < 1e-9 test (src/sfm_utils.cpp:853)        */
    int32_t select_fallback;    /* 1 when a hypothesis overflowed the candidate buffer and the median was     */
                                /* finished by the remaining radix passes (same result, more sweeps)          */
    int32_t gpu_launches;       /* kernels of this library launched by the call                               */
    double  ms_device;          /* CUDA-event time of the device work of the call (copies included)           */
    double  ms_kernels;         /* the kernels alone                                                          */
    int32_t reserved[4];
} b200post_plane;

/* device: CUDA ordinal, -1 = current.  select_cap: capacity of the per-hypothesis candidate buffer used to finish the
 * median after two radix passes; 0 = default (4096); < 0 = never gather, run all six radix passes. */
int b200post_create(int32_t device, int32_t select_cap, b200post_handle** out);
void b200post_destroy(b200post_handle* h);
const char* b200post_last_error(const b200post_handle* h);

/* The reference's hypothesis draw (src/sfm_utils.cpp:837-845), consuming the process-wide rand() stream exactly like it:
 * nA = rand() % M; nB, nC redrawn until distinct.  triples: n_ransac x 3. */
int b200post_draw_triples(int32_t M, int32_t n_ransac, int32_t* triples);

/* PointsPlaneFit_RANSAC on pts (M x 3) with the given hypotheses.  keep (M, may be NULL) receives cpFilter.
 * n_ransac <= 4096.  B200BA_E_INVALID when M < 10 (the reference returns -1 there), when an index is out of range, when a coordinate is not
 * finite, or when every hypothesis is degenerate (the reference reads uninitialised memory in that case). */
int b200post_plane_fit(b200post_handle* h, int32_t M, const double* pts, int32_t n_ransac, const int32_t* triples,
                       double k_threshold, b200post_plane* out, int32_t* keep);

/* CalcCameraGround_axes: host arithmetic on 28 scalars (cam = the first camera of the map, src/sfm_tracker.cpp:899). */
int b200post_ground_axes(const double* cam_RT, const double* plane4, const double* plane_se3, double* axes_se3);

/* TransCloudPoint + TransCamera on the device, in place in the caller's arrays; either array may be NULL / empty. */
int b200post_transform(b200post_handle* h, const double* axes_se3, int32_t M, double* pts, int32_t N, double* cam_RT);

/* The whole stage as the tracker strings it together (src/sfm_tracker.cpp:884-904) with the points uploaded once:
 * fit -> mask -> ground axes from cam_RT[0] -> re-axis of the kept points and of all N cameras -> order-preserving
 * compaction.  pts is overwritten with the *m_out kept, transformed points (the rest of the array is left alone),
 * cam_RT (N x 12) with the transformed cameras, keep (M, may be NULL) with cpFilter, axes_se3 (12, may be NULL). */
int b200post_ground_stage(b200post_handle* h, int32_t M, double* pts, int32_t N, double* cam_RT,
                          int32_t n_ransac, const int32_t* triples, double k_threshold,
                          b200post_plane* out, int32_t* keep, int32_t* m_out, double* axes_se3);

/* The same chained onto a finished bundle adjustment WITHOUT the points leaving the device:
 * the adjusted points are taken from `ba` through b200ba_get_result_device (caller order, de-normalised, bit-identical to
 * b200ba_get_result), fitted, masked, re-axed and compacted; only the kept points come back.  `ba` must have returned from
 * b200ba_solve / b200ba_finish and live on the same CUDA device as `h`.  pts_out: room for M x 3 doubles (M = the BA
 * problem's point count), the first *m_out rows are written; cam_RT: N x 12, receives the adjusted AND re-axed cameras;
 * keep: M (may be NULL).  Replaces the sequence adjustBundle_* -> write-back -> PointsPlaneFit_RANSAC -> ... -> TransCamera
 * of src/sfm_tracker.cpp:862-904 with one upload (the BA problem) and one download (the kept points). */
int b200post_ground_stage_ba(b200post_handle* h, b200ba_handle* ba, int32_t n_ransac, const int32_t* triples,
                             double k_threshold, b200post_plane* out, int32_t* keep, int32_t* m_out, double* pts_out,
                             double* cam_RT, double* axes_se3);

/* Time `reps` launches of one kernel of the last b200post_plane_fit problem (still resident) with CUDA events.
 * which: 0 radix histogram sweep, 1 candidate gather sweep, 2 score sweep, 3 mask + moments sweep. */
int b200post_debug_time_kernel(b200post_handle* h, int32_t which, int32_t reps, double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_POST_H */
