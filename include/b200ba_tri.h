/* b200ba_tri.h -- C ABI of the batched two-view triangulation (libb200ba.so), SURVEY.md section 8(f) rank 4.
 *
 * The step that feeds new points to the bundle adjustment.  Reference interface replaced (paths relative to the
 * reference tree):
 *
 *   double TriangulatePoints(const vector<KeyPoint>& pt_set1, const vector<KeyPoint>& pt_set2, const Mat& K,
 *                            const Mat& Kinv, const Mat& distcoeff, const Matx34d& P, const Matx34d& P1,
 *                            vector<CloudPoint>& pointcloud, vector<KeyPoint>& correspImg1Pt)
 *                                                                  src/sfm_triangulation.cpp:95-202
 *   Mat_<double> IterativeLinearLSTriangulation(Point3d u, Matx34d P, Point3d u1, Matx34d P1)     :63-100
 *
 * called by the tracker for every new keyframe pair (src/sfm_tracker.cpp:713) and four times per candidate pose by the
 * essential-matrix disambiguation (src/sfm_cameraMatrices.cpp:289-330).  One CUDA thread per match does everything the
 * reference's loop body does: undistort both pixels (OpenCV 2.4.9 cvUndistortPoints: 5 fixed-point iterations, result
 * rounded to float like the CV_32FC2 matrix it lands in), u = Kinv (x, y, 1), up to 10 re-weighted 4x3 least-squares
 * solves with the reference's float-rounded stopping test |w - P_3 X| <= 1e-4, re-projection through K P1 and the
 * float-rounded pixel error.  The mean error is a fixed-order tree sum (bit-reproducible run to run).
 *
 * Parity: the reference solves each 4x3 system with cv::solve(DECOMP_SVD) (OpenCV's Jacobi SVD, not in the reference
 * tree: OpenCV 2.4.9 is an external dependency); this library uses Householder QR in registers.  Both are backward
 * stable, so points agree to (condition number) x 1e-16; tests/test_gpu_tri.py states the tolerance (1e-9 relative on
 * well-conditioned pairs) and pins the CPU checker against cv::solve / cv::undistortPoints of the image's cv2 wheel.
 * A rank-deficient system (both rays identical) gives NaN here where the SVD returns the minimum-norm solution; the
 * count comes back in n_degenerate.
 *
 * Same conventions as b200ba.h: 0 or a negative B200BA_E_* code, never throws, caller-owned host buffers, no CPU fallback.
 */
#ifndef B200BA_TRI_H
#define B200BA_TRI_H

#include <stdint.h>
#include "b200ba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200tri_handle b200tri_handle;

typedef struct b200tri_report {
    double  mean_reproj_error;  /* what TriangulatePoints returns (mean of the per-match pixel errors in view 2) */
    int32_t n_degenerate;       /* matches whose 4x3 system was rank deficient (point = NaN)                     */
    int32_t gpu_launches;
    double  ms_device;          /* CUDA-event time of copies + kernels                                           */
    double  ms_kernels;
    int32_t reserved[4];
} b200tri_report;

int b200tri_create(int32_t device, b200tri_handle** out);
void b200tri_destroy(b200tri_handle* h);
const char* b200tri_last_error(const b200tri_handle* h);

/* n matches.  pt1, pt2: n x 2 float pixels (KeyPoint::pt of the two views); K9, Kinv9: 3x3 row-major (the reference is
 * handed both); dist: n_dist in {0, 4, 5, 8} distortion coefficients k1 k2 p1 p2 [k3 [k4 k5 k6]] (NULL / 0 = none);
 * P, P1: 3x4 row-major.  Outputs (any may be NULL): pts n x 3 (CloudPoint::pt), reproj_err n
 * (CloudPoint::reprojection_error), pt1_undistorted n x 2 floats (correspImg1Pt). */
int b200tri_triangulate(b200tri_handle* h, int32_t n, const float* pt1, const float* pt2, const double* K9,
                        const double* Kinv9, const double* dist, int32_t n_dist, const double* P, const double* P1,
                        double* pts, double* reproj_err, float* pt1_undistorted, b200tri_report* report);

/* The gating the tracker applies to a fresh triangulation (SfM_Tracker::triangulateTwoFrames, src/sfm_tracker.cpp:718-771;
 * TestTriangulation, src/sfm_cameraMatrices.cpp:180-204), on the results of the LAST b200tri_triangulate call, which are
 * still resident on the device: points in front of P and of P1 (z > 0, at least 75 % each), cut-off = 2.4 x the
 * sorted error at index 4 n / 5 (an exact order statistic: the errors are radix-sorted on the device), keep[i] = in front
 * of P1 and error < cut-off (the reference's status vector is overwritten by its second TestTriangulation call, :725-726).
 * status: 0 ok, -1 fewer than 10 points, -2 fewer than 75 % in front of a camera, -3 at most 10 points survive -
 * the reference's return codes.  keep: n bytes (may be NULL). */
typedef struct b200tri_gate_report {
    int32_t status;
    int32_t n_front_P, n_front_P1, n_kept;
    double  cutoff;
    int32_t gpu_launches;
    double  ms_device;
    int32_t reserved[4];
} b200tri_gate_report;
int b200tri_gate(b200tri_handle* h, uint8_t* keep, b200tri_gate_report* report);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_TRI_H */
