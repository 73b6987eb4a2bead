// oracle/post_ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the reference's OWN post-BA stage on raw arrays:
//   PointsPlaneFit_RANSAC   src/sfm_utils.cpp:807-977   (with FindSigmaSquared :792-805)
//   CalcCameraGround_axes   src/sfm_utils.cpp:1098-1146 (with LocalCoord2G :994, point2Plane :1070)
//   TransCloudPoint         src/sfm_utils.cpp:1148-1179
//   TransCamera (map form)  src/sfm_utils.cpp:1215-1259
// exactly as SfM_Tracker::bundleAdjustFrames strings them together (src/sfm_tracker.cpp:876-908).
//
// src/sfm_utils.cpp as a whole cannot be compiled in this image (it includes Qt, QGLViewer and
// OpenSURF headers), so oracle/Makefile cuts the two function ranges above out of the file WHERE IT
// LIES under /root/reference into oracle/_ref/sfm_utils_slice.inc (generated, git-ignored, never
// committed) and this harness includes that slice beside the reference's vendored OpenCV 2.4.9
// headers (cv::Matx / cv::Vec are header-only) and its vendored Eigen 3.1.3. The arithmetic that
// runs is therefore the reference's, compiled by the same g++ with the same association order.
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <map>
#include <algorithm>

#include <opencv2/core/core.hpp>
#include <Eigen/Dense>

#include "sfm_data.h"                       // CloudPoint (src/sfm_data.h:33-42)

#define sqr(x) ((x)*(x))                    // src/utils.h:49

using namespace std;
using namespace cv;

// prototypes with the defaults of src/sfm_utils.h:87-107
int PointsPlaneFit_RANSAC(std::vector<CloudPoint> &cp, cv::Vec4d &plane, cv::Matx34d &plane_se3,
                          std::vector<int> &cpFilter, int nRansacs, double kThreshold);
int CalcCameraGround_axes(cv::Matx34d &cam, cv::Vec4d &p, cv::Matx34d &p_se3, cv::Matx34d &axes_se3);
int TransCloudPoint(cv::Matx34d &t_se3, vector<CloudPoint> &cp);
int TransCamera(cv::Matx34d &t_se3, map<int, cv::Matx34d> &cams);

// the reference draws its hypotheses with rand(); the harness interposes a replayable stream so the
// engine under test can be handed the very same triples
static const int *g_stream = 0; static int g_stream_len = 0, g_stream_pos = 0;
static int stream_rand() {
    if (g_stream && g_stream_pos < g_stream_len) return g_stream[g_stream_pos++];
    return ::rand();
}
#define rand stream_rand
#define printf(...) ((void)0)
#include "_ref/sfm_utils_slice.inc"
#undef printf
#undef rand

static Matx34d to_m34(const double *a) { Matx34d m; for (int i = 0; i < 12; ++i) m.val[i] = a[i]; return m; }
static void from_m34(const Matx34d &m, double *a) { for (int i = 0; i < 12; ++i) a[i] = m.val[i]; }

extern "C" {

// rand_stream: the values rand() returns, in order (n_rand of them); the count consumed comes back.
int post_ref_plane_fit(int M, const double *pts, int n_ransac, double k_threshold,
                       const int *rand_stream, int n_rand, int *n_rand_used,
                       double *plane4, double *plane_se3, int *keep)
{
    vector<CloudPoint> cp(M);
    for (int i = 0; i < M; ++i) { cp[i].pt.x = pts[3*i]; cp[i].pt.y = pts[3*i+1]; cp[i].pt.z = pts[3*i+2]; }
    Vec4d plane; Matx34d se3; vector<int> flt;
    g_stream = rand_stream; g_stream_len = n_rand; g_stream_pos = 0;
    int rc = PointsPlaneFit_RANSAC(cp, plane, se3, flt, n_ransac, k_threshold);
    if (n_rand_used) *n_rand_used = g_stream_pos;
    g_stream = 0;
    if (rc != 0) return rc;
    for (int i = 0; i < 4; ++i) plane4[i] = plane[i];
    from_m34(se3, plane_se3);
    for (int i = 0; i < M; ++i) keep[i] = flt[i];
    return 0;
}

int post_ref_ground_axes(const double *cam, const double *plane4, const double *plane_se3, double *axes_se3)
{
    Matx34d c = to_m34(cam), p = to_m34(plane_se3), a;
    Vec4d pl(plane4[0], plane4[1], plane4[2], plane4[3]);
    int rc = CalcCameraGround_axes(c, pl, p, a);
    from_m34(a, axes_se3);
    return rc;
}

int post_ref_transform(const double *axes_se3, int M, double *pts, int N, double *cams)
{
    Matx34d a = to_m34(axes_se3);
    vector<CloudPoint> cp(M);
    for (int i = 0; i < M; ++i) { cp[i].pt.x = pts[3*i]; cp[i].pt.y = pts[3*i+1]; cp[i].pt.z = pts[3*i+2]; }
    TransCloudPoint(a, cp);
    for (int i = 0; i < M; ++i) { pts[3*i] = cp[i].pt.x; pts[3*i+1] = cp[i].pt.y; pts[3*i+2] = cp[i].pt.z; }
    map<int, Matx34d> cm;
    for (int i = 0; i < N; ++i) cm[i] = to_m34(cams + 12*i);
    TransCamera(a, cm);
    for (int i = 0; i < N; ++i) from_m34(cm[i], cams + 12*i);
    return 0;
}

}  // extern "C"
