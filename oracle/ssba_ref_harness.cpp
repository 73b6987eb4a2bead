// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference SSBA-3.0 sources that live under
// /root/reference/Thirdparty/SSBA-3.0 through the exact sequence of calls SequenceSFM makes in
// src/sfm_ba_ssba.cpp:85-233, but on raw arrays instead of OpenCV containers.  Built by
// oracle/Makefile into oracle/_ref/libssba_ref.so (git-ignored).  Nothing in the product path links it.
//
// What is replayed (reference file:line):
//   * K/f0 normalisation, Knorm[2][2] = 1                       src/sfm_ba_ssba.cpp:88-102
//   * cams[i] = (Knorm, R, T) in Pmats order                    src/sfm_ba_ssba.cpp:124-147
//   * measurement = cv::Point(pt) / f0 (integer-rounded pixel)  src/sfm_ba_ssba.cpp:171-178
//   * inlierThreshold = 2/|f0|                                  src/sfm_ba_ssba.cpp:193
//   * ScopedBundleExtrinsicNormalizer + IntrinsicNormalizer     src/sfm_ba_ssba.cpp:202-203
//   * CommonInternalsMetricBundleOptimizer(FULL_BUNDLE_METRIC), tau = 1e-3, maxIterations = 50
//                                                               src/sfm_ba_ssba.cpp:205-217
//   * good_adjustment = (status != 2)                           src/sfm_ba_ssba.cpp:221
// The per-iteration trace is scraped from the optimizer's own verbose output
// (Math/v3d_optimization.cpp:508-978, optimizerVerbosenessLevel = 2) with std::cout switched to 17
// significant digits, so the numbers are the reference's own doubles, not a re-derivation.
#define V3DLIB_ENABLE_SUITESPARSE

#include <Math/v3d_linear.h>
#include <Geometry/v3d_metricbundle.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace V3D;
using namespace std;

namespace {

struct TraceRow {
    int iter;
    double err, lambda, new_err, rho;
    int accepted;   // 1 accepted, 0 rejected, -1 no step taken (stop / LDL failure)
};

// Parse the optimizer's verbose log into one row per LM loop iteration.
static void parse_trace(const string& log, vector<TraceRow>& rows)
{
    istringstream in(log);
    string line;
    TraceRow cur; bool have = false;
    double last_err = NAN, last_lambda = NAN;
    auto flush = [&]() { if (have) rows.push_back(cur); have = false; };
    while (getline(in, line)) {
        double a, b;
        int it;
        if (sscanf(line.c_str(), "SparseLevenbergOptimizer: currentIteration: %d", &it) == 1) {
            flush();
            cur.iter = it; cur.err = last_err; cur.lambda = last_lambda;
            cur.new_err = NAN; cur.rho = NAN; cur.accepted = -1; have = true;
        } else if (sscanf(line.c_str(), "SparseLevenbergOptimizer: |residual|^2 = %lf", &a) == 1) {
            last_err = a; if (have) cur.err = a;
        } else if (sscanf(line.c_str(), "SparseLevenbergOptimizer: lambda = %lf", &a) == 1) {
            // printed before lambda is initialised on iteration 0; fixed up by "initial lambda" below
            last_lambda = a; if (have) cur.lambda = a;
        } else if (sscanf(line.c_str(), "SparseLevenbergOptimizer: initial lambda = %lf", &a) == 1) {
            last_lambda = a; if (have) cur.lambda = a;
        } else if (sscanf(line.c_str(), "SparseLevenbergOptimizer: |new residual|^2 = %lf rho = %lf", &a, &b) == 2) {
            if (have) cur.new_err = a;
        } else if (sscanf(line.c_str(), "SparseLevenbergOptimizer: rho = %lf", &a) == 1) {
            if (have) cur.rho = a;
        } else if (line.find("Improved solution") != string::npos) {
            if (have) cur.accepted = 1;
            last_lambda = std::max(1e-10, last_lambda * 0.1);   // v3d_optimization.h:50
        } else if (line.find("Inferior solution") != string::npos) {
            if (have) cur.accepted = 0;
            last_lambda = last_lambda * 10.0;                   // v3d_optimization.h:49
        }
    }
    flush();
}

} // namespace

extern "C" {

// Returns the value adjustBundle_ssba would return (0 good, -1 discarded).  cam_RT / pts are updated in
// place only when the reference would write them back (status != CONVERGED), exactly like the reference.
// trace: up to trace_cap rows of 6 doubles {iter, err, lambda, new_err, rho, accepted}.
int ssba_ref_adjust(int N, int M, int K,
                    const double* K5,        // fx, skew, cx, fy, cy
                    double* cam_RT,          // N x 12 row-major [R|T]
                    double* pts,             // M x 3
                    const double* obs_xy,    // K x 2 pixels
                    const int* obs_cam, const int* obs_pt,
                    int round_pixels,        // 1 = replicate cv::Point (int) rounding of the SSBA path
                    int max_iterations,
                    double* trace, int trace_cap, int* n_trace,
                    int* status_out, int* iterations_out,
                    double* mean_reproj_px /*[2] before, after*/,
                    double* seconds_minimize)
{
    StdDistortionFunction distortion;

    Matrix3x3d KMat;
    makeIdentityMatrix(KMat);
    KMat[0][0] = K5[0]; KMat[0][1] = K5[1]; KMat[0][2] = K5[2];
    KMat[1][1] = K5[3]; KMat[1][2] = K5[4];
    double const f0 = KMat[0][0];
    Matrix3x3d Knorm = KMat;
    scaleMatrixIP(1.0 / f0, Knorm);
    Knorm[2][2] = 1.0;

    vector<Vector3d> Xs(M);
    for (int j = 0; j < M; ++j) { Xs[j][0] = pts[3*j]; Xs[j][1] = pts[3*j+1]; Xs[j][2] = pts[3*j+2]; }

    vector<CameraMatrix> cams(N);
    for (int i = 0; i < N; ++i) {
        Matrix3x3d R; Vector3d T;
        const double* P = cam_RT + 12 * i;
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r][c] = P[4*r+c]; T[r] = P[4*r+3]; }
        cams[i].setIntrinsic(Knorm);
        cams[i].setRotation(R);
        cams[i].setTranslation(T);
    }

    vector<Vector2d> measurements; vector<int> view, point;
    measurements.reserve(K); view.reserve(K); point.reserve(K);
    for (int k = 0; k < K; ++k) {
        Vector3d p;
        if (round_pixels) {
            // cv::Point cvp = Point2f: saturate_cast<int>(float) == cvRound == round-half-to-even
            p[0] = (double)(int)lrintf((float)obs_xy[2*k]);
            p[1] = (double)(int)lrintf((float)obs_xy[2*k+1]);
        } else { p[0] = obs_xy[2*k]; p[1] = obs_xy[2*k+1]; }
        p[2] = 1.0;
        scaleVectorIP(1.0 / f0, p);
        measurements.push_back(Vector2d(p[0], p[1]));
        view.push_back(obs_cam[k]);
        point.push_back(obs_pt[k]);
    }

    auto mean_err = [&]() {
        double s = 0;
        for (int k = 0; k < K; ++k) {
            Vector2d p = cams[view[k]].projectPoint(distortion, Xs[point[k]]);
            s += norm_L2(f0 * (p - measurements[k]));
        }
        return K ? s / K : 0.0;
    };
    if (mean_reproj_px) mean_reproj_px[0] = mean_err();

    double const inlierThreshold = 2.0 / fabs(f0);
    Matrix3x3d K0 = cams[0].getIntrinsic();
    bool good = false;
    int status = 0, iters = 0;
    string log;
    double secs = 0;
    {
        ScopedBundleExtrinsicNormalizer extNorm(cams, Xs);
        ScopedBundleIntrinsicNormalizer intNorm(cams, measurements, view);
        CommonInternalsMetricBundleOptimizer opt(V3D::FULL_BUNDLE_METRIC, inlierThreshold, K0, distortion,
                                                 cams, Xs, measurements, view, point);
        opt.tau = 1e-3;
        opt.maxIterations = max_iterations;

        ostringstream cap;
        streambuf* old = cout.rdbuf(cap.rdbuf());
        streamsize oldp = cout.precision(17);
        int oldv = V3D::optimizerVerbosenessLevel;
        V3D::optimizerVerbosenessLevel = trace ? 2 : 0;
        auto t0 = chrono::steady_clock::now();
        opt.minimize();
        auto t1 = chrono::steady_clock::now();
        V3D::optimizerVerbosenessLevel = oldv;
        cout.precision(oldp);
        cout.rdbuf(old);
        secs = chrono::duration<double>(t1 - t0).count();
        log = cap.str();
        status = opt.status;
        iters = opt.currentIteration;
        good = (opt.status != 2);
    }
    for (int i = 0; i < N; ++i) cams[i].setIntrinsic(K0);
    if (mean_reproj_px) mean_reproj_px[1] = mean_err();

    if (trace && n_trace) {
        vector<TraceRow> rows; parse_trace(log, rows);
        int n = (int)rows.size() < trace_cap ? (int)rows.size() : trace_cap;
        for (int r = 0; r < n; ++r) {
            double* t = trace + 6 * r;
            t[0] = rows[r].iter; t[1] = rows[r].err; t[2] = rows[r].lambda;
            t[3] = rows[r].new_err; t[4] = rows[r].rho; t[5] = rows[r].accepted;
        }
        *n_trace = n;
    } else if (n_trace) *n_trace = 0;
    if (status_out) *status_out = status;
    if (iterations_out) *iterations_out = iters;
    if (seconds_minimize) *seconds_minimize = secs;

    if (good) {
        for (int j = 0; j < M; ++j) { pts[3*j] = Xs[j][0]; pts[3*j+1] = Xs[j][1]; pts[3*j+2] = Xs[j][2]; }
        for (int i = 0; i < N; ++i) {
            Matrix3x3d R = cams[i].getRotation(); Vector3d T = cams[i].getTranslation();
            double* P = cam_RT + 12 * i;
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) P[4*r+c] = R[r][c]; P[4*r+3] = T[r]; }
        }
        return 0;
    }
    return -1;
}

} // extern "C"
