// sfm_ba_b200.cpp -- drop-in replacement body for SequenceSFM's two BA entry points (src/sfm_ba.h:29-43) on top
// of the B200 engine's C ABI (include/b200ba.h).  It replaces src/sfm_ba_ssba.cpp and src/sfm_ba_pba.cpp in the
// reference build; src/sfm_ba.h, the tracker (src/sfm_tracker.cpp:858-866), triangulation and the Qt/QGLViewer
// front end stay untouched.  See INTEGRATION.md for the CMake change.
//
// Marshalling follows the reference line by line so that the engine sees the numbers SSBA / PBA would see:
//   cameras in std::map key order                         src/sfm_ba_ssba.cpp:130-147 / src/sfm_ba_pba.cpp:75-91
//   points in vector order                                src/sfm_ba_ssba.cpp:109-120
//   observations by point, then by frame id               src/sfm_ba_ssba.cpp:161-182 / src/sfm_ba_pba.cpp:107-127
//   write-back of points and [R|T] only on success        src/sfm_ba_ssba.cpp:235-283
// The SSBA path's integer rounding of measurements (`Point cvp = Point2f`, :171) and its
// "discard a gradient-converged result" rule (:221) are applied inside the engine (options.round_pixels,
// options.discard_converged), not here.
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "sfm_ba.h"          // the reference's own header (or adapter/shim/sfm_ba.h for the stand-alone test)
#include "b200ba.h"

namespace {

struct Flat {
    std::vector<double> cam, pts, xy;
    std::vector<int32_t> oc, op;
    std::vector<int> cam_ids;
    double K5[5];
};

void marshal(const std::vector<CloudPoint>& pointcloud, const cv::Mat& K,
             const std::map<int, std::vector<cv::KeyPoint> >& imgpts, const std::map<int, cv::Matx34d>& Pmats, Flat& f)
{
    f.K5[0] = K.at<double>(0, 0); f.K5[1] = K.at<double>(0, 1); f.K5[2] = K.at<double>(0, 2);
    f.K5[3] = K.at<double>(1, 1); f.K5[4] = K.at<double>(1, 2);
    std::map<int, int> cam_index;
    f.cam.reserve(12 * Pmats.size());
    for (std::map<int, cv::Matx34d>::const_iterator it = Pmats.begin(); it != Pmats.end(); ++it) {
        cam_index[it->first] = (int)f.cam_ids.size();
        f.cam_ids.push_back(it->first);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) f.cam.push_back(it->second(r, c));
    }
    f.pts.reserve(3 * pointcloud.size());
    for (size_t j = 0; j < pointcloud.size(); ++j) {
        const CloudPoint& cp = pointcloud[j];
        f.pts.push_back(cp.pt.x); f.pts.push_back(cp.pt.y); f.pts.push_back(cp.pt.z);
        for (std::map<int, int>::const_iterator m = cp.img_kp.begin(); m != cp.img_kp.end(); ++m) {
            const cv::KeyPoint& kp = imgpts.find(m->first)->second[m->second];
            f.xy.push_back((double)kp.pt.x); f.xy.push_back((double)kp.pt.y);
            f.oc.push_back(cam_index[m->first]);
            f.op.push_back((int32_t)j);
        }
    }
}

// ONE engine handle per option set lives as long as the process (the tracker calls the BA once per keyframe from a single
// worker thread, src/utils_gui.cpp:825): stream, events, pinned staging pool and the grow-only device arena are created once
// instead of on every keyframe (src/sfm_ba_ssba.cpp builds and tears down its whole optimizer per call).  The map between two
// calls is not a pure extension - the post-BA stage drops outliers and re-axes everything (src/sfm_tracker.cpp:876-908) - so
// the problem is uploaded with b200ba_set_problem each time; b200ba_append / b200ba_set_state (include/b200ba.h) serve callers
// whose map only grows.  b200ba_adapter_release() frees the handle (optional, e.g. at shutdown).
b200ba_handle* g_handle = 0;
b200ba_options g_handle_opt;

b200ba_handle* persistent_handle(const b200ba_options& opt, const char* name)
{
    if (g_handle && memcmp(&g_handle_opt, &opt, sizeof opt) != 0) { b200ba_destroy(g_handle); g_handle = 0; }
    if (!g_handle) {
        if (b200ba_create(&opt, &g_handle) != B200BA_OK) { fprintf(stderr, "%s: %s\n", name, b200ba_last_error(0)); g_handle = 0; return 0; }
        g_handle_opt = opt;
    }
    return g_handle;
}

int run(const b200ba_options& opt, const char* name, std::vector<CloudPoint>& pointcloud, const cv::Mat& K,
        std::map<int, std::vector<cv::KeyPoint> >& imgpts, std::map<int, cv::Matx34d>& Pmats)
{
    Flat f; marshal(pointcloud, K, imgpts, Pmats, f);
    const int N = (int)f.cam_ids.size(), M = (int)pointcloud.size(), Kobs = (int)f.oc.size();
    printf("\n%s: \n  N (cams) = %d, M (points) = %d, K (measurements) = %d\n", name, N, M, Kobs);
    b200ba_handle* h = persistent_handle(opt, name);
    if (!h) return -1;
    b200ba_report rep = b200ba_report();
    int rc = b200ba_set_problem(h, N, M, Kobs, f.cam.data(), f.K5, f.pts.data(), f.xy.data(), f.oc.data(), f.op.data());
    if (rc == B200BA_OK) rc = b200ba_solve(h, &rep);
    if (rc != B200BA_OK) { fprintf(stderr, "%s: %s\n", name, b200ba_last_error(h)); return -1; }
    printf("mean reprojection error (in pixels): %g\n", rep.mean_px[0]);
    printf("optimizer status = %d\n", rep.status);
    printf("mean reprojection error (in pixels): %g\n", rep.mean_px[1]);
    int ret = rep.ret;
    if (ret == 0) {
        rc = b200ba_get_result(h, f.cam.data(), f.pts.data());
        if (rc != B200BA_OK) { fprintf(stderr, "%s: %s\n", name, b200ba_last_error(h)); return -1; }
        for (size_t j = 0; j < pointcloud.size(); ++j) {
            pointcloud[j].pt.x = f.pts[3 * j]; pointcloud[j].pt.y = f.pts[3 * j + 1]; pointcloud[j].pt.z = f.pts[3 * j + 2];
        }
        for (int i = 0; i < N; ++i) {
            cv::Matx34d P;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P(r, c) = f.cam[12 * (size_t)i + 4 * r + c];
            Pmats[f.cam_ids[i]] = P;
        }
    }
    return ret;
}

} // namespace

extern "C" void b200ba_adapter_release(void) { if (g_handle) { b200ba_destroy(g_handle); g_handle = 0; } }

int adjustBundle_ssba(std::vector<CloudPoint>& pointcloud, cv::Mat& camIntrinsic,
                      std::map<int, std::vector<cv::KeyPoint> >& imgpts, std::map<int, cv::Matx34d>& Pmats)
{
    b200ba_options opt;
    b200ba_default_options(&opt);          // = the constants hard-coded at src/sfm_ba_ssba.cpp:193-217
    return run(opt, "adjustBundle_ssba", pointcloud, camIntrinsic, imgpts, Pmats);
}

int adjustBundle_pba(std::vector<CloudPoint>& pointcloud, cv::Mat& camIntrinsic, cv::Mat& camDistortion,
                     std::map<int, std::vector<cv::KeyPoint> >& imgpts, std::map<int, cv::Matx34d>& Pmats)
{
#ifdef USE_PBA
    // Built with USE_PBA the reference runs the PBA personality and always returns 0 (src/sfm_ba_pba.cpp:49-166,173).
    (void)camDistortion;                   // read but unused by the reference too (src/sfm_ba_pba.cpp:64-65,87,136)
    b200ba_options opt;
    b200ba_pba_lm_preset(&opt);            // camera 0 constant, no robustifier, sub-pixel measurements + PBA's own LM / PCG / normalisation
    cv::Mat Kp = camIntrinsic.clone();
    double f = (camIntrinsic.at<double>(0, 0) + camIntrinsic.at<double>(1, 1)) / 2.0;   // src/sfm_ba_pba.cpp:60
    Kp.at<double>(0, 0) = f; Kp.at<double>(1, 1) = f; Kp.at<double>(0, 1) = 0.0;
    run(opt, "adjustBundle_pba", pointcloud, Kp, imgpts, Pmats);
    return 0;                              // src/sfm_ba_pba.cpp:173
#else
    // The SHIPPED build leaves USE_PBA undefined: the function forwards to the SSBA path and returns ITS value
    // (src/sfm_ba_pba.cpp:167-171) - integer-rounded pixels, Huber weights, all cameras free, -1 on CONVERGED.  This is the
    // call the tracker makes for every BA with >= 5 keyframes (src/sfm_tracker.cpp:863-866).
    (void)camDistortion;
    int K = 0;                             // the reference prints its own banner first (src/sfm_ba_pba.cpp:43-48)
    for (size_t j = 0; j < pointcloud.size(); ++j) K += (int)pointcloud[j].img_kp.size();
    printf("\nadjustBundle_pba: \n");
    printf("  N (cams) = %d, M (points) = %d, K (measurements) = %d\n", (int)Pmats.size(), (int)pointcloud.size(), K);
    return adjustBundle_ssba(pointcloud, camIntrinsic, imgpts, Pmats);
#endif
}
