// sfm_post_b200.cpp -- drop-in replacement bodies for the post-BA ground-plane functions of SequenceSFM
// (src/sfm_utils.h:87-107; bodies src/sfm_utils.cpp:807-977, 1098-1259) on top of include/b200ba_post.h.
// SfM_Tracker::bundleAdjustFrames (src/sfm_tracker.cpp:876-908) keeps calling the same five names; in the reference
// build these definitions replace the ones in src/sfm_utils.cpp (INTEGRATION.md shows the #if around them).
//
// The hypotheses are still drawn with rand() in the reference's order (b200post_draw_triples), so a run with the same
// srand() state tests the same planes; every loop over the point cloud runs on the GPU.
//
// GroundStage_b200 is the optional fused form of the tracker's five calls (one upload of the points instead of three).
#include <cstdio>
#include <map>
#include <vector>

#include "sfm_utils.h"       // the reference's own header (or adapter/shim/sfm_utils.h for the stand-alone test)
#include "b200ba_post.h"

namespace {

b200post_handle* post_handle()
{
    static b200post_handle* h = 0;          // one worker thread calls the BA stage (src/utils_gui.cpp:825)
    if (!h && b200post_create(-1, 0, &h) != B200BA_OK) {
        fprintf(stderr, "b200post: %s\n", b200post_last_error(0));
        h = 0;
    }
    return h;
}

void flatten(const std::vector<CloudPoint>& cp, std::vector<double>& pts)
{
    pts.resize(3 * cp.size());
    for (size_t i = 0; i < cp.size(); ++i) { pts[3 * i] = cp[i].pt.x; pts[3 * i + 1] = cp[i].pt.y; pts[3 * i + 2] = cp[i].pt.z; }
}
void to_flat(const cv::Matx34d& m, double* a) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) a[4 * r + c] = m(r, c); }
void from_flat(const double* a, cv::Matx34d& m) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) m(r, c) = a[4 * r + c]; }

}  // namespace

int PointsPlaneFit_RANSAC(std::vector<CloudPoint>& cp, cv::Vec4d& plane, cv::Matx34d& plane_se3,
                          std::vector<int>& cpFilter, int nRansacs, double kThreshold)
{
    const unsigned int nPoints = (unsigned int)cp.size();
    if (nPoints < 10) {
        printf("ERR: updateRansac: too few points (%d) to calc plane.\n", nPoints);
        return -1;
    }
    cpFilter.resize(cp.size(), 0);
    b200post_handle* h = post_handle();
    if (!h) return -1;
    std::vector<double> pts; flatten(cp, pts);
    std::vector<int32_t> triples(3 * (size_t)(nRansacs > 0 ? nRansacs : 0)), keep(cp.size());
    b200post_draw_triples((int32_t)nPoints, nRansacs, triples.data());
    b200post_plane out;
    if (b200post_plane_fit(h, (int32_t)nPoints, pts.data(), nRansacs, triples.data(), kThreshold, &out, keep.data()) != B200BA_OK) {
        fprintf(stderr, "PointsPlaneFit_RANSAC (b200): %s\n", b200post_last_error(h));
        return -1;
    }
    for (size_t i = 0; i < cp.size(); ++i) if (keep[i]) cpFilter[i] = 1;     // the reference only ever sets ones (:899-900)
    printf("Inliners = %d / %d\n", out.n_inliers, nPoints);
    for (int i = 0; i < 4; ++i) plane[i] = out.plane[i];
    from_flat(out.plane_se3, plane_se3);
    return 0;
}

int CalcCameraGround_axes(cv::Matx34d& cam, cv::Vec4d& p, cv::Matx34d& p_se3, cv::Matx34d& axes_se3)
{
    double c[12], pl[4], ps[12], ax[12];
    to_flat(cam, c); to_flat(p_se3, ps);
    for (int i = 0; i < 4; ++i) pl[i] = p[i];
    b200post_ground_axes(c, pl, ps, ax);
    from_flat(ax, axes_se3);
    return 0;
}

int TransCloudPoint(cv::Matx34d& t_se3, std::vector<CloudPoint>& cp)
{
    b200post_handle* h = post_handle();
    if (!h || cp.empty()) return 0;
    double ax[12]; to_flat(t_se3, ax);
    std::vector<double> pts; flatten(cp, pts);
    if (b200post_transform(h, ax, (int32_t)cp.size(), pts.data(), 0, 0) != B200BA_OK) {
        fprintf(stderr, "TransCloudPoint (b200): %s\n", b200post_last_error(h));
        return 0;
    }
    for (size_t i = 0; i < cp.size(); ++i) { cp[i].pt.x = pts[3 * i]; cp[i].pt.y = pts[3 * i + 1]; cp[i].pt.z = pts[3 * i + 2]; }
    return 0;
}

int TransCamera(cv::Matx34d& t_se3, std::map<int, cv::Matx34d>& cams)
{
    b200post_handle* h = post_handle();
    if (!h || cams.empty()) return 0;
    double ax[12]; to_flat(t_se3, ax);
    std::vector<double> flat(12 * cams.size());
    size_t n = 0;
    for (std::map<int, cv::Matx34d>::iterator it = cams.begin(); it != cams.end(); ++it, ++n) to_flat(it->second, &flat[12 * n]);
    if (b200post_transform(h, ax, 0, 0, (int32_t)cams.size(), flat.data()) != B200BA_OK) {
        fprintf(stderr, "TransCamera (b200): %s\n", b200post_last_error(h));
        return 0;
    }
    n = 0;
    for (std::map<int, cv::Matx34d>::iterator it = cams.begin(); it != cams.end(); ++it, ++n) from_flat(&flat[12 * n], it->second);
    return 0;
}

int TransCamera(cv::Matx34d& t_se3, cv::Matx34d& cam_se3)
{
    std::map<int, cv::Matx34d> one;
    one[0] = cam_se3;
    TransCamera(t_se3, one);
    cam_se3 = one[0];
    return 0;
}

// Fused form of src/sfm_tracker.cpp:884-904: fit, drop the outliers from pointCloud (order kept, all other CloudPoint
// fields carried along), ground axes from the first camera of the map, re-axis of points and cameras.
int GroundStage_b200(std::vector<CloudPoint>& pointCloud, std::map<int, cv::Matx34d>& kf2CamMap,
                     cv::Vec4d& p_est, cv::Matx34d& p_se3, cv::Matx34d& axes_se3, int nRansacs, double kThreshold)
{
    const int32_t M = (int32_t)pointCloud.size(), N = (int32_t)kf2CamMap.size();
    b200post_handle* h = post_handle();
    if (!h || M < 10 || N < 1) return -1;
    std::vector<double> pts; flatten(pointCloud, pts);
    std::vector<double> cams(12 * (size_t)N);
    size_t n = 0;
    for (std::map<int, cv::Matx34d>::iterator it = kf2CamMap.begin(); it != kf2CamMap.end(); ++it, ++n) to_flat(it->second, &cams[12 * n]);
    std::vector<int32_t> triples(3 * (size_t)nRansacs), keep((size_t)M);
    b200post_draw_triples(M, nRansacs, triples.data());
    b200post_plane out; int32_t m_out = 0; double ax[12];
    if (b200post_ground_stage(h, M, pts.data(), N, cams.data(), nRansacs, triples.data(), kThreshold, &out, keep.data(), &m_out, ax) != B200BA_OK) {
        fprintf(stderr, "GroundStage_b200: %s\n", b200post_last_error(h));
        return -1;
    }
    printf("Inliners = %d / %d\n", out.n_inliers, M);
    size_t w = 0;
    for (size_t i = 0; i < (size_t)M; ++i) if (keep[i]) {
        if (w != i) pointCloud[w] = pointCloud[i];
        pointCloud[w].pt.x = pts[3 * w]; pointCloud[w].pt.y = pts[3 * w + 1]; pointCloud[w].pt.z = pts[3 * w + 2];
        ++w;
    }
    pointCloud.resize(w);
    n = 0;
    for (std::map<int, cv::Matx34d>::iterator it = kf2CamMap.begin(); it != kf2CamMap.end(); ++it, ++n) from_flat(&cams[12 * n], it->second);
    for (int i = 0; i < 4; ++i) p_est[i] = out.plane[i];
    from_flat(out.plane_se3, p_se3);
    from_flat(ax, axes_se3);
    return 0;
}
