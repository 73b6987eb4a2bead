// sfm_tri_b200.cpp -- drop-in replacement body for SequenceSFM's TriangulatePoints (src/sfm_triangulation.h,
// body src/sfm_triangulation.cpp:95-202) on top of include/b200ba_tri.h.  The callers (src/sfm_tracker.cpp:713,
// src/sfm_cameraMatrices.cpp:289-330) stay untouched; in the reference build this definition replaces the one in
// src/sfm_triangulation.cpp.
#include <cstdio>
#include <vector>

#include "sfm_ba.h"          // CloudPoint (the reference's sfm_data.h through its own headers, or adapter/shim)
#include "b200ba_tri.h"

double TriangulatePoints(const std::vector<cv::KeyPoint>& pt_set1, const std::vector<cv::KeyPoint>& pt_set2,
                         const cv::Mat& K, const cv::Mat& Kinv, const cv::Mat& distcoeff,
                         const cv::Matx34d& P, const cv::Matx34d& P1,
                         std::vector<CloudPoint>& pointcloud, std::vector<cv::KeyPoint>& correspImg1Pt)
{
    correspImg1Pt.clear();                                   // pointcloud is appended to, like the reference (:106-107)
    static b200tri_handle* h = 0;                            // one worker thread calls the tracker (src/utils_gui.cpp:825)
    if (!h && b200tri_create(-1, &h) != B200BA_OK) { fprintf(stderr, "b200tri: %s\n", b200tri_last_error(0)); h = 0; return 0.0; }
    const int n = (int)pt_set1.size();
    std::vector<float> a(2 * (size_t)n), b(2 * (size_t)n), a_und(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { a[2 * i] = pt_set1[i].pt.x; a[2 * i + 1] = pt_set1[i].pt.y; b[2 * i] = pt_set2[i].pt.x; b[2 * i + 1] = pt_set2[i].pt.y; }
    double K9[9], Ki9[9], p[12], p1[12], dist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { K9[3 * r + c] = K.at<double>(r, c); Ki9[3 * r + c] = Kinv.at<double>(r, c); }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) { p[4 * r + c] = P(r, c); p1[4 * r + c] = P1(r, c); }
    int nd = distcoeff.rows * distcoeff.cols;
    if (!(nd == 0 || nd == 4 || nd == 5 || nd == 8)) nd = 0;
    for (int k = 0; k < nd; ++k) dist[k] = distcoeff.at<double>(k);
    std::vector<double> pts(3 * (size_t)n), err((size_t)n);
    b200tri_report rep;
    if (b200tri_triangulate(h, n, a.data(), b.data(), K9, Ki9, dist, nd, p, p1, pts.data(), err.data(), a_und.data(), &rep) != B200BA_OK) {
        fprintf(stderr, "TriangulatePoints (b200): %s\n", b200tri_last_error(h));
        return 0.0;
    }
    for (int i = 0; i < n; ++i) {
        CloudPoint cp;
        cp.pt = cv::Point3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        cp.reprojection_error = err[i];
        cp.match_id = i;
        pointcloud.push_back(cp);
        cv::KeyPoint k = pt_set1[i];
        k.pt.x = a_und[2 * i]; k.pt.y = a_und[2 * i + 1];
        correspImg1Pt.push_back(k);
    }
    return rep.mean_reproj_error;
}
