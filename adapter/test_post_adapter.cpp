// Stand-alone driver for the post-BA adapter test: rebuilds the tracker containers from a flat file, runs the sequence of
// src/sfm_tracker.cpp:884-904 through the drop-in functions (mode "calls") or the fused helper (mode "fused") after
// srand(seed), and writes what the tracker would see afterwards.  Frame ids are sparse (5*i+2) to exercise std::map order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "sfm_utils.h"
int GroundStage_b200(std::vector<CloudPoint>&, std::map<int, cv::Matx34d>&, cv::Vec4d&, cv::Matx34d&, cv::Matx34d&, int, double);
int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s calls|fused seed in.txt out.txt\n", argv[0]); return 2; }
    FILE* f = fopen(argv[3], "r"); if (!f) return 2;
    int N, M;
    if (fscanf(f, "%d %d", &N, &M) != 2) return 2;
    std::map<int, cv::Matx34d> cams; std::vector<CloudPoint> cloud(M);
    for (int i = 0; i < N; ++i) { cv::Matx34d P; for (int a = 0; a < 12; ++a) if (fscanf(f, "%lf", &P.val[a]) != 1) return 2; cams[5 * i + 2] = P; }
    for (int j = 0; j < M; ++j) { if (fscanf(f, "%lf %lf %lf", &cloud[j].pt.x, &cloud[j].pt.y, &cloud[j].pt.z) != 3) return 2; cloud[j].match_id = j; }
    fclose(f);
    srand((unsigned)atoi(argv[2]));
    cv::Vec4d p_est; cv::Matx34d p_se3, axes_se3;
    int ret;
    if (strcmp(argv[1], "fused") == 0) {
        ret = GroundStage_b200(cloud, cams, p_est, p_se3, axes_se3, 100, 2.0);
    } else {
        std::vector<int> p_flt;
        ret = PointsPlaneFit_RANSAC(cloud, p_est, p_se3, p_flt);
        if (ret == 0) {
            std::vector<CloudPoint> pc_new;
            for (size_t i = 0; i < cloud.size(); ++i) if (p_flt[i]) pc_new.push_back(cloud[i]);
            cloud = pc_new;
            cv::Matx34d& c1 = cams.begin()->second;
            CalcCameraGround_axes(c1, p_est, p_se3, axes_se3);
            TransCloudPoint(axes_se3, cloud);
            TransCamera(axes_se3, cams);
        }
    }
    FILE* o = fopen(argv[4], "w"); if (!o) return 2;
    fprintf(o, "%d %d\n", ret, (int)cloud.size());
    for (int a = 0; a < 4; ++a) fprintf(o, "%.17g ", p_est[a]);
    fprintf(o, "\n");
    for (int a = 0; a < 12; ++a) fprintf(o, "%.17g ", axes_se3.val[a]);
    fprintf(o, "\n");
    for (std::map<int, cv::Matx34d>::iterator it = cams.begin(); it != cams.end(); ++it) { for (int a = 0; a < 12; ++a) fprintf(o, "%.17g ", it->second.val[a]); fprintf(o, "\n"); }
    for (size_t j = 0; j < cloud.size(); ++j) fprintf(o, "%d %.17g %.17g %.17g\n", cloud[j].match_id, cloud[j].pt.x, cloud[j].pt.y, cloud[j].pt.z);
    fclose(o);
    return 0;
}
