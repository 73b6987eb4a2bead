// Stand-alone driver for the adapter test: reads a flat problem, rebuilds the tracker containers
// (vector<CloudPoint>, map<int,vector<KeyPoint>>, map<int,Matx34d>), calls the drop-in entry point and writes
// back what the tracker would see afterwards.  Frame ids are deliberately sparse (7*i+3) to exercise the std::map order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "sfm_ba.h"
int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s ssba|pba in.txt out.txt\n", argv[0]); return 2; }
    FILE* f = fopen(argv[2], "r"); if (!f) return 2;
    int N, M, K; double K5[5];
    if (fscanf(f, "%d %d %d %lf %lf %lf %lf %lf", &N, &M, &K, K5, K5 + 1, K5 + 2, K5 + 3, K5 + 4) != 8) return 2;
    std::map<int, cv::Matx34d> Pmats; std::map<int, std::vector<cv::KeyPoint> > imgpts; std::vector<CloudPoint> cloud(M);
    for (int i = 0; i < N; ++i) { cv::Matx34d P; for (int a = 0; a < 12; ++a) if (fscanf(f, "%lf", &P.val[a]) != 1) return 2; Pmats[7 * i + 3] = P; imgpts[7 * i + 3]; }
    for (int j = 0; j < M; ++j) { if (fscanf(f, "%lf %lf %lf", &cloud[j].pt.x, &cloud[j].pt.y, &cloud[j].pt.z) != 3) return 2; cloud[j].match_id = j; cloud[j].reprojection_error = 0; }
    for (int k = 0; k < K; ++k) {
        int c, p; double x, y; if (fscanf(f, "%d %d %lf %lf", &c, &p, &x, &y) != 4) return 2;
        cv::KeyPoint kp; kp.pt.x = (float)x; kp.pt.y = (float)y;
        std::vector<cv::KeyPoint>& v = imgpts[7 * c + 3];
        cloud[p].img_kp[7 * c + 3] = (int)v.size(); v.push_back(kp);
    }
    fclose(f);
    cv::Mat Kmat(3, 3), dist(1, 4);
    Kmat.at<double>(0, 0) = K5[0]; Kmat.at<double>(0, 1) = K5[1]; Kmat.at<double>(0, 2) = K5[2];
    Kmat.at<double>(1, 1) = K5[3]; Kmat.at<double>(1, 2) = K5[4]; Kmat.at<double>(2, 2) = 1.0;
    int ret = strcmp(argv[1], "pba") == 0 ? adjustBundle_pba(cloud, Kmat, dist, imgpts, Pmats) : adjustBundle_ssba(cloud, Kmat, imgpts, Pmats);
    FILE* o = fopen(argv[3], "w"); if (!o) return 2;
    fprintf(o, "%d\n", ret);
    for (std::map<int, cv::Matx34d>::iterator it = Pmats.begin(); it != Pmats.end(); ++it) { for (int a = 0; a < 12; ++a) fprintf(o, "%.17g ", it->second.val[a]); fprintf(o, "\n"); }
    for (int j = 0; j < M; ++j) fprintf(o, "%.17g %.17g %.17g\n", cloud[j].pt.x, cloud[j].pt.y, cloud[j].pt.z);
    fclose(o);
    return 0;
}
