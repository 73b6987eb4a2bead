// Stand-in for the reference's src/sfm_ba.h + src/sfm_data.h (interface declarations only) used by the
// stand-alone adapter test; the SequenceSFM build uses the reference's own headers.
#pragma once
#include <map>
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
struct CloudPoint {                      // src/sfm_data.h:33-43
    cv::Point3d pt; cv::Vec3b rgb; int match_id; std::map<int, int> img_kp; double reprojection_error;
};
int adjustBundle_ssba(std::vector<CloudPoint>& pointcloud, cv::Mat& camIntrinsic,
                      std::map<int, std::vector<cv::KeyPoint> >& imgpts, std::map<int, cv::Matx34d>& Pmats);
int adjustBundle_pba(std::vector<CloudPoint>& pointcloud, cv::Mat& camIntrinsic, cv::Mat& camDistortion,
                     std::map<int, std::vector<cv::KeyPoint> >& imgpts, std::map<int, cv::Matx34d>& Pmats);
