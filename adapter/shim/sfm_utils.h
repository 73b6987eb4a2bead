// Stand-in for the four post-BA declarations of the reference's src/sfm_utils.h:87-107 (interface only) used by the
// stand-alone adapter test; the SequenceSFM build uses the reference's own header.
#pragma once
#include <map>
#include <vector>
#include <opencv2/core/core.hpp>
#include "sfm_ba.h"     // CloudPoint
int PointsPlaneFit_RANSAC(std::vector<CloudPoint> &cp, cv::Vec4d &plane, cv::Matx34d &plane_se3,
                          std::vector<int> &cpFilter, int nRansacs = 100, double kThreshold = 2.0);
int CalcCameraGround_axes(cv::Matx34d &cam, cv::Vec4d &p, cv::Matx34d &p_se3, cv::Matx34d &axes_se3);
int TransCloudPoint(cv::Matx34d &t_se3, std::vector<CloudPoint> &cp);
int TransCamera(cv::Matx34d &t_se3, cv::Matx34d &cam_se3);
int TransCamera(cv::Matx34d &t_se3, std::map<int, cv::Matx34d> &cams);
