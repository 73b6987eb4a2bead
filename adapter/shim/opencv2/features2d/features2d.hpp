#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
}
