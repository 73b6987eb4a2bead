// Minimal stand-in for the handful of OpenCV 2.4 types that cross SequenceSFM's BA boundary, so that
// adapter/sfm_ba_b200.cpp can be compiled, linked and tested on machines without OpenCV (this image has no
// OpenCV C++ libraries).  In the real SequenceSFM build the genuine <opencv2/core/core.hpp> is used instead.
#pragma once
#include <vector>
#include <cstring>
namespace cv {
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<double> Point3d;
template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
struct Vec3b { unsigned char v[3]; Vec3b() { v[0] = v[1] = v[2] = 0; } };
struct Matx34d {
    double val[12];
    Matx34d() { std::memset(val, 0, sizeof val); }
    double& operator()(int r, int c) { return val[4 * r + c]; }
    const double& operator()(int r, int c) const { return val[4 * r + c]; }
};
struct Vec4d {
    double val[4];
    Vec4d() { std::memset(val, 0, sizeof val); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
    double& operator()(int i) { return val[i]; }
    const double& operator()(int i) const { return val[i]; }
};
// CV_64F matrix, row-major, only what the BA boundary touches: at<double>(r,c), at<double>(i), clone()
class Mat {
public:
    int rows, cols; std::vector<double> d;
    Mat() : rows(0), cols(0) {}
    Mat(int r, int c) : rows(r), cols(c), d((size_t)r * c, 0.0) {}
    template <class T> T& at(int r, int c) { return d[(size_t)r * cols + c]; }
    template <class T> const T& at(int r, int c) const { return d[(size_t)r * cols + c]; }
    template <class T> T& at(int i) { return d[i]; }
    template <class T> const T& at(int i) const { return d[i]; }
    Mat clone() const { return *this; }
};
}
