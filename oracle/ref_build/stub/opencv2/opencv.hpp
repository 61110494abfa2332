// TEST INFRASTRUCTURE ONLY -- stand-in for <opencv2/opencv.hpp> so that the reference's self-contained
// AP3P solver (src/MiniCVNative/ap3p.cpp, which only needs cv::Matx33d / cv::Vec3d as plain value types)
// can be compiled from where it lies under /root/reference without the OpenCV C++ SDK (not installed in
// this image).  Declares just enough names for MiniCVNative.h and fivepoint.h to parse.  See Makefile.
#pragma once
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef unsigned char uchar;

namespace cv {
struct Point2f { float x, y; };
struct Point2d { double x, y; };
struct Point3d { double x, y, z; };

// 3x3 row-major doubles, 72 bytes, same layout as cv::Matx33d
struct Matx33d {
    double val[9];
    Matx33d() : val{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
    Matx33d(double a0, double a1, double a2, double a3, double a4, double a5, double a6, double a7, double a8)
        : val{a0, a1, a2, a3, a4, a5, a6, a7, a8} {}
};
// 3 doubles, 24 bytes, same layout as cv::Vec3d
struct Vec3d {
    double val[3];
    Vec3d() : val{0, 0, 0} {}
    Vec3d(double a, double b, double c) : val{a, b, c} {}
};

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

struct AKAZE { enum DescriptorType { DESCRIPTOR_KAZE_UPRIGHT = 2, DESCRIPTOR_KAZE = 3, DESCRIPTOR_MLDB_UPRIGHT = 4, DESCRIPTOR_MLDB = 5 }; };
struct KAZE { enum DiffusivityType { DIFF_PM_G1 = 0, DIFF_PM_G2 = 1, DIFF_WEICKERT = 2, DIFF_CHARBONNIER = 3 }; };
struct ORB { enum ScoreType { HARRIS_SCORE = 0, FAST_SCORE = 1 }; };
}  // namespace cv
