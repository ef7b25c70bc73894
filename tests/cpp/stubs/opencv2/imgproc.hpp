// TEST-ONLY stand-in for <opencv2/imgproc.hpp> (see core.hpp in this directory).
#pragma once
#include "core.hpp"
namespace cv {
enum { COLOR_BGR2Lab = 44, COLOR_Lab2BGR = 56, BORDER_DEFAULT = 4 };
void cvtColor(const Mat& src, Mat& dst, int code);
void bilateralFilter(const Mat& src, Mat& dst, int d, double sigmaColor, double sigmaSpace, int borderType = BORDER_DEFAULT);
}  // namespace cv
