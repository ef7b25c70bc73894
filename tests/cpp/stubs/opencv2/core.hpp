// TEST-ONLY stand-in for <opencv2/core.hpp>: just enough declarations (no definitions) for the reference's
// include/filter.hpp + include/utils.hpp and integration/filter_b200.cpp to TYPE-CHECK in an image without the OpenCV
// C++ SDK (tests/test_shim_typecheck.py runs g++ -fsyntax-only).  Signatures follow OpenCV 4's cv::Mat.
#pragma once
#include <cstddef>
#include <vector>
typedef unsigned char uchar;
#define CV_8U 0
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
namespace cv {
struct Size { int width, height; };
class Mat {
public:
    Mat();
    Mat(int rows, int cols, int type);
    Mat(int rows, int cols, int type, void* data);
    Mat(Size size, int type);
    Mat(const Mat&);
    Mat& operator=(const Mat&);
    ~Mat();
    int rows, cols;
    uchar* data;
    int channels() const;
    int type() const;
    size_t total() const;
    bool empty() const;
    bool isContinuous() const;
    Size size() const;
    Mat clone() const;
    void convertTo(Mat& m, int rtype) const;
    template <typename T> T* ptr(int row = 0);
    template <typename T> const T* ptr(int row = 0) const;
    template <typename T> T& at(int row, int col);
    template <typename T> const T& at(int row, int col) const;
};
void split(const Mat& m, std::vector<Mat>& mv);
void merge(const std::vector<Mat>& mv, Mat& dst);
}  // namespace cv
