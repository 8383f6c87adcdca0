// emvs_cv_shim.h — a CONTAINER-ONLY stand-in for the few OpenCV names that the reference's
// cartesian3dgrid sources mention, so that /root/reference/cartesian3dgrid/src/cartesian3dgrid.cpp
// and its header compile IN PLACE into oracle/_ref/libgrid3d_ref.so (see oracle/Makefile, target ref).
//
// TEST INFRASTRUCTURE ONLY.  cv::Mat here is a dense row-major 2-D buffer with the handful of
// members the Grid3D hot path and huangMedianFilter touch (constructor, at<T>, rows, cols, data,
// type, clone).  Everything that only
// the reference's unused focus-collapse methods call (Sobel, GaussianBlur, Mat arithmetic, ...)
// is declared so those methods type-check, and aborts if it is ever executed: nothing on the
// mapping hot path reaches it (cartesian3dgrid.cpp:192-483 are never called with method = -1,
// mapper_emvs_stereo.cpp:367-369).  No arithmetic of the reference is restated in this file.
#ifndef EMVS_CV_SHIM_H_
#define EMVS_CV_SHIM_H_

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

[[noreturn]] inline void shim_unreachable(const char* what)
{
  std::fprintf(stderr, "emvs_cv_shim: %s is not part of the mapping hot path and has no stand-in\n", what);
  std::abort();
}

struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int x_ = 0, int y_ = 0, int w = 0, int h = 0) : x(x_), y(y_), width(w), height(h) {} };
struct Scalar { double v[4]; double operator[](int i) const { return v[i]; } };

enum { BORDER_REFLECT = 2, THRESH_TOZERO = 3 };

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((unsigned char*)ext), type_(type) {}
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }  // create() zero-fills
  void create(int r, int c, int type)
  {
    rows = r; cols = c; type_ = type;
    own_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elem() + 1, (unsigned char)0);
    data = own_->data();
  }
  template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data)[(size_t)y * cols + x]; }
  template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data)[(size_t)y * cols + x]; }
  size_t elem() const { return type_ == CV_8U ? 1 : 4; }
  int type() const { return type_; }
  Mat clone() const
  {
    Mat m(rows, cols, type_);
    if (data) std::memcpy(m.data, data, (size_t)rows * cols * elem());
    return m;
  }
  // members used only by the reference's unused focus collapses
  Mat mul(const Mat&) const { shim_unreachable("Mat::mul"); }
  Mat t() const { shim_unreachable("Mat::t"); }
  Mat operator()(const Rect&) const { shim_unreachable("Mat::operator()(Rect)"); }
  void copyTo(Mat&) const { shim_unreachable("Mat::copyTo"); }

 private:
  int type_ = CV_32F;
  std::shared_ptr<std::vector<unsigned char>> own_;
};

inline Mat operator+(const Mat&, const Mat&) { shim_unreachable("Mat + Mat"); }
inline Mat operator-(const Mat&, const Mat&) { shim_unreachable("Mat - Mat"); }
inline Mat operator*(const Mat&, const Mat&) { shim_unreachable("Mat * Mat"); }
inline Mat& operator-=(Mat&, const Mat&) { shim_unreachable("Mat -= Mat"); }
inline Mat getGaussianKernel(int, double, int) { shim_unreachable("getGaussianKernel"); }
inline void Sobel(const Mat&, Mat&, int, int, int) { shim_unreachable("Sobel"); }
inline void Laplacian(const Mat&, Mat&, int, int) { shim_unreachable("Laplacian"); }
inline void GaussianBlur(const Mat&, Mat&, Size, double, double = 0, int = 0) { shim_unreachable("GaussianBlur"); }
inline double threshold(const Mat&, Mat&, double, double, int) { shim_unreachable("threshold"); }
inline void sqrt(const Mat&, Mat&) { shim_unreachable("sqrt(Mat)"); }
inline Scalar mean(const Mat&) { shim_unreachable("mean"); }

}  // namespace cv

#endif
