// pagk_cv_shim.hpp -- stand-in for the part of OpenCV the reference's hot-path sources use.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  It exists so that the reference's own, unmodified
// src/gyro_aided_tracker.cpp, src/patch_match.cpp and src/utils.cpp compile here (OpenCV is not installed
// and there is no network) and can serve as the checker for the restatement in oracle/pagk_oracle.cpp.
//
// What is real and what is restated when oracle/_ref/libpagk_ref.so runs:
//   * every line of the reference's three sources: REAL (compiled from /root/reference where it lies);
//   * cv::Point_/Point3_ arithmetic: restated from OpenCV's core/types.hpp templates (saturate_cast<float>
//     of the float/double expression);
//   * cv::Mat / cv::MatExpr: a small 2-D matrix with OpenCV's header-sharing semantics and the lazy
//     expression rules of modules/core/src/matop.cpp for the operators the reference writes
//     (`*`, `+`, `* scalar`, `/ scalar`, `.t()`, `.inv()`, `*=`); the kernels underneath (gemm, invert,
//     scaleAdd, add, transpose) follow the rules measured bit-for-bit against cv2 4.13
//     (tests/golden/matexpr.npz): see small_gemm below;
//   * cv::resize (u8, INTER_LINEAR): bit-exact against cv2 4.13 (tests/golden/pyramid.npz);
//   * cv::parallel_for_: a static split over std::thread at multiples of 64 indices (see parallel_for_ below);
//   * findHomography / findFundamentalMat / calcOpticalFlowPyrLK / BFMatcher / undistortPoints /
//     initUndistortRectifyMap: declared so the sources compile; defined in oracle/ref_harness.cpp
//     (the two RANSAC estimators return a model injected by the test, the others abort: they are
//     not on the path).
//
// Image buffers: every u8 cv::Mat this shim allocates (cv::resize outputs) is followed by one guard row
// that replicates the last row plus one byte, the out-of-bounds convention of oracle/pagk_oracle.cpp
// (PatchMatch::GetPixelValue reads data[step + 1] past the last row, src/patch_match.cpp:399-403).
#pragma once

// (OpenCV's headers pull in most of the standard library; the reference relies on that for <set>, <cassert>, ...)
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_FM_RANSAC 8
#define CV_GRAY2BGR 8
#define CV_RGB2GRAY 7
#define CV_RGBA2GRAY 11

namespace cv {

[[noreturn]] inline void shim_abort(const char *what) {
  std::fprintf(stderr, "pagk cv shim: %s is not available in this stand-in\n", what);
  std::abort();
}

template <typename T> static inline T saturate_cast(float v) { return (T)v; }
template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <typename T> static inline T saturate_cast(int v) { return (T)v; }

// ------------------------------------------------------------------------------------------ points
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U> &o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
};
template <typename T> static inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) {
  return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y));
}
template <typename T> static inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) {
  return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y));
}
template <typename T> static inline Point_<T> operator-(const Point_<T> &a) { return Point_<T>(-a.x, -a.y); }
template <typename T> static inline Point_<T> operator*(const Point_<T> &a, int b) {
  return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T> static inline Point_<T> operator*(const Point_<T> &a, float b) {
  return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T> static inline Point_<T> operator*(const Point_<T> &a, double b) {
  return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T> static inline Point_<T> operator*(float a, const Point_<T> &b) { return b * a; }
template <typename T> static inline Point_<T> operator*(double a, const Point_<T> &b) { return b * a; }
template <typename T> static inline Point_<T> &operator/=(Point_<T> &a, int b) {
  a.x = saturate_cast<T>(a.x / b); a.y = saturate_cast<T>(a.y / b); return a;
}
template <typename T> static inline Point_<T> &operator/=(Point_<T> &a, float b) {
  a.x = saturate_cast<T>(a.x / b); a.y = saturate_cast<T>(a.y / b); return a;
}
template <typename T> static inline Point_<T> &operator/=(Point_<T> &a, double b) {
  a.x = saturate_cast<T>(a.x / b); a.y = saturate_cast<T>(a.y / b); return a;
}
template <typename T> static inline Point_<T> operator/(const Point_<T> &a, int b) { Point_<T> t(a); t /= b; return t; }
template <typename T> static inline Point_<T> operator/(const Point_<T> &a, float b) { Point_<T> t(a); t /= b; return t; }
template <typename T> static inline Point_<T> operator/(const Point_<T> &a, double b) { Point_<T> t(a); t /= b; return t; }
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator+=(Point_<T> &a, const Point_<T> &b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> static inline Point_<T> &operator-=(Point_<T> &a, const Point_<T> &b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> static inline bool operator==(const Point_<T> &a, const Point_<T> &b) { return a.x == b.x && a.y == b.y; }
template <typename T> static inline std::ostream &operator<<(std::ostream &o, const Point_<T> &p) {
  return o << "[" << p.x << ", " << p.y << "]";
}
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
template <typename T> static inline Point3_<T> operator+(const Point3_<T> &a, const Point3_<T> &b) {
  return Point3_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y), saturate_cast<T>(a.z + b.z));
}
template <typename T> static inline Point3_<T> operator-(const Point3_<T> &a, const Point3_<T> &b) {
  return Point3_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y), saturate_cast<T>(a.z - b.z));
}
template <typename T> static inline Point3_<T> operator*(const Point3_<T> &a, int b) {
  return Point3_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b), saturate_cast<T>(a.z * b));
}
template <typename T> static inline Point3_<T> operator*(const Point3_<T> &a, float b) {
  return Point3_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b), saturate_cast<T>(a.z * b));
}
template <typename T> static inline Point3_<T> operator*(const Point3_<T> &a, double b) {
  return Point3_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b), saturate_cast<T>(a.z * b));
}
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range {
  int start, end;
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
};

template <typename T, int N> struct Vec {
  T val[N];
  T &operator[](int i) { return val[i]; }
  const T &operator[](int i) const { return val[i]; }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<double, 4> Scalar_;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
};

struct TermCriteria {
  enum { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
  int type, maxCount;
  double epsilon;
  TermCriteria() : type(0), maxCount(0), epsilon(0) {}
  TermCriteria(int t, int m, double e) : type(t), maxCount(m), epsilon(e) {}
};

enum { NORM_L2 = 4, RANSAC = 8, OPTFLOW_USE_INITIAL_FLOW = 4, INTER_LINEAR = 1, DECOMP_LU = 0 };
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };

// ------------------------------------------------------------------------------------------ Mat
class MatExpr;

class Mat {
 public:
  int rows, cols;
  uchar *data;
  size_t step;  // bytes per row

  Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(int r, int c, int type, void *ext, size_t ext_step = 0) : rows(r), cols(c), data((uchar *)ext), type_(type) {
    step = ext_step ? ext_step : (size_t)c * elemSize();
  }
  // header over a vector of points: N x 1, two channels, no copy (cv::Mat(const std::vector<_Tp>&, copyData=false))
  explicit Mat(const std::vector<Point2f> &v)
      : rows((int)v.size()), cols(1), data((uchar *)v.data()), step(sizeof(Point2f)), type_(CV_32FC2) {}
  Mat(int r, int c, int type, const struct Scalar &s);  // filled with s[0] (single channel use only)
  Mat(const MatExpr &e);
  Mat &operator=(const MatExpr &e);

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * elemSize();
    // one guard row + one byte behind every allocation (see the header comment)
    buf_ = std::shared_ptr<uchar>(new uchar[(size_t)(r + 1) * step + 16](), std::default_delete<uchar[]>());
    data = buf_.get();
  }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize1() const { return depth() == CV_8U ? 1 : depth() == CV_32F ? 4 : 8; }
  size_t elemSize() const { return elemSize1() * channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  size_t step1() const { return step / elemSize1(); }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows == 1; }
  Size size() const { return Size(cols, rows); }

  template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
  template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
  uchar *ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> T &at(int i) { return const_cast<T &>(static_cast<const Mat *>(this)->at<T>(i)); }
  template <typename T> const T &at(int i) const {  // Mat::at(int i0), core/mat.inl.hpp
    if (isContinuous() || rows == 1) return ((const T *)data)[i];
    if (cols == 1) return *(const T *)(data + step * i);
    const int r = i / cols;
    return ((const T *)(data + step * r))[i - r * cols];
  }

  Mat clone() const {
    Mat m;
    copyTo(m);
    return m;
  }
  void copyTo(Mat &dst) const {
    if (empty()) { dst = Mat(); return; }
    if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || !dst.data) dst.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
  }
  Mat operator()(const Rect &r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  void copyTo(Mat &&dst) const { Mat d(dst); copyTo(d); }  // into a view (mMask(roi_rect), im_out.colRange(...))
  Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
  Mat reshape(int cn, int new_rows = 0) const {
    if (!isContinuous() && !(new_rows == 0 || new_rows == rows)) shim_abort("Mat::reshape of a non-continuous matrix");
    Mat m(*this);
    const size_t row_elems = (size_t)cols * channels();
    m.type_ = CV_MAKETYPE(depth(), cn);
    if (new_rows == 0 || new_rows == rows) {
      m.cols = (int)(row_elems / cn);
    } else {
      m.rows = new_rows;
      m.cols = (int)((size_t)rows * row_elems / new_rows / cn);
      m.step = (size_t)m.cols * m.elemSize();
    }
    return m;
  }
  MatExpr t() const;
  MatExpr inv(int method = DECOMP_LU) const;
  static Mat eye(int r, int c, int type) {
    Mat m = zeros(r, c, type);
    for (int i = 0; i < std::min(r, c); ++i) {
      if (m.depth() == CV_32F) m.at<float>(i, i) = 1.f;
      else if (m.depth() == CV_64F) m.at<double>(i, i) = 1.0;
      else m.at<uchar>(i, i) = 1;
    }
    return m;
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }  // create() value-initialises
  static Mat ones(int r, int c, int type) {
    Mat m(r, c, type);
    if (m.depth() != CV_8U || m.channels() != 1) shim_abort("Mat::ones of this type");
    for (int y = 0; y < r; ++y) std::memset(m.data + (size_t)y * m.step, 1, (size_t)c);
    return m;
  }

  std::shared_ptr<uchar> buf_;
  int type_;
};

inline Mat::Mat(int r, int c, int type, const Scalar &sc) : Mat() {
  create(r, c, type);
  if (depth() != CV_8U) shim_abort("Mat(rows, cols, type, Scalar) of this depth");
  for (int y = 0; y < r; ++y) std::memset(data + (size_t)y * step, (int)sc.val[0], (size_t)c * elemSize());
}

template <typename T> struct MatCommaInitializer_ {
  Mat m;
  int idx;
  MatCommaInitializer_(const Mat &m_) : m(m_), idx(0) {}
  MatCommaInitializer_ &operator,(T v) { m.at<T>(idx / m.cols, idx % m.cols) = v; ++idx; return *this; }
  operator Mat() const { return m; }
};
template <typename T> struct Mat_ : public Mat {
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : sizeof(T) == 8 ? CV_64F : CV_8U) {}
};
template <typename T, typename U> static inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, U v) {
  MatCommaInitializer_<T> ci(m);
  return (ci, (T)v);
}

// ---------------------------------------------------------------------------- kernels under MatExpr
// cv::gemm on CV_32F, as measured against cv2 4.13 (tests/golden/make_golden.py):
//   no transpose flag and inner dimension <= 4 ... float accumulation left to right, no FMA;
//   any transpose flag (or a longer inner dimension) ... double accumulation in four interleaved partial sums,
//   one cast to float;   alpha != 1 ... the float result times float(alpha).
// CV_64F: double accumulation.
namespace shim {
inline Mat gemm(const Mat &A, const Mat &B, double alpha, int flags) {
  const bool ta = flags & GEMM_1_T, tb = flags & GEMM_2_T;
  const int m = ta ? A.cols : A.rows, k = ta ? A.rows : A.cols, n = tb ? B.rows : B.cols;
  if ((tb ? B.cols : B.rows) != k || A.type() != B.type() || A.channels() != 1) shim_abort("gemm on these operands");
  Mat D(m, n, A.type());
  if (A.depth() == CV_64F) {
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int t = 0; t < k; ++t) s += (ta ? A.at<double>(t, i) : A.at<double>(i, t)) * (tb ? B.at<double>(j, t) : B.at<double>(t, j));
        D.at<double>(i, j) = s * alpha;
      }
    return D;
  }
  if (A.depth() != CV_32F) shim_abort("gemm on this depth");
  const float fa = (float)alpha;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      float r;
      if (!ta && !tb && k <= 4) {
        float s = A.at<float>(i, 0) * B.at<float>(0, j);
        for (int t = 1; t < k; ++t) s = s + A.at<float>(i, t) * B.at<float>(t, j);
        r = s;
      } else {
        double s[4] = {0, 0, 0, 0};
        int t = 0;
        for (; t <= k - 4; t += 4)
          for (int u = 0; u < 4; ++u)
            s[u] += (double)(ta ? A.at<float>(t + u, i) : A.at<float>(i, t + u)) * (double)(tb ? B.at<float>(j, t + u) : B.at<float>(t + u, j));
        for (; t < k; ++t) s[0] += (double)(ta ? A.at<float>(t, i) : A.at<float>(i, t)) * (double)(tb ? B.at<float>(j, t) : B.at<float>(t, j));
        s[0] += s[1] + s[2] + s[3];
        r = (float)s[0];
      }
      D.at<float>(i, j) = (alpha == 1.0) ? r : r * fa;
    }
  return D;
}
// cv::invert, DECOMP_LU, 2x2 and 3x3 special cases: adjugate in double times 1/det, cast to the element type
inline Mat invert(const Mat &S) {
  if (S.rows != S.cols || (S.rows != 2 && S.rows != 3) || S.channels() != 1) shim_abort("invert on this shape");
  Mat D(S.rows, S.cols, S.type());
  auto get = [&](int r, int c) -> double { return S.depth() == CV_32F ? (double)S.at<float>(r, c) : S.at<double>(r, c); };
  auto put = [&](int r, int c, double v) { if (S.depth() == CV_32F) D.at<float>(r, c) = (float)v; else D.at<double>(r, c) = v; };
  if (S.rows == 2) {
    const double det = get(0, 0) * get(1, 1) - get(0, 1) * get(1, 0);
    if (det == 0.) return D;
    const double d = 1. / det;
    put(0, 0, get(1, 1) * d); put(0, 1, -get(0, 1) * d); put(1, 0, -get(1, 0) * d); put(1, 1, get(0, 0) * d);
    return D;
  }
  const double s00 = get(0, 0), s01 = get(0, 1), s02 = get(0, 2), s10 = get(1, 0), s11 = get(1, 1), s12 = get(1, 2),
               s20 = get(2, 0), s21 = get(2, 1), s22 = get(2, 2);
  const double det = s00 * (s11 * s22 - s12 * s21) - s01 * (s10 * s22 - s12 * s20) + s02 * (s10 * s21 - s11 * s20);
  if (det == 0.) return D;
  const double d = 1. / det;
  put(0, 0, (s11 * s22 - s12 * s21) * d); put(0, 1, (s02 * s21 - s01 * s22) * d); put(0, 2, (s01 * s12 - s02 * s11) * d);
  put(1, 0, (s12 * s20 - s10 * s22) * d); put(1, 1, (s00 * s22 - s02 * s20) * d); put(1, 2, (s02 * s10 - s00 * s12) * d);
  put(2, 0, (s10 * s21 - s11 * s20) * d); put(2, 1, (s01 * s20 - s00 * s21) * d); put(2, 2, (s00 * s11 - s01 * s10) * d);
  return D;
}
inline Mat transpose(const Mat &S) {
  if (S.channels() != 1) shim_abort("transpose of a multi-channel matrix");
  Mat D(S.cols, S.rows, S.type());
  const size_t es = S.elemSize();
  for (int r = 0; r < S.rows; ++r)
    for (int c = 0; c < S.cols; ++c) std::memcpy(D.data + (size_t)c * D.step + (size_t)r * es, S.data + (size_t)r * S.step + (size_t)c * es, es);
  return D;
}
// dst = a*alpha + b*beta elementwise on CV_32F, with the kernels MatOp_AddEx::assign picks:
// alpha == beta == 1 -> cv::add;  beta == 1 -> cv::scaleAdd(a, alpha, b): a*float(alpha) + b;  alpha == 1 likewise.
inline Mat add_ex(const Mat &a, double alpha, const Mat &b, double beta) {
  if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type() || a.depth() != CV_32F || a.channels() != 1)
    shim_abort("add on these operands");
  Mat D(a.rows, a.cols, a.type());
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) {
      const float x = a.at<float>(r, c), y = b.at<float>(r, c);
      float v;
      if (alpha == 1 && beta == 1) v = x + y;
      else if (alpha == 1 && beta == -1) v = x - y;
      else if (alpha == 1) v = y * (float)beta + x;
      else if (beta == 1 && alpha == -1) v = y - x;
      else if (beta == 1) v = x * (float)alpha + y;
      else shim_abort("addWeighted");
      D.at<float>(r, c) = v;
    }
  return D;
}
inline Mat scale(const Mat &a, double alpha) {  // Mat::convertTo(m, type, alpha): float(double(x) * alpha)
  if (a.depth() != CV_32F || a.channels() != 1) shim_abort("scale on this type");
  Mat D(a.rows, a.cols, a.type());
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) D.at<float>(r, c) = (float)((double)a.at<float>(r, c) * alpha);
  return D;
}
}  // namespace shim

// ------------------------------------------------------------------------------------------ MatExpr
// The lazy-expression rules of modules/core/src/matop.cpp for the operator forms the reference writes.
class MatExpr {
 public:
  enum Op { IDENT, T, GEMM, INV, ADDEX };
  Op op;
  Mat a, b;
  double alpha, beta;
  int flags;
  MatExpr() : op(IDENT), alpha(1), beta(0), flags(0) {}
  MatExpr(const Mat &m) : op(IDENT), a(m), alpha(1), beta(0), flags(0) {}
  static MatExpr make(Op o, const Mat &a_, const Mat &b_, double al, double be, int fl) {
    MatExpr e; e.op = o; e.a = a_; e.b = b_; e.alpha = al; e.beta = be; e.flags = fl; return e;
  }
  bool isT() const { return op == T; }
  bool isScaled() const { return op == ADDEX && !b.data; }
  Mat eval() const {
    switch (op) {
      case IDENT: return a;
      case T: { Mat t = shim::transpose(a); return alpha == 1 ? t : shim::scale(t, alpha); }
      case GEMM: return shim::gemm(a, b, alpha, flags);
      case INV: return shim::invert(a);
      case ADDEX: return b.data ? shim::add_ex(a, alpha, b, beta) : shim::scale(a, alpha);
    }
    return Mat();
  }
  operator Mat() const { return eval(); }
  MatExpr t() const {
    if (op == IDENT) return make(T, a, Mat(), 1, 0, 0);
    if (op == T) return alpha == 1 ? MatExpr(a) : make(ADDEX, a, Mat(), alpha, 0, 0);
    return make(T, eval(), Mat(), 1, 0, 0);
  }
  MatExpr inv(int = DECOMP_LU) const { return make(INV, eval(), Mat(), 1, 0, 0); }  // MatOp::invert: evaluate, then MatOp_Invert
  template <typename U> U &at(int, int) { shim_abort("MatExpr::at"); }
};
inline Mat::Mat(const MatExpr &e) : Mat() { *this = e.eval(); }
inline Mat &Mat::operator=(const MatExpr &e) { *this = e.eval(); return *this; }
inline MatExpr Mat::t() const { return MatExpr::make(MatExpr::T, *this, Mat(), 1, 0, 0); }
inline MatExpr Mat::inv(int) const { return MatExpr::make(MatExpr::INV, *this, Mat(), 1, 0, 0); }

// MatOp::matmul: a transposed or scaled operand folds into the gemm's flags / alpha, anything else is evaluated first
inline MatExpr operator*(const MatExpr &e1, const MatExpr &e2) {
  int flags = 0;
  double sc = 1;
  Mat m1, m2;
  if (e1.isT()) { flags |= GEMM_1_T; sc *= e1.alpha; m1 = e1.a; }
  else if (e1.isScaled()) { sc *= e1.alpha; m1 = e1.a; }
  else m1 = e1.eval();
  if (e2.isT()) { flags |= GEMM_2_T; sc *= e2.alpha; m2 = e2.a; }
  else if (e2.isScaled()) { sc *= e2.alpha; m2 = e2.a; }
  else m2 = e2.eval();
  return MatExpr::make(MatExpr::GEMM, m1, m2, sc, 0, flags);
}
inline MatExpr operator*(const Mat &a, const Mat &b) { return MatExpr(a) * MatExpr(b); }
inline MatExpr operator*(const Mat &a, const MatExpr &b) { return MatExpr(a) * b; }
inline MatExpr operator*(const MatExpr &a, const Mat &b) { return a * MatExpr(b); }
// MatOp_AddEx / MatOp_GEMM / MatOp_T ::multiply(e, s): the scalar folds into alpha (double)
inline MatExpr operator*(const MatExpr &e, double s) {
  MatExpr r = e;
  if (e.op == MatExpr::IDENT) return MatExpr::make(MatExpr::ADDEX, e.a, Mat(), s, 0, 0);
  if (e.op == MatExpr::INV) return MatExpr::make(MatExpr::ADDEX, e.eval(), Mat(), s, 0, 0);
  r.alpha *= s; r.beta *= s;
  return r;
}
inline MatExpr operator*(const Mat &a, double s) { return MatExpr(a) * s; }
inline MatExpr operator*(double s, const MatExpr &e) { return e * s; }
inline MatExpr operator*(double s, const Mat &a) { return MatExpr(a) * s; }
inline MatExpr operator/(const MatExpr &e, double s) { return e * (1. / s); }
inline MatExpr operator/(const Mat &a, double s) { return MatExpr(a) * (1. / s); }
// MatOp::add: operands that are plain / purely scaled matrices fold into one AddEx, anything else is evaluated
inline MatExpr operator+(const MatExpr &e1, const MatExpr &e2) {
  // MatOp_GEMM::add folds `A*B + C` (C plain, scaled or transposed) into one gemm; the reference never writes that form
  auto foldable = [](const MatExpr &e) { return e.op == MatExpr::IDENT || e.isScaled() || e.isT(); };
  if ((e1.op == MatExpr::GEMM && foldable(e2)) || (e2.op == MatExpr::GEMM && foldable(e1))) shim_abort("gemm + C folding");
  double alpha = 1, beta = 1;
  Mat m1, m2;
  if (e1.isScaled()) { m1 = e1.a; alpha = e1.alpha; } else m1 = e1.eval();
  if (e2.isScaled()) { m2 = e2.a; beta = e2.alpha; } else m2 = e2.eval();
  return MatExpr::make(MatExpr::ADDEX, m1, m2, alpha, beta, 0);
}
inline MatExpr operator+(const Mat &a, const Mat &b) { return MatExpr(a) + MatExpr(b); }
inline MatExpr operator+(const Mat &a, const MatExpr &b) { return MatExpr(a) + b; }
inline MatExpr operator+(const MatExpr &a, const Mat &b) { return a + MatExpr(b); }
// operator*=(Mat&, const Mat&): gemm(a, b, 1, Mat(), 0, a, 0)
inline Mat &operator*=(Mat &a, const Mat &b) { a = shim::gemm(a, b, 1.0, 0); return a; }

// ------------------------------------------------------------------------------------------ functions
struct _InputArray {
  const Mat *m;
  _InputArray() : m(nullptr) {}
  _InputArray(const Mat &m_) : m(&m_) {}
  Mat getMat() const { return m ? *m : Mat(); }
  bool empty() const { return !m || m->empty(); }
};
struct _OutputArray {
  Mat *m;
  _OutputArray() : m(nullptr) {}
  _OutputArray(Mat &m_) : m(&m_) {}
  void create(int r, int c, int type) const { if (m) m->create(r, c, type); }
  Mat getMat() const { return m ? *m : Mat(); }
  void release() const { if (m) *m = Mat(); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

// cv::resize(src, dst, dsize) with the default INTER_LINEAR on CV_8UC1 (src/patch_match.cpp:69-70)
void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);

inline int &shim_num_threads() { static int n = 1; return n; }
// Static split over std::thread.  Stripe boundaries sit at multiples of 64 indices: PatchMatch::mvSuccess is a
// std::vector<bool> (include/patch_match.h:94) that the parallel body writes per feature (src/patch_match.cpp:351), so two
// stripes sharing a 64-bit word is a data race in the reference (a lost status bit now and then); word-aligned stripes
// keep the reference build deterministic.
template <typename F> inline void parallel_for_(const Range &range, F body) {
  const int n = range.end - range.start;
  const int blocks = (n + 63) / 64, nt = std::max(1, std::min(shim_num_threads(), blocks));
  if (nt <= 1) { body(range); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    const int a = range.start + std::min(n, 64 * (int)((long long)blocks * t / nt));
    const int b = range.start + std::min(n, 64 * (int)((long long)blocks * (t + 1) / nt));
    if (a < b) th.emplace_back([=, &body]() { body(Range(a, b)); });
  }
  for (auto &x : th) x.join();
}

Mat findHomography(const std::vector<Point2f> &src, const std::vector<Point2f> &dst, int method, double thresh, Mat &mask);
Mat findFundamentalMat(const std::vector<Point2f> &p1, const std::vector<Point2f> &p2, int method, double p1_, double p2_, Mat &mask);
void calcOpticalFlowPyrLK(const Mat &prev, const Mat &next, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                          std::vector<uchar> &status, std::vector<float> &err, Size winSize, int maxLevel,
                          TermCriteria criteria, int flags, double minEigThreshold);
void undistortPoints(const Mat &src, Mat &dst, const Mat &K, const Mat &dist, const Mat &R, const Mat &P);
void initUndistortRectifyMap(const Mat &K, const Mat &dist, const Mat &R, const Mat &newK, Size size, int type, Mat &m1, Mat &m2);
Mat getOptimalNewCameraMatrix(const Mat &K, const Mat &dist, Size size, double alpha, Size newSize, void *roi = nullptr);

// ---- for src/ORBextractor.cc (the keypoint top-up; only DetectFeatures / ComputeKeyPointsOctTree are exercised) ----
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif
inline int cvFloor(double v) { const int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { const int i = (int)v; return i + (i < v); }
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline float fastAtan2(float y, float x) {  // degrees in [0, 360); the keypoint angle is never read on the path
  float a = std::atan2(y, x) * 57.29577951308232f;
  return a < 0.f ? a + 360.f : a;
}
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint> &, int) { shim_abort("KeyPointsFilter::retainBest"); }
};
inline void GaussianBlur(const Mat &, Mat &, Size, double, double = 0, int = BORDER_REFLECT_101) { shim_abort("GaussianBlur"); }
// src into the centre of a (rows + top + bottom) x (cols + left + right) dst, borders mirrored without repeating the edge pixel
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType);
// cv::FAST, TYPE_9_16: oracle/pagk_cv_fast.h (bit-exact against cv2.FastFeatureDetector fixtures)
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true);

// ---- not on the path: declared so that src/frame.cpp (Frame::SetPredictKeyPointsAndMask lives there) compiles ----
enum { FONT_HERSHEY_PLAIN = 1, FONT_HERSHEY_DUPLEX = 2, LINE_AA = 16 };
class Exception : public std::exception {
 public:
  std::string err;
  const char *what() const noexcept override { return err.c_str(); }
};
inline void cvtColor(const Mat &, Mat &, int) { shim_abort("cvtColor"); }
inline void goodFeaturesToTrack(const Mat &, std::vector<Point2f> &, int, double, double, const Mat &, int, bool, double) { shim_abort("goodFeaturesToTrack"); }
inline void line(Mat &, Point, Point, const Scalar &, int = 1, int = 8, int = 0) { shim_abort("line"); }
inline void circle(Mat &, Point, int, const Scalar &, int = 1, int = 8, int = 0) { shim_abort("circle"); }
inline void putText(Mat &, const std::string &, Point, int, double, const Scalar &, int = 1, int = 8, bool = false) { shim_abort("putText"); }
inline Size getTextSize(const std::string &, int, double, int, int *) { shim_abort("getTextSize"); }
inline void imshow(const std::string &, const Mat &) { shim_abort("imshow"); }
inline int waitKey(int = 0) { shim_abort("waitKey"); }

struct BFMatcher {
  BFMatcher(int = NORM_L2, bool = false) {}
  void radiusMatch(const Mat &, const Mat &, std::vector<std::vector<DMatch>> &, float) { shim_abort("BFMatcher::radiusMatch"); }
};

}  // namespace cv
