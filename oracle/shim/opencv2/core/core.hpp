// Minimal stand-in for the OpenCV 3.2 types the reference's hot-path sources touch, so those
// sources compile VERBATIM from /root/reference for oracle/_ref (test infrastructure only).
// Not OpenCV: only the members the reference uses, with OpenCV's documented semantics
// (saturate_cast<int>(float) == cvRound == round-half-to-even; Mat is a continuous row-major
// buffer with an optional owning allocation).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

namespace cv {

typedef unsigned char uchar;

template <typename T> static inline T saturate_cast(float v) { return static_cast<T>(v); }
template <> inline int saturate_cast<int>(float v) { return static_cast<int>(lrintf(v)); }
template <typename T> static inline T saturate_cast(double v) { return static_cast<T>(v); }
template <> inline int saturate_cast<int>(double v) { return static_cast<int>(lrint(v)); }
template <typename T> static inline T saturate_cast(int v) { return static_cast<T>(v); }

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U>
  Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
  Point_& operator+=(const Point_& o) {
    x = saturate_cast<T>(x + o.x);
    y = saturate_cast<T>(y + o.y);
    return *this;
  }
  Point_& operator-=(const Point_& o) {
    x = saturate_cast<T>(x - o.x);
    y = saturate_cast<T>(y - o.y);
    return *this;
  }
};
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) {
  return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y));
}
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) {
  return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y));
}
template <typename T> static inline Point_<T> operator-(const Point_<T>& a) {
  return Point_<T>(saturate_cast<T>(-a.x), saturate_cast<T>(-a.y));
}
// OpenCV (core/types.hpp): Point_<T> times float / double / int, both orders; the product is
// formed in the usual C++ promoted type, then saturate_cast back to T.
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, float b) {
  return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T> static inline Point_<T> operator*(float a, const Point_<T>& b) {
  return Point_<T>(saturate_cast<T>(b.x * a), saturate_cast<T>(b.y * a));
}
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, double b) {
  return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T> static inline Point_<T> operator*(double a, const Point_<T>& b) {
  return Point_<T>(saturate_cast<T>(b.x * a), saturate_cast<T>(b.y * a));
}
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, int b) {
  return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T> static inline Point_<T> operator*(int a, const Point_<T>& b) {
  return Point_<T>(saturate_cast<T>(b.x * a), saturate_cast<T>(b.y * a));
}
template <typename T> static inline bool operator==(const Point_<T>& a, const Point_<T>& b) {
  return a.x == b.x && a.y == b.y;
}
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
  Point_<T> tl() const { return Point_<T>(x, y); }
  Point_<T> br() const { return Point_<T>(x + width, y + height); }
  // OpenCV 3.2 (core/types.hpp): contains() takes Point_<T>, so a Point2f handed to a cv::Rect is first
  // converted with saturate_cast<int> (round half to even) and the ROUNDED point is tested.
  bool contains(const Point_<T>& pt) const {
    return x <= pt.x && pt.x < x + width && y <= pt.y && pt.y < y + height;
  }
};
typedef Rect_<int> Rect;

template <typename T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
  Vec(T a, T b) { val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
  Vec(T a, T b, T c, T d) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uint8_t, 3> Vec3b;
typedef Vec<uint8_t, 4> Vec4b;
typedef Vec<int, 2> Vec2i;
typedef Vec<int, 3> Vec3i;
typedef Vec<float, 3> Vec3f;

template <typename T> struct Scalar_ {
  T val[4];
  Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar_(T a, T b = 0, T c = 0, T d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

template <typename T> struct DataType { enum { type = -1, depth = -1, channels = 1 }; };
template <> struct DataType<uint8_t> { enum { type = 0, depth = 0, channels = 1 }; };
template <> struct DataType<float> { enum { type = 5, depth = 5, channels = 1 }; };
template <> struct DataType<double> { enum { type = 6, depth = 6, channels = 1 }; };
template <> struct DataType<Vec3b> { enum { type = 16, depth = 0, channels = 3 }; };
namespace traits {
template <typename T> struct Type { enum { value = DataType<T>::type }; };
}  // namespace traits

#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_32F 5
#define CV_8U 0
#endif

static inline size_t shim_elem_size(int type) {
  switch (type) {
    case 0: return 1;
    case 5: return 4;
    case 6: return 8;
    case 16: return 3;
    default: return 1;
  }
}

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

// Continuous row-major matrix header (optionally owning).
struct Mat {
  uint8_t* data;
  int rows, cols;
  size_t step;
  int type_;
  std::shared_ptr<std::vector<uint8_t> > store;

  Mat() : data(NULL), rows(0), cols(0), step(0), type_(0) {}
  Mat(int r, int c, int type) : data(NULL), rows(0), cols(0), step(0), type_(type) {
    create(r, c, type);
  }
  Mat(int r, int c, int type, void* ext)
      : data(static_cast<uint8_t*>(ext)), rows(r), cols(c), step(c * shim_elem_size(type)),
        type_(type) {}
  void create(int r, int c, int type) {
    type_ = type;
    rows = r;
    cols = c;
    step = c * shim_elem_size(type);
    store.reset(new std::vector<uint8_t>(step * r + 64, 0));
    data = store->data();
  }
  Size size() const { return Size(cols, rows); }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  int type() const { return type_; }
  int channels() const { return type_ == 16 ? 3 : 1; }
  bool isContinuous() const { return true; }
  bool empty() const { return data == NULL; }
  size_t elemSize() const { return shim_elem_size(type_); }
  template <typename T> T& at(int y, int x) {
    return *reinterpret_cast<T*>(data + y * step + x * sizeof(T));
  }
  template <typename T> const T& at(int y, int x) const {
    return *reinterpret_cast<const T*>(data + y * step + x * sizeof(T));
  }
  template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + y * step); }
  template <typename T> const T* ptr(int y = 0) const {
    return reinterpret_cast<const T*>(data + y * step);
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    if (data) std::memcpy(m.data, data, step * rows);
    return m;
  }
};

template <typename T>
struct Mat_ : public Mat {
  Mat_() : Mat() { type_ = DataType<T>::type; }
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(int r, int c, T* ext) : Mat(r, c, DataType<T>::type, ext) {}
  Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { *this = v; }  // filled (cv::Mat_(rows, cols, value))
  Mat_(const Mat& m) : Mat(m) {}
  void create(int r, int c) { Mat::create(r, c, DataType<T>::type); }
  void create(Size sz) { Mat::create(sz.height, sz.width, DataType<T>::type); }
  T& operator()(int y, int x) { return this->template at<T>(y, x); }
  const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
  T& operator()(int i) { return reinterpret_cast<T*>(data)[i]; }
  const T& operator()(int i) const { return reinterpret_cast<const T*>(data)[i]; }
  Mat_& operator=(const T& v) {
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) (*this)(y, x) = v;
    return *this;
  }
};
typedef Mat_<uint8_t> Mat1b;
typedef Mat_<float> Mat1f;
typedef Mat_<double> Mat1d;
typedef Mat_<Vec3b> Mat3b;

// cv::KeyPoint (core/types.hpp): what Flame::detectFeatures collects before KeyPoint::convert
struct KeyPoint {
  Point2f pt;
  float size;
  KeyPoint() : pt(), size(0) {}
  KeyPoint(Point2f p, float s) : pt(p), size(s) {}
  static void convert(const std::vector<KeyPoint>& kps, std::vector<Point2f>& pts) {
    pts.resize(kps.size());
    for (size_t i = 0; i < kps.size(); ++i) pts[i] = kps[i].pt;
  }
};

// Stub (only the legacy scan-line rasterisers, which Flame never calls, use it). Declared so
// rasterization.cc compiles; NOT a faithful cv::LineIterator.
struct LineIterator {
  int count;
  Mat* img_;
  Point p_;
  LineIterator(Mat& img, Point p1, Point p2) : img_(&img), p_(p1) {
    int dx = std::abs(p2.x - p1.x), dy = std::abs(p2.y - p1.y);
    count = (dx > dy ? dx : dy) + 1;
  }
  uint8_t* operator*() { return img_->data + p_.y * img_->step + p_.x * img_->elemSize(); }
  LineIterator operator++(int) { return *this; }
  LineIterator& operator++() { return *this; }
  Point pos() const { return p_; }
};

enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4 };
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right,
                    int borderType, const Scalar& value = Scalar());
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1,
           double delta = 0, int borderType = BORDER_REFLECT_101);

}  // namespace cv
