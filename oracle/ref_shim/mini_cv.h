// mini_cv.h -- TEST INFRASTRUCTURE: the few OpenCV types the reference's UNMODIFIED
// event_detector.{h,cc} touches, so that it compiles here without OpenCV's C++ headers (see
// mini_eigen.h and oracle/Makefile, target _ref/libesvio_ref.so).
//
// OpenCV is a third-party dependency that is not in /root/reference; what is restated here is
// its published behaviour for the calls at event_detector.cc:233,254-264:
//   * Mat::zeros / at<T>: plain storage;
//   * `255.0 * (m + 1.0) / 2.0` and `255.0 * m`: MatExpr folds scalar factors into ONE
//     m * alpha + beta (MatOp_AddEx: alpha = 255 * (1/2) = 127.5, beta = 127.5), evaluated by
//     convertTo in double;
//   * convertTo(CV_8U): cvRound (round half to even) + saturate;
//   * medianBlur: BORDER_REPLICATE median of a k x k window.
// The same arithmetic is pinned against cv2 4.13 by tests/golden (convertTo rounding, median).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_64F 6
#define CV_8UC3 16

namespace cv {

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Point {
  int x, y;
  Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Vec3b {
  unsigned char v[3];
  Vec3b() : v{0, 0, 0} {}
  Vec3b(unsigned char a, unsigned char b, unsigned char c) : v{a, b, c} {}
};

struct MatExpr;

struct Mat {
  int rows, cols, type_;
  std::shared_ptr<std::vector<unsigned char>> buf;  // shared like cv::Mat's refcounted data
  Mat() : rows(0), cols(0), type_(CV_8U) {}
  static size_t elem(int type) { return type == CV_64F ? 8 : (type == CV_8UC3 ? 3 : 1); }
  static Mat zeros(Size s, int type) {
    Mat m;
    m.rows = s.height;
    m.cols = s.width;
    m.type_ = type;
    m.buf = std::make_shared<std::vector<unsigned char>>((size_t)m.rows * m.cols * elem(type), 0);
    return m;
  }
  int type() const { return type_; }
  bool empty() const { return !buf || buf->empty(); }
  template <class T>
  T& at(int y, int x) {
    return reinterpret_cast<T*>(buf->data())[(size_t)y * cols + x];
  }
  template <class T>
  const T& at(int y, int x) const {
    return reinterpret_cast<const T*>(buf->data())[(size_t)y * cols + x];
  }
  template <class T>
  T& at(Point p) {
    return at<T>(p.y, p.x);
  }
  // dst = saturate_cast<dtype>(src * alpha + beta)
  void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
    Mat out = zeros(Size(cols, rows), rtype);
    const size_t n = (size_t)rows * cols;
    for (size_t i = 0; i < n; ++i) {
      double v = type_ == CV_64F ? reinterpret_cast<const double*>(buf->data())[i]
                                 : (double)buf->data()[i];
      v = v * alpha + beta;
      if (rtype == CV_64F) {
        reinterpret_cast<double*>(out.buf->data())[i] = v;
      } else {
        long r = std::lrint(v);  // cvRound: round half to even (default rounding mode)
        out.buf->data()[i] = (unsigned char)(r < 0 ? 0 : (r > 255 ? 255 : r));
      }
    }
    dst = out;
  }
  Mat& operator=(const MatExpr& e);
};

// a * alpha + beta, lazily (MatOp_AddEx with an empty second operand)
struct MatExpr {
  Mat a;
  double alpha, beta;
};
inline MatExpr operator+(const Mat& m, double s) { return MatExpr{m, 1.0, s}; }
inline MatExpr operator*(double s, const Mat& m) { return MatExpr{m, s, 0.0}; }
inline MatExpr operator*(double s, const MatExpr& e) { return MatExpr{e.a, e.alpha * s, e.beta * s}; }
inline MatExpr operator/(const MatExpr& e, double s) {
  const double r = 1.0 / s;  // MatOp::divide(expr, s) = multiply(expr, 1 / s)
  return MatExpr{e.a, e.alpha * r, e.beta * r};
}
inline Mat& Mat::operator=(const MatExpr& e) {
  Mat out;
  e.a.convertTo(out, e.a.type(), e.alpha, e.beta);
  *this = out;
  return *this;
}

inline void medianBlur(const Mat& src, Mat& dst, int ksize) {
  Mat out = Mat::zeros(Size(src.cols, src.rows), CV_8U);
  const int r = ksize / 2;
  std::vector<unsigned char> w((size_t)ksize * ksize);
  for (int y = 0; y < src.rows; ++y)
    for (int x = 0; x < src.cols; ++x) {
      size_t k = 0;
      for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx) {
          const int yy = std::min(std::max(y + dy, 0), src.rows - 1);
          const int xx = std::min(std::max(x + dx, 0), src.cols - 1);
          w[k++] = src.at<unsigned char>(yy, xx);
        }
      std::nth_element(w.begin(), w.begin() + w.size() / 2, w.end());
      out.at<unsigned char>(y, x) = w[w.size() / 2];
    }
  dst = out;
}

}  // namespace cv
