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
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_64FC1 CV_64F
#define CV_8UC3 16
#define CV_MINMAX 32     /* = cv::NORM_MINMAX */
#define CV_GRAY2RGB 8    /* = cv::COLOR_GRAY2RGB */

typedef unsigned char uchar;

// The third-party OpenCV ALGORITHMS that feature_tracker.cpp calls (filled-circle raster,
// pyramidal LK, findFundamentalMat, CLAHE, normalize, goodFeaturesToTrack) are not restated a
// second time here: the stand-ins below forward to the oracle's C restatements
// (oracle/esvio_oracle.c), each of which is pinned bit for bit / to 1e-3 px against real cv2
// outputs (tests/golden).  Only oracle/_ref/libesvio_ref_ft.so links them.
extern "C" {
void ora_disc_half_widths(int r, int* half_width);
void ora_calc_optical_flow_pyr_lk(const uint8_t* prev, const uint8_t* next, int W, int H,
                                  const float* prev_pts, float* next_pts, int n, uint8_t* status,
                                  int win, int max_level, int max_count, double epsilon,
                                  int use_initial_flow, double min_eig_threshold);
void ora_clahe_u8(const uint8_t* src, int W, int H, double clip_limit, int tiles, uint8_t* dst);
void ora_normalize_minmax_u8(const uint8_t* src, size_t n, uint8_t* dst);
int ora_good_features_to_track(const uint8_t* img, int W, int H, const uint8_t* mask,
                               int max_corners, double quality, double min_distance,
                               float* out_xy);
int ora_find_fundamental_mask(const float* pts1, const float* pts2, int n, double thresh,
                              double confidence, int max_iters, uint8_t* mask);
// Optional replacements installed by the harness (ref_ft_set_cv_hooks): the same calls served by
// REAL OpenCV (cv2, through ctypes callbacks) instead of the oracle's restatements -- the
// reference's code then runs on the library it was written for.  Null = the restatement.
typedef void (*esvio_ref_lk_hook)(const uint8_t* prev, const uint8_t* next, int W, int H, const float* prev_pts,
                                  float* next_pts, int n, uint8_t* status, int max_level, int use_initial_flow);
typedef int (*esvio_ref_fm_hook)(const float* pts1, const float* pts2, int n, double thresh, uint8_t* mask);
typedef void (*esvio_ref_img_hook)(const uint8_t* src, int W, int H, uint8_t* dst);
typedef int (*esvio_ref_gftt_hook)(const uint8_t* img, int W, int H, const uint8_t* mask, int max_corners,
                                   double quality, double min_distance, float* out_xy);
extern esvio_ref_lk_hook esvio_ref_hook_lk;
extern esvio_ref_fm_hook esvio_ref_hook_fm;
extern esvio_ref_img_hook esvio_ref_hook_clahe, esvio_ref_hook_normalize;
extern esvio_ref_gftt_hook esvio_ref_hook_gftt;
}

namespace cv {

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
// saturate_cast<int>(float / double) is cvRound: round half to even (SSE cvtss2si / lrint)
template <class T, class U>
inline T saturate_to(U v) { return static_cast<T>(v); }
template <>
inline int saturate_to<int, float>(float v) { return (int)std::lrintf(v); }
template <>
inline int saturate_to<int, double>(double v) { return (int)std::lrint(v); }
template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  // cv::Point_<T>::operator Point_<T2>(): per-coordinate saturate_cast (types.hpp)
  template <class T2>
  operator Point_<T2>() const { return Point_<T2>(saturate_to<T2, T>(x), saturate_to<T2, T>(y)); }
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct Scalar {
  double val[4];
  Scalar() : val{0, 0, 0, 0} {}
  Scalar(double a) : val{a, 0, 0, 0} {}
  Scalar(double a, double b, double c = 0, double d = 0) : val{a, b, c, d} {}
};
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
struct TermCriteria {
  enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
  int type, maxCount;
  double epsilon;
  TermCriteria() : type(COUNT + EPS), maxCount(30), epsilon(0.01) {}
  TermCriteria(int t, int n, double e) : type(t), maxCount(n), epsilon(e) {}
};
enum { OPTFLOW_USE_INITIAL_FLOW = 4, FM_RANSAC = 8 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7 };
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
  static size_t elem(int type) {
    return type == CV_64F ? 8 : (type == CV_32F ? 4 : (type == CV_8UC3 ? 3 : 1));
  }
  // cv::Mat(rows, cols, type, Scalar): every element set to the scalar's first value
  Mat(int r, int c, int type, const Scalar& s) { *this = zeros(Size(c, r), type); fill(s); }
  void fill(const Scalar& s) {
    const size_t n = (size_t)rows * cols;
    if (type_ == CV_64F) std::fill_n(reinterpret_cast<double*>(buf->data()), n, s.val[0]);
    else if (type_ == CV_32F) std::fill_n(reinterpret_cast<float*>(buf->data()), n, (float)s.val[0]);
    else if (type_ == CV_8U) std::fill_n(buf->data(), n, (unsigned char)s.val[0]);
    else
      for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) buf->data()[3 * i + k] = (unsigned char)s.val[k];
  }
  Mat clone() const {  // deep copy (plain assignment shares the data, like cv::Mat)
    Mat m = *this;
    if (buf) m.buf = std::make_shared<std::vector<unsigned char>>(*buf);
    return m;
  }
  Size size() const { return Size(cols, rows); }
  const unsigned char* ptr() const { return buf->data(); }
  unsigned char* ptr() { return buf->data(); }
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
  template <class T>
  const T& at(Point p) const {
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


// ---- what feature_tracker.cpp calls (oracle/_ref/libesvio_ref_ft.so) ------------------------------
// cv::circle: only the FILLED circle (thickness < 0) carries data on this path (the selection
// masks, feature_tracker.cpp:32,118,148); its raster is OpenCV's drawing.cpp Circle(), restated
// once in the oracle (ora_disc_half_widths, pinned against cv2 for r = 1..40, 64).  Outlines
// are drawn only by the SHOW_TRACK visualisation, which this build never enables.
inline void circle(Mat& img, Point c, int radius, const Scalar& color, int thickness = 1, int = 8, int = 0) {
  if (thickness >= 0 || img.empty() || radius < 0) return;
  std::vector<int> hw((size_t)radius + 1);
  ora_disc_half_widths(radius, hw.data());
  for (int dy = -radius; dy <= radius; ++dy) {
    const int y = c.y + dy;
    if (y < 0 || y >= img.rows) continue;
    const int h = hw[(size_t)(dy < 0 ? -dy : dy)];
    const int x0 = std::max(c.x - h, 0), x1 = std::min(c.x + h, img.cols - 1);
    for (int x = x0; x <= x1; ++x) {
      if (img.type() == CV_64F) img.at<double>(y, x) = color.val[0];
      else if (img.type() == CV_32F) img.at<float>(y, x) = (float)color.val[0];
      else if (img.type() == CV_8U) img.at<unsigned char>(y, x) = (unsigned char)color.val[0];
      else img.at<Vec3b>(y, x) = Vec3b((unsigned char)color.val[0], (unsigned char)color.val[1], (unsigned char)color.val[2]);
    }
  }
}
// visualisation only (SHOW_TRACK): present so that the draw functions compile, never executed
inline void arrowedLine(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0, double = 0.1) {}
inline void hconcat(const Mat& a, const Mat&, Mat& dst) { dst = a.clone(); }
inline void vconcat(const Mat& a, const Mat&, Mat& dst) { dst = a.clone(); }
inline void cvtColor(const Mat& a, Mat& dst, int) { Mat t = a.clone(); dst = t; }
template <class S>
inline bool imwrite(const S&, const Mat&) { return false; }
// the image node resizes frames whose size differs from the configuration (never the case here)
inline void resize(const Mat& a, Mat& dst, Size) { Mat t = a.clone(); dst = t; }

// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, winSize, maxLevel,
//                          criteria = (COUNT+EPS, 30, 0.01), flags = 0, minEigThreshold = 1e-4)
inline void calcOpticalFlowPyrLK(const Mat& prev, const Mat& next, const std::vector<Point2f>& prev_pts,
                                 std::vector<Point2f>& next_pts, std::vector<unsigned char>& status,
                                 std::vector<float>& err, Size win = Size(21, 21), int max_level = 3,
                                 TermCriteria crit = TermCriteria(), int flags = 0, double min_eig = 1e-4) {
  const int n = (int)prev_pts.size();
  const bool init = (flags & OPTFLOW_USE_INITIAL_FLOW) != 0;
  if (!init || (int)next_pts.size() != n) next_pts.assign(n, Point2f());
  status.assign(n, 0);
  err.assign(n, 0.f);
  if (n == 0) return;
  if (esvio_ref_hook_lk && win.width == 21 && crit.maxCount == 30 && crit.epsilon == 0.01 && min_eig == 1e-4) {
    esvio_ref_hook_lk(prev.ptr(), next.ptr(), prev.cols, prev.rows, reinterpret_cast<const float*>(prev_pts.data()),
                      reinterpret_cast<float*>(next_pts.data()), n, status.data(), max_level, init ? 1 : 0);
    return;
  }
  ora_calc_optical_flow_pyr_lk(prev.ptr(), next.ptr(), prev.cols, prev.rows,
                               reinterpret_cast<const float*>(prev_pts.data()),
                               reinterpret_cast<float*>(next_pts.data()), n, status.data(), win.width,
                               max_level, crit.maxCount, crit.epsilon, init ? 1 : 0, min_eig);
}

// cv::findFundamentalMat(points1, points2, method, ransacReprojThreshold, confidence, mask)
inline Mat findFundamentalMat(const std::vector<Point2f>& p1, const std::vector<Point2f>& p2, int /*FM_RANSAC*/,
                              double thresh, double confidence, std::vector<unsigned char>& mask) {
  const int n = (int)p1.size();
  mask.assign(n, 0);
  if (n > 0 && esvio_ref_hook_fm && confidence == 0.99)
    esvio_ref_hook_fm(reinterpret_cast<const float*>(p1.data()), reinterpret_cast<const float*>(p2.data()), n, thresh,
                      mask.data());
  else if (n > 0)
    ora_find_fundamental_mask(reinterpret_cast<const float*>(p1.data()), reinterpret_cast<const float*>(p2.data()),
                              n, thresh, confidence, 1000, mask.data());
  return Mat();
}

// cv::createCLAHE() (clipLimit 40, 8 x 8 tiles) -> apply
struct CLAHE {
  void apply(const Mat& src, Mat& dst) const {
    Mat out = Mat::zeros(src.size(), CV_8U);
    if (esvio_ref_hook_clahe) esvio_ref_hook_clahe(src.ptr(), src.cols, src.rows, out.ptr());
    else ora_clahe_u8(src.ptr(), src.cols, src.rows, 40.0, 8, out.ptr());
    dst = out;
  }
};
template <class T>
using Ptr = std::shared_ptr<T>;
inline Ptr<CLAHE> createCLAHE() { return std::make_shared<CLAHE>(); }
// cv::normalize(src, dst, 0, 255, NORM_MINMAX) on CV_8U
inline void normalize(const Mat& src, Mat& dst, double, double, int) {
  Mat out = Mat::zeros(src.size(), CV_8U);
  if (esvio_ref_hook_normalize) esvio_ref_hook_normalize(src.ptr(), src.cols, src.rows, out.ptr());
  else ora_normalize_minmax_u8(src.ptr(), (size_t)src.rows * src.cols, out.ptr());
  dst = out;
}
// cv::goodFeaturesToTrack(image, corners, maxCorners, qualityLevel, minDistance, mask)
inline void goodFeaturesToTrack(const Mat& img, std::vector<Point2f>& corners, int max_corners, double quality,
                                double min_distance, const Mat& mask) {
  corners.clear();
  const size_t cap = max_corners > 0 ? (size_t)max_corners : (size_t)img.rows * img.cols;
  std::vector<float> xy(2 * cap);
  const int k = esvio_ref_hook_gftt
                    ? esvio_ref_hook_gftt(img.ptr(), img.cols, img.rows, mask.empty() ? nullptr : mask.ptr(),
                                          max_corners, quality, min_distance, xy.data())
                    : ora_good_features_to_track(img.ptr(), img.cols, img.rows, mask.empty() ? nullptr : mask.ptr(),
                                                 max_corners, quality, min_distance, xy.data());
  for (int i = 0; i < k; ++i) corners.push_back(Point2f(xy[2 * i], xy[2 * i + 1]));
}

}  // namespace cv

// cvRound(float / double): round half to even, as OpenCV's SSE2 path
inline int cvRound(double v) { return (int)std::lrint(v); }
inline int cvRound(float v) { return (int)std::lrintf(v); }
