// mini_eigen.h -- TEST INFRASTRUCTURE: the smallest stand-in for Eigen that lets the reference's
// UNMODIFIED feature_tracker/src/event_detector/event_detector.{h,cc} compile here (Eigen is not
// in /root/reference -- eigen_catkin downloads it -- nor in this image).  See oracle/Makefile,
// target _ref/libesvio_ref.so.
//
// What the reference's SAE / time-surface / Arc* code needs from Eigen is storage only:
// MatrixXd(W, H) indexed (x, y), column-major like Eigen's default, so offset = x + y * W
// (event_detector.h:74-79, event_detector.cc:47-70,149-166,230-305,308-544).  That part of
// this shim has no arithmetic in it, so those stages run as the reference wrote them.
//
// EventDetector::motioncorrection (event_detector.cc:547-591) additionally uses fixed-size
// float arithmetic (products, transpose, inverse, Matrix3f::exp()).  Those operators are
// provided with the evaluation order the oracle documents for Eigen 3.3 (sum of a 3-term dot
// product as p0 + (p1 + p2), cofactor inverse, scalar operands converted to the matrix scalar
// type first); exp() is forwarded to a hook that the test harness points at the oracle's
// restatement of unsupported/MatrixFunctions.  For that path the shim therefore pins the
// reference's control flow (gates, warp, floor, range checks), not Eigen's kernels.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

extern "C" void esvio_ref_shim_mat3_exp_f(const float* a_row_major, float* out_row_major);

namespace Eigen {

const int Dynamic = -1;

template <class Derived>
struct MatrixBase {
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Derived& derived() { return *static_cast<Derived*>(this); }
  template <class I>
  auto operator[](I i) const -> decltype(std::declval<const Derived&>()[i]) {
    return derived()[i];
  }
};

template <class S, int R, int C>
struct Matrix;

// comma initialiser: m << a, b, c, ...  (row-major fill order, every scalar converted to S)
template <class M>
struct CommaInit {
  M& m;
  int k;
  template <class T>
  CommaInit& operator,(const T& v) {
    m.fill_seq(k++, static_cast<typename M::Scalar>(v));
    return *this;
  }
};

// ---- fixed-size matrices / vectors ------------------------------------------------------------
template <class S, int R, int C>
struct Matrix : MatrixBase<Matrix<S, R, C>> {
  typedef S Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
  S d[R * C];  // column-major, like Eigen
  Matrix() {
    for (int i = 0; i < R * C; ++i) d[i] = S(0);
  }
  template <class S2>
  Matrix(const Matrix<S2, R, C>& o) {
    for (int i = 0; i < R * C; ++i) d[i] = static_cast<S>(o.d[i]);
  }
  Matrix(S a, S b) {
    static_assert(R * C == 2, "2-vector");
    d[0] = a;
    d[1] = b;
  }
  S& operator()(int r, int c) { return d[r + c * R]; }
  const S& operator()(int r, int c) const { return d[r + c * R]; }
  S& operator[](int i) { return d[i]; }
  const S& operator[](int i) const { return d[i]; }
  // vector accessors (feature_tracker.cpp, PinholeCamera::liftProjective)
  S& operator()(int i) { return d[i]; }
  const S& operator()(int i) const { return d[i]; }
  S& x() { return d[0]; }
  S& y() { return d[1]; }
  S& z() { return d[2]; }
  const S& x() const { return d[0]; }
  const S& y() const { return d[1]; }
  const S& z() const { return d[2]; }
  Matrix(S a, S b, S c) {
    static_assert(R * C == 3, "3-vector");
    d[0] = a;
    d[1] = b;
    d[2] = c;
  }
  Matrix operator-(const Matrix& b) const {
    Matrix o;
    for (int i = 0; i < R * C; ++i) o.d[i] = d[i] - b.d[i];
    return o;
  }
  struct Head {  // v.head(n) = w  (visualisation code only)
    Matrix& m;
    int n;
    template <class M2>
    Head& operator=(const M2& w) {
      for (int i = 0; i < n; ++i) m.d[i] = w[i];
      return *this;
    }
  };
  Head head(int n) { return Head{*this, n}; }
  void fill_seq(int k, S v) { d[(k / C) + (k % C) * R] = v; }
  template <class T>
  CommaInit<Matrix> operator<<(const T& v) {
    fill_seq(0, static_cast<S>(v));
    return CommaInit<Matrix>{*this, 1};
  }
  Matrix<S, C, R> transpose() const {
    Matrix<S, C, R> t;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c);
    return t;
  }
  Matrix operator-() const {
    Matrix o;
    for (int i = 0; i < R * C; ++i) o.d[i] = -d[i];
    return o;
  }
  Matrix operator+(const Matrix& b) const {
    Matrix o;
    for (int i = 0; i < R * C; ++i) o.d[i] = d[i] + b.d[i];
    return o;
  }
  // 3x3 only: cofactor inverse (Eigen's compute_inverse<Matrix3f>)
  Matrix inverse() const {
    static_assert(R == 3 && C == 3, "3x3 inverse only");
    const Matrix& m = *this;
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    const S c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const S det = c0 * m(0, 0) + (c1 * m(1, 0) + c2 * m(2, 0));
    const S invdet = S(1) / det;
    Matrix r;
    r(0, 0) = c0 * invdet;
    r(0, 1) = c1 * invdet;
    r(0, 2) = c2 * invdet;
    r(1, 0) = cof(0, 1) * invdet;
    r(1, 1) = cof(1, 1) * invdet;
    r(1, 2) = cof(2, 1) * invdet;
    r(2, 0) = cof(0, 2) * invdet;
    r(2, 1) = cof(1, 2) * invdet;
    r(2, 2) = cof(2, 2) * invdet;
    return r;
  }
  // unsupported/Eigen/MatrixFunctions: forwarded to the harness (see the header comment)
  Matrix exp() const {
    static_assert(R == 3 && C == 3 && sizeof(S) == sizeof(float), "Matrix3f::exp() only");
    float a[9], o[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) a[r * 3 + c] = (*this)(r, c);
    esvio_ref_shim_mat3_exp_f(a, o);
    Matrix e;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) e(r, c) = o[r * 3 + c];
    return e;
  }
};

// coefficient-based product, 3-term sums as p0 + (p1 + p2)
template <class S, int R, int K, int C>
Matrix<S, R, C> operator*(const Matrix<S, R, K>& a, const Matrix<S, K, C>& b) {
  static_assert(K == 3, "inner dimension 3 only");
  Matrix<S, R, C> o;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c)
      o(r, c) = a(r, 0) * b(0, c) + (a(r, 1) * b(1, c) + a(r, 2) * b(2, c));
  return o;
}
// scalar multiples: the scalar is converted to the matrix scalar type first (Eigen 3.3
// promote_scalar_arg), then multiplied coefficient-wise
template <class S, int R, int C, class T,
          class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<S, R, C> operator*(const Matrix<S, R, C>& a, const T& s) {
  const S f = static_cast<S>(s);
  Matrix<S, R, C> o;
  for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] * f;
  return o;
}
template <class S, int R, int C>
Matrix<S, R, C> operator*(const double& s, const Matrix<S, R, C>& a) {
  const S f = static_cast<S>(s);
  Matrix<S, R, C> o;
  for (int i = 0; i < R * C; ++i) o.d[i] = f * a.d[i];
  return o;
}
template <class S, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<S, R, C>& m) {
  for (int r = 0; r < R; ++r) {
    for (int c = 0; c < C; ++c) os << (c ? " " : "") << m(r, c);
    if (r + 1 < R) os << "\n";
  }
  return os;
}

// ---- MatrixXd: storage only ---------------------------------------------------------------------
template <class S>
struct Matrix<S, Dynamic, Dynamic> : MatrixBase<Matrix<S, Dynamic, Dynamic>> {
  typedef S Scalar;
  int rows_, cols_;
  std::vector<S> d;  // column-major: (r, c) at r + c * rows
  Matrix() : rows_(0), cols_(0) {}
  static Matrix Zero(int rows, int cols) {
    Matrix m;
    m.rows_ = rows;
    m.cols_ = cols;
    m.d.assign((size_t)rows * cols, S(0));
    return m;
  }
  S& operator()(int r, int c) { return d[(size_t)r + (size_t)c * rows_]; }
  const S& operator()(int r, int c) const { return d[(size_t)r + (size_t)c * rows_]; }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  const S* data() const { return d.data(); }
  S* data() { return d.data(); }
};

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 4, 1> Vector4d;

}  // namespace Eigen
