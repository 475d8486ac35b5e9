// STAND-IN for the subset of Eigen 3 that the reference's hot-path headers use
// (/root/reference/include/BALM/tools.hpp, include/BALM/bavoxel.hpp, include/utils.hpp).
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE, and NOT Eigen.  Eigen is not installed in this image and there is no
// network; this file exists so that the reference's own source files can be compiled WHERE THEY LIE (oracle/Makefile,
// target _ref/libbalm_ref.so) and their outputs used to pin the oracles (tests/golden/make_golden_ref.py).  It is
// written from scratch against the call sites in those three headers; it is eager (every operator returns a value, no
// expression templates), which is what makes `const Vector3d& x = a / n;` in the reference bind to a temporary safely.
//
// What differs from real Eigen, and therefore what a fixture made with it does NOT pin:
//   * SelfAdjointEigenSolver: cyclic Jacobi here, tridiagonal QL in Eigen.  Same eigenvalues to a few ulp of ||A||,
//     eigenvector SIGNS differ (the Hessian of bavoxel.hpp:68-174 is even in every eigenvector; `direct` of
//     judge_eigen :349 is not, the tests compare it up to sign);
//   * SimplicialLDLT: natural ordering, envelope storage, no pivoting here; AMD ordering in Eigen.  Same solution up to
//     the rounding of a different elimination order;
//   * colPivHouseholderQr / AngleAxis: minimal versions, reached only from functions outside the hot path
//     (tools.hpp:130-145, :478-506).
// Only the members the three headers use exist.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <type_traits>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;

template <typename T>
using aligned_allocator = std::allocator<T>;

namespace mini {
template <typename T, int R, int C, bool Dyn = (R < 0 || C < 0)>
struct Store {
  T d[R * C];
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int, int) {}
};
template <typename T, int R, int C>
struct Store<T, R, C, true> {
  std::vector<T> d;
  int r = (R < 0 ? 0 : R), c = (C < 0 ? 0 : C);
  int rows() const { return r; }
  int cols() const { return c; }
  void resize(int rr, int cc) { r = rr; c = cc; d.assign((size_t)rr * cc, T(0)); }
};
template <typename S>
using if_scalar = typename std::enable_if<std::is_arithmetic<S>::value, int>::type;
}  // namespace mini

template <typename T, int R, int C> class Matrix;
template <typename T, int R, int C> class BlockRef;
template <typename T> class DiagRef;
template <typename M> class SelfAdjointEigenSolver;

template <typename M>
struct CommaInit {
  M& m;
  int k;
  template <typename S, mini::if_scalar<S> = 0>
  CommaInit& operator,(S v) {
    const int c = m.cols();
    m(k / c, k % c) = (typename M::Scalar)v;     // row by row, as Eigen's comma initialiser fills
    ++k;
    return *this;
  }
};

template <typename T, int R, int C>
struct ColPivQrStandIn {            // least squares through the normal equations; outside the hot path
  Matrix<T, R, C> a;
  template <int RB>
  Matrix<T, C, 1> solve(const Matrix<T, RB, 1>& b) const;
};

template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
  mini::Store<T, R, C> s;

  Matrix() {
    if (R >= 0 && C >= 0) for (int i = 0; i < R * C; ++i) s.d[i] = T(0);
  }
  // dynamic sizes
  template <int RR = R, int CC = C, typename std::enable_if<(RR < 0 && CC < 0), int>::type = 0>
  Matrix(int r, int c) { s.resize(r, c); }
  template <int RR = R, int CC = C, typename std::enable_if<(RR < 0 && CC == 1), int>::type = 0>
  explicit Matrix(int n) { s.resize(n, 1); }
  // coefficients of a fixed 3-vector / 4-vector
  template <typename A, typename B, typename D, int RR = R, int CC = C,
            typename std::enable_if<(RR == 3 && CC == 1), int>::type = 0>
  Matrix(A x, B y, D z) { s.d[0] = (T)x; s.d[1] = (T)y; s.d[2] = (T)z; }
  template <typename A, typename B, typename D, typename E, int RR = R, int CC = C,
            typename std::enable_if<(RR == 4 && CC == 1), int>::type = 0>
  Matrix(A x, B y, D z, E w) { s.d[0] = (T)x; s.d[1] = (T)y; s.d[2] = (T)z; s.d[3] = (T)w; }
  // same shape up to Dynamic
  template <int R2, int C2, typename std::enable_if<(R2 != R || C2 != C), int>::type = 0>
  Matrix(const Matrix<T, R2, C2>& o) { assign(o); }
  template <int R2, int C2, typename std::enable_if<(R2 != R || C2 != C), int>::type = 0>
  Matrix& operator=(const Matrix<T, R2, C2>& o) { assign(o); return *this; }

  int rows() const { return s.rows(); }
  int cols() const { return s.cols(); }
  int size() const { return rows() * cols(); }
  void resize(int r, int c) { s.resize(r, c); }
  void resize(int n) { s.resize(n, 1); }

  // column-major like Eigen's default (nothing in the three headers looks at the raw storage)
  T& operator()(int i, int j) { return s.d[(size_t)j * rows() + i]; }
  const T& operator()(int i, int j) const { return s.d[(size_t)j * rows() + i]; }
  T& operator()(int i) { return s.d[i]; }
  const T& operator()(int i) const { return s.d[i]; }
  T& operator[](int i) { return s.d[i]; }
  const T& operator[](int i) const { return s.d[i]; }
  T* data() { return &s.d[0]; }
  const T* data() const { return &s.d[0]; }
  T& w() { return s.d[3]; }
  const T& w() const { return s.d[3]; }
  template <int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v(i) = s.d[i]; return v; }
  template <int N> Matrix<T, N, 1> tail() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v(i) = s.d[size() - N + i]; return v; }
  bool isZero(T prec = T(1e-12)) const { for (int i = 0; i < size(); ++i) if (!(std::abs(s.d[i]) <= prec)) return false; return true; }   // |x| <= prec * 1
  Matrix normalized() const { Matrix m(*this); m.normalize(); return m; }
  T& x() { return s.d[0]; }
  T& y() { return s.d[1]; }
  T& z() { return s.d[2]; }
  const T& x() const { return s.d[0]; }
  const T& y() const { return s.d[1]; }
  const T& z() const { return s.d[2]; }

  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Zero() { return Matrix(); }
  static Matrix UnitX() { Matrix m; m.s.d[0] = T(1); return m; }
  static Matrix UnitY() { Matrix m; m.s.d[1] = T(1); return m; }
  static Matrix UnitZ() { Matrix m; m.s.d[2] = T(1); return m; }

  Matrix& setZero() { for (int i = 0; i < size(); ++i) s.d[i] = T(0); return *this; }
  Matrix& setOnes() { for (int i = 0; i < size(); ++i) s.d[i] = T(1); return *this; }
  Matrix& setIdentity() {
    setZero();
    for (int i = 0; i < std::min(rows(), cols()); ++i) (*this)(i, i) = T(1);
    return *this;
  }

  template <typename S, mini::if_scalar<S> = 0>
  CommaInit<Matrix> operator<<(S v) {
    CommaInit<Matrix> c{*this, 0};
    c, v;
    return c;
  }

  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> t;
    t.resize(cols(), rows());
    for (int i = 0; i < rows(); ++i) for (int j = 0; j < cols(); ++j) t(j, i) = (*this)(i, j);
    return t;
  }
  T trace() const { T t = T(0); for (int i = 0; i < std::min(rows(), cols()); ++i) t += (*this)(i, i); return t; }
  T squaredNorm() const { T t = T(0); for (int i = 0; i < size(); ++i) t += s.d[i] * s.d[i]; return t; }
  T norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { const T n = norm(); for (int i = 0; i < size(); ++i) s.d[i] /= n; }
  bool allFinite() const { for (int i = 0; i < size(); ++i) if (!std::isfinite(s.d[i])) return false; return true; }
  template <int R2, int C2>
  T dot(const Matrix<T, R2, C2>& o) const {      // both operands are vectors of either orientation (tools.hpp:496)
    T t = T(0);
    for (int i = 0; i < size(); ++i) t += s.d[i] * o.s.d[i];
    return t;
  }
  Matrix<T, 3, 1> cross(const Matrix<T, 3, 1>& o) const {
    return Matrix<T, 3, 1>(s.d[1] * o[2] - s.d[2] * o[1], s.d[2] * o[0] - s.d[0] * o[2], s.d[0] * o[1] - s.d[1] * o[0]);
  }
  Matrix cwiseMin(const Matrix& o) const { Matrix m(*this); for (int i = 0; i < size(); ++i) m.s.d[i] = std::min(s.d[i], o.s.d[i]); return m; }
  Matrix cwiseMax(const Matrix& o) const { Matrix m(*this); for (int i = 0; i < size(); ++i) m.s.d[i] = std::max(s.d[i], o.s.d[i]); return m; }

  Matrix<T, R, 1> col(int j) const {
    Matrix<T, R, 1> v;
    v.resize(rows(), 1);
    for (int i = 0; i < rows(); ++i) v(i) = (*this)(i, j);
    return v;
  }
  Matrix<T, 1, C> row(int i) const {
    Matrix<T, 1, C> v;
    v.resize(1, cols());
    for (int j = 0; j < cols(); ++j) v(j) = (*this)(i, j);
    return v;
  }

  template <int BR, int BC> BlockRef<T, BR, BC> block(int i, int j) { return BlockRef<T, BR, BC>(&(*this)(0, 0), rows(), i, j); }
  template <int BR, int BC>
  Matrix<T, BR, BC> block(int i, int j) const {
    Matrix<T, BR, BC> b;
    for (int a = 0; a < BR; ++a) for (int c = 0; c < BC; ++c) b(a, c) = (*this)(i + a, j + c);
    return b;
  }
  DiagRef<T> diagonal() { return DiagRef<T>(&(*this)(0, 0), rows(), std::min(rows(), cols())); }
  Matrix<T, Dynamic, 1> diagonal() const {
    Matrix<T, Dynamic, 1> d(std::min(rows(), cols()));
    for (int i = 0; i < d.rows(); ++i) d(i) = (*this)(i, i);
    return d;
  }

  template <int R2, int C2> Matrix& operator+=(const Matrix<T, R2, C2>& o) { for (int i = 0; i < size(); ++i) s.d[i] += o.s.d[i]; return *this; }
  template <int R2, int C2> Matrix& operator-=(const Matrix<T, R2, C2>& o) { for (int i = 0; i < size(); ++i) s.d[i] -= o.s.d[i]; return *this; }
  template <typename S, mini::if_scalar<S> = 0> Matrix& operator*=(S v) { for (int i = 0; i < size(); ++i) s.d[i] *= (T)v; return *this; }
  template <typename S, mini::if_scalar<S> = 0> Matrix& operator/=(S v) { for (int i = 0; i < size(); ++i) s.d[i] /= (T)v; return *this; }
  Matrix operator-() const { Matrix m(*this); for (int i = 0; i < size(); ++i) m.s.d[i] = -s.d[i]; return m; }

  ColPivQrStandIn<T, R, C> colPivHouseholderQr() const { return ColPivQrStandIn<T, R, C>{*this}; }

 private:
  template <int R2, int C2>
  void assign(const Matrix<T, R2, C2>& o) {
    s.resize(o.rows(), o.cols());
    for (int j = 0; j < o.cols(); ++j) for (int i = 0; i < o.rows(); ++i) (*this)(i, j) = o(i, j);
  }
};

// ---------------------------------------------------------------- arithmetic (eager)
template <typename T, int R, int C>
Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m(a); m += b; return m; }
template <typename T, int R, int C>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m(a); m -= b; return m; }
template <typename T, int R, int C, typename S, mini::if_scalar<S> = 0>
Matrix<T, R, C> operator*(S v, const Matrix<T, R, C>& a) { Matrix<T, R, C> m(a); m *= v; return m; }
template <typename T, int R, int C, typename S, mini::if_scalar<S> = 0>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, S v) { Matrix<T, R, C> m(a); m *= v; return m; }
template <typename T, int R, int C, typename S, mini::if_scalar<S> = 0>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, S v) { Matrix<T, R, C> m(a); m /= v; return m; }
template <typename T, int R, int K, int K2, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K2, C>& b) {
  Matrix<T, R, C> m;
  m.resize(a.rows(), b.cols());
  for (int j = 0; j < b.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) {
      T t = T(0);
      for (int k = 0; k < a.cols(); ++k) t += a(i, k) * b(k, j);
      m(i, j) = t;
    }
  return m;
}

// ---------------------------------------------------------------- writable views
// A view IS-A matrix holding a snapshot of the viewed coefficients (so that it takes part in every expression above
// through derived-to-base deduction) and writes through on assignment.
template <typename T, int R, int C>
class BlockRef : public Matrix<T, R, C> {
  T* base_;
  int ld_, i0_, j0_;
  void store() { for (int a = 0; a < R; ++a) for (int c = 0; c < C; ++c) base_[(size_t)(j0_ + c) * ld_ + i0_ + a] = (*this)(a, c); }

 public:
  BlockRef(T* base, int ld, int i0, int j0) : base_(base), ld_(ld), i0_(i0), j0_(j0) {
    for (int a = 0; a < R; ++a) for (int c = 0; c < C; ++c) (*this)(a, c) = base_[(size_t)(j0_ + c) * ld_ + i0_ + a];
  }
  BlockRef(const BlockRef&) = default;
  BlockRef& operator=(const BlockRef& o) { Matrix<T, R, C>::operator=(static_cast<const Matrix<T, R, C>&>(o)); store(); return *this; }
  template <int R2, int C2> BlockRef& operator=(const Matrix<T, R2, C2>& o) { for (int a = 0; a < R; ++a) for (int c = 0; c < C; ++c) (*this)(a, c) = o(a, c); store(); return *this; }
  template <int R2, int C2> BlockRef& operator+=(const Matrix<T, R2, C2>& o) { for (int a = 0; a < R; ++a) for (int c = 0; c < C; ++c) (*this)(a, c) += o(a, c); store(); return *this; }
  template <int R2, int C2> BlockRef& operator-=(const Matrix<T, R2, C2>& o) { for (int a = 0; a < R; ++a) for (int c = 0; c < C; ++c) (*this)(a, c) -= o(a, c); store(); return *this; }
  template <typename S, mini::if_scalar<S> = 0> BlockRef& operator*=(S v) { Matrix<T, R, C>::operator*=(v); store(); return *this; }
  template <typename S, mini::if_scalar<S> = 0> BlockRef& operator/=(S v) { Matrix<T, R, C>::operator/=(v); store(); return *this; }
};

template <typename T>
class DiagRef : public Matrix<T, Dynamic, 1> {
  T* base_;
  int ld_, n_;
  void store() { for (int i = 0; i < n_; ++i) base_[(size_t)i * ld_ + i] = (*this)(i); }

 public:
  DiagRef(T* base, int ld, int n) : Matrix<T, Dynamic, 1>(n), base_(base), ld_(ld), n_(n) {
    for (int i = 0; i < n_; ++i) (*this)(i) = base_[(size_t)i * ld_ + i];
  }
  DiagRef(const DiagRef&) = default;
  DiagRef& operator=(const DiagRef& o) { for (int i = 0; i < n_; ++i) (*this)(i) = o(i); store(); return *this; }
  template <int R2, int C2> DiagRef& operator=(const Matrix<T, R2, C2>& o) { for (int i = 0; i < n_; ++i) (*this)(i) = o(i); store(); return *this; }
};

template <typename T, int R, int C>
template <int RB>
Matrix<T, C, 1> ColPivQrStandIn<T, R, C>::solve(const Matrix<T, RB, 1>& b) const {
  const int n = a.cols();
  std::vector<T> N((size_t)n * n), y(n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) { T t = T(0); for (int k = 0; k < a.rows(); ++k) t += a(k, i) * a(k, j); N[(size_t)i * n + j] = t; }
    T t = T(0); for (int k = 0; k < a.rows(); ++k) t += a(k, i) * b(k); y[i] = t;
  }
  for (int c = 0; c < n; ++c) {                         // Gaussian elimination with partial pivoting
    int p = c;
    for (int r = c + 1; r < n; ++r) if (std::abs(N[(size_t)r * n + c]) > std::abs(N[(size_t)p * n + c])) p = r;
    for (int k = 0; k < n; ++k) std::swap(N[(size_t)c * n + k], N[(size_t)p * n + k]);
    std::swap(y[c], y[p]);
    for (int r = c + 1; r < n; ++r) {
      const T f = N[(size_t)r * n + c] / N[(size_t)c * n + c];
      for (int k = c; k < n; ++k) N[(size_t)r * n + k] -= f * N[(size_t)c * n + k];
      y[r] -= f * y[c];
    }
  }
  Matrix<T, C, 1> x;
  x.resize(n, 1);
  for (int r = n - 1; r >= 0; --r) {
    T t = y[r];
    for (int k = r + 1; k < n; ++k) t -= N[(size_t)r * n + k] * x(k);
    x(r) = t / N[(size_t)r * n + r];
  }
  return x;
}

typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;

// ---------------------------------------------------------------- symmetric eigen-solver (3x3 on the path, 4x4 in the DLT)
template <typename M>
class SelfAdjointEigenSolver {
  typedef typename M::Scalar T;
  M vec_;
  Matrix<T, Dynamic, 1> val_dyn_;
  Matrix<T, 3, 1> val3_;
  Matrix<T, 4, 1> val4_;
  int info_ = 0;

 public:
  SelfAdjointEigenSolver() {}
  explicit SelfAdjointEigenSolver(const M& m) { compute(m); }
  // Ascending eigenvalues, eigenvectors in the columns.  Reads the lower triangle only, as Eigen documents.  Cyclic Jacobi.
  SelfAdjointEigenSolver& compute(const M& m) {
    const int n = m.rows();
    std::vector<T> a((size_t)n * n), v((size_t)n * n, T(0));
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { a[i * n + j] = (i >= j) ? m(i, j) : m(j, i); }
    for (int i = 0; i < n; ++i) v[i * n + i] = T(1);
    for (int sweep = 0; sweep < 64; ++sweep) {
      T off = T(0);
      for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off += std::abs(a[p * n + q]);
      if (off == T(0)) break;
      for (int p = 0; p < n - 1; ++p)
        for (int q = p + 1; q < n; ++q) {
          if (a[p * n + q] == T(0)) continue;
          const T theta = (a[q * n + q] - a[p * n + p]) / (T(2) * a[p * n + q]);
          const T t = (theta >= T(0) ? T(1) : T(-1)) / (std::abs(theta) + std::sqrt(theta * theta + T(1)));
          const T c = T(1) / std::sqrt(t * t + T(1)), sn = t * c;
          for (int k = 0; k < n; ++k) { const T akp = a[k * n + p], akq = a[k * n + q]; a[k * n + p] = c * akp - sn * akq; a[k * n + q] = sn * akp + c * akq; }
          for (int k = 0; k < n; ++k) { const T apk = a[p * n + k], aqk = a[q * n + k]; a[p * n + k] = c * apk - sn * aqk; a[q * n + k] = sn * apk + c * aqk; }
          for (int k = 0; k < n; ++k) { const T vkp = v[k * n + p], vkq = v[k * n + q]; v[k * n + p] = c * vkp - sn * vkq; v[k * n + q] = sn * vkp + c * vkq; }
          a[p * n + q] = a[q * n + p] = T(0);
        }
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int x, int y) { return a[x * n + x] < a[y * n + y]; });
    vec_ = m;
    val_dyn_.resize(n);
    for (int j = 0; j < n; ++j) {
      const T lam = a[idx[j] * n + idx[j]];
      val_dyn_(j) = lam;
      if (n == 3) val3_(j) = lam;
      if (n == 4) val4_(j) = lam;
      for (int i = 0; i < n; ++i) vec_(i, j) = v[i * n + idx[j]];
    }
    info_ = 0;
    return *this;
  }
  template <int R = M::RowsAtCompileTime> typename std::enable_if<R == 3, const Matrix<T, 3, 1>&>::type eigenvalues() const { return val3_; }
  template <int R = M::RowsAtCompileTime> typename std::enable_if<R == 4, const Matrix<T, 4, 1>&>::type eigenvalues() const { return val4_; }
  template <int R = M::RowsAtCompileTime> typename std::enable_if<(R != 3 && R != 4), const Matrix<T, Dynamic, 1>&>::type eigenvalues() const { return val_dyn_; }
  const M& eigenvectors() const { return vec_; }
  int info() const { return info_; }
};
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

// ---------------------------------------------------------------- rotations (outside the hot path)
template <typename T> class AngleAxis;
template <typename T>
class Quaternion {
 public:
  T w_, x_, y_, z_;
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
  // rotation matrix -> quaternion by the branch on the trace / largest diagonal entry that Eigen documents (w >= 0 on the first branch)
  explicit Quaternion(const Matrix<T, 3, 3>& m) {
    T t = m.trace();
    if (t > T(0)) {
      t = std::sqrt(t + T(1));
      w_ = T(0.5) * t; t = T(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t; y_ = (m(0, 2) - m(2, 0)) * t; z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1));
      T v[3];
      v[i] = T(0.5) * t; t = T(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t; v[j] = (m(j, i) + m(i, j)) * t; v[k] = (m(k, i) + m(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  T& w() { return w_; } T& x() { return x_; } T& y() { return y_; } T& z() { return z_; }
  const T& w() const { return w_; } const T& x() const { return x_; } const T& y() const { return y_; } const T& z() const { return z_; }
  T norm() const { return std::sqrt(w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_); }
  void normalize() { const T n = norm(); w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& p) const { return toRotationMatrix() * p; }
  Quaternion operator*(const Quaternion& o) const {
    return Quaternion(w_ * o.w_ - x_ * o.x_ - y_ * o.y_ - z_ * o.z_, w_ * o.x_ + x_ * o.w_ + y_ * o.z_ - z_ * o.y_,
                      w_ * o.y_ - x_ * o.z_ + y_ * o.w_ + z_ * o.x_, w_ * o.z_ + x_ * o.y_ - y_ * o.x_ + z_ * o.w_);
  }
  Quaternion operator*(const AngleAxis<T>& o) const;
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> R;
    R(0, 0) = 1 - 2 * (y_ * y_ + z_ * z_); R(0, 1) = 2 * (x_ * y_ - w_ * z_); R(0, 2) = 2 * (x_ * z_ + w_ * y_);
    R(1, 0) = 2 * (x_ * y_ + w_ * z_); R(1, 1) = 1 - 2 * (x_ * x_ + z_ * z_); R(1, 2) = 2 * (y_ * z_ - w_ * x_);
    R(2, 0) = 2 * (x_ * z_ - w_ * y_); R(2, 1) = 2 * (y_ * z_ + w_ * x_); R(2, 2) = 1 - 2 * (x_ * x_ + y_ * y_);
    return R;
  }
};
template <typename T>
class AngleAxis {
  T ang_;
  Matrix<T, 3, 1> ax_;

 public:
  AngleAxis() : ang_(0), ax_(1, 0, 0) {}
  AngleAxis(T ang, const Matrix<T, 3, 1>& ax) : ang_(ang), ax_(ax) {}
  explicit AngleAxis(const Matrix<T, 3, 3>& R) {
    const T c = std::min(T(1), std::max(T(-1), (R.trace() - T(1)) / T(2)));
    ang_ = std::acos(c);
    Matrix<T, 3, 1> k(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
    const T n = k.norm();
    ax_ = n > T(0) ? Matrix<T, 3, 1>(k / n) : Matrix<T, 3, 1>(1, 0, 0);
  }
  T angle() const { return ang_; }
  const Matrix<T, 3, 1>& axis() const { return ax_; }
  Quaternion<T> quat() const { const T h = ang_ / 2, sn = std::sin(h); return Quaternion<T>(std::cos(h), sn * ax_[0], sn * ax_[1], sn * ax_[2]); }
  Quaternion<T> operator*(const AngleAxis& o) const { return quat() * o.quat(); }
  Matrix<T, 3, 3> toRotationMatrix() const { return quat().toRotationMatrix(); }
};
template <typename T>
Quaternion<T> Quaternion<T>::operator*(const AngleAxis<T>& o) const { return *this * o.quat(); }
typedef AngleAxis<double> AngleAxisd;
typedef Quaternion<double> Quaterniond;

// ---------------------------------------------------------------- sparse stand-ins (bavoxel.hpp:695-710)
template <typename T>
class Triplet {
  int r_, c_;
  T v_;

 public:
  Triplet() : r_(0), c_(0), v_(0) {}
  Triplet(int r, int c, const T& v) : r_(r), c_(c), v_(v) {}
  int row() const { return r_; }
  int col() const { return c_; }
  const T& value() const { return v_; }
};

template <typename T>
class SparseMatrix {
 public:
  int n_rows, n_cols;
  std::vector<Triplet<T>> entries;
  SparseMatrix() : n_rows(0), n_cols(0) {}
  SparseMatrix(int r, int c) : n_rows(r), n_cols(c) {}
  int rows() const { return n_rows; }
  int cols() const { return n_cols; }
  template <typename It>
  void setFromTriplets(It b, It e) { entries.assign(b, e); }      // the path never passes a duplicate (row, col)
  void makeCompressed() {}
};

// LDL^T of the LOWER triangle (Eigen's default UpLo), natural ordering, envelope storage, no pivoting.
template <typename SM>
class SimplicialLDLT {
  int n_ = 0;
  std::vector<int> first_;              // first column of row i inside the envelope
  std::vector<size_t> off_;             // start of row i in l_
  std::vector<double> l_, d_;

 public:
  SimplicialLDLT() {}
  void compute(const SM& A) {
    n_ = A.rows();
    first_.resize(n_);
    for (int i = 0; i < n_; ++i) first_[i] = i;
    for (const auto& t : A.entries) if (t.row() >= t.col()) first_[t.row()] = std::min(first_[t.row()], t.col());
    off_.assign(n_ + 1, 0);
    for (int i = 0; i < n_; ++i) off_[i + 1] = off_[i] + (size_t)(i - first_[i]);
    l_.assign(off_[n_], 0.0);
    d_.assign(n_, 0.0);
    for (const auto& t : A.entries) {
      if (t.row() > t.col()) l_[off_[t.row()] + (t.col() - first_[t.row()])] = t.value();
      else if (t.row() == t.col()) d_[t.row()] = t.value();
    }
    for (int i = 0; i < n_; ++i) {
      double* li = &l_[off_[i]] - first_[i];                 // li[j] = L(i, j) for first_[i] <= j < i
      for (int j = first_[i]; j < i; ++j) {
        const double* lj = &l_[off_[j]] - first_[j];
        double t = li[j];
        for (int k = std::max(first_[i], first_[j]); k < j; ++k) t -= li[k] * lj[k];   // li[k] still holds L(i,k) * d_k
        li[j] = t;
      }
      double dii = d_[i];
      for (int j = first_[i]; j < i; ++j) { const double w = li[j]; li[j] = w / d_[j]; dii -= w * li[j]; }
      d_[i] = dii;
    }
  }
  template <int R, int C>
  Matrix<double, R, C> solve(const Matrix<double, R, C>& b) const {
    Matrix<double, R, C> x(b);
    for (int i = 0; i < n_; ++i) {
      const double* li = &l_[off_[i]] - first_[i];
      double t = x(i);
      for (int j = first_[i]; j < i; ++j) t -= li[j] * x(j);
      x(i) = t;
    }
    for (int i = 0; i < n_; ++i) x(i) /= d_[i];
    for (int i = n_ - 1; i >= 0; --i) {
      const double* li = &l_[off_[i]] - first_[i];
      const double xi = x(i);
      for (int j = first_[i]; j < i; ++j) x(j) -= li[j] * xi;
    }
    return x;
  }
};

}  // namespace Eigen
