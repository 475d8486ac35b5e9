// STAND-IN for the slice of ceres-solver include/utils.hpp names (NOT Ceres; test infrastructure, see ../mini_eigen.h).
// ceres-solver 2.1.0 is a third-party dependency of the reference (README.md:21, CMakeLists.txt:33), neither vendored
// under /root/reference nor installed here.  This header lets the reference's two cost functors
// (include/utils.hpp:51-147) be compiled from their own source and evaluated
//   * with T = double  (the residuals), and
//   * with T = ceres::Jet<double, N>  (the Jacobians auto-diff hands to the solver),
// which is exactly how ceres::AutoDiffCostFunction uses them.  Jet below is first-order forward-mode dual numbers written
// from the published definition (value + N partials, product / quotient / chain rule); the trust-region solver itself
// is NOT here — it stays a restatement (oracle/visual_oracle.py).
#pragma once
#include <cmath>
namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(T(0)) { for (int i = 0; i < N; ++i) v[i] = T(0); }
  Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(0); }     // NOLINT: implicit on purpose, as in Ceres
  Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(0); v[k] = T(1); }
  Jet& operator+=(const Jet& o) { a += o.a; for (int i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
  Jet& operator-=(const Jet& o) { a -= o.a; for (int i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
};
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h(f); h += g; return h; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h(f); h -= g; return h; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <typename T, int N> Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h;
}
template <typename T, int N> Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // (f / g)' = (f' - (f / g) g') / g, the form Ceres documents
  const T gi = T(1) / g.a, q = f.a * gi;
  Jet<T, N> h; h.a = q; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h;
}
// Jet (op) scalar, scalar (op) Jet
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h(f); h.a += s; return h; }
template <typename T, int N> Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h(f); h.a += s; return h; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h(f); h.a -= s; return h; }
template <typename T, int N> Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h(-f); h.a += s; return h; }
template <typename T, int N> Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
template <typename T, int N> Jet<T, N> operator/(const Jet<T, N>& f, T s) { return f * (T(1) / s); }
template <typename T, int N> Jet<T, N> operator/(T s, const Jet<T, N>& g) {
  const T gi = T(1) / g.a, q = s * gi;
  Jet<T, N> h; h.a = q; for (int i = 0; i < N; ++i) h.v[i] = -q * g.v[i] * gi; return h;
}
// comparisons look at the value only
#define LVBA_JET_CMP(op)                                                                                   \
  template <typename T, int N> bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  template <typename T, int N> bool operator op(const Jet<T, N>& f, T s) { return f.a op s; }                  \
  template <typename T, int N> bool operator op(T s, const Jet<T, N>& g) { return s op g.a; }
LVBA_JET_CMP(<) LVBA_JET_CMP(<=) LVBA_JET_CMP(>) LVBA_JET_CMP(>=) LVBA_JET_CMP(==) LVBA_JET_CMP(!=)
#undef LVBA_JET_CMP

inline double sqrt(double x) { return std::sqrt(x); }
template <typename T, int N> Jet<T, N> sqrt(const Jet<T, N>& f) {
  const T r = std::sqrt(f.a), two_r_inv = T(1) / (T(2) * r);
  Jet<T, N> h; h.a = r; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_r_inv; return h;
}

class CostFunction {
 public:
  virtual ~CostFunction() {}
};
// Sized like ceres::AutoDiffCostFunction<Functor, kNumResiduals, N0, N1, ...>; evaluation lives in the oracle's driver
// (oracle/ref_driver.cpp), which seeds the Jets the same way Ceres does (one unit partial per parameter coordinate).
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  Functor* f_;
 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) {}
  ~AutoDiffCostFunction() override { delete f_; }
  const Functor& functor() const { return *f_; }
};

}  // namespace ceres
