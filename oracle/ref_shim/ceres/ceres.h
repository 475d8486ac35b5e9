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
#include <functional>
#include <map>
#include <string>
#include <vector>
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
class LossFunction { public: virtual ~LossFunction() {} };
class HuberLoss : public LossFunction { public: double a; explicit HuberLoss(double a_) : a(a_) {} };
class CauchyLoss : public LossFunction { public: double a; explicit CauchyLoss(double a_) : a(a_) {} };
class Manifold { public: virtual ~Manifold() {} };
class EigenQuaternionManifold : public Manifold {};
class QuaternionManifold : public Manifold {};
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
// ceres::Problem here RECORDS what the reference hands to it (parameter blocks with their manifolds and constancy, residual blocks
// with cost functor, loss and parameter pointers); ceres::Solve passes the record and the options to a hook the oracle's driver
// installs.  There is no solver in this file.
class Problem {
 public:
  struct ParamBlock { double* p; int size; Manifold* manifold; bool constant; };
  struct ResidualBlock { CostFunction* cost; LossFunction* loss; std::vector<double*> params; };
  std::vector<ParamBlock> params;
  std::map<double*, size_t> index;
  std::vector<ResidualBlock> residuals;
  ~Problem() { for (auto& r : residuals) delete r.cost; for (auto& b : params) delete b.manifold; }
  void AddParameterBlock(double* p, int size, Manifold* m = nullptr) {
    auto it = index.find(p);
    if (it == index.end()) { index[p] = params.size(); params.push_back({p, size, m, false}); }
    else if (m != nullptr) params[it->second].manifold = m;
  }
  void SetParameterBlockConstant(double* p) { params[index.at(p)].constant = true; }
  void SetManifold(double* p, Manifold* m) { params[index.at(p)].manifold = m; }
  template <typename... Ps>
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, Ps... ps) { residuals.push_back({cost, loss, {ps...}}); }
  int NumResidualBlocks() const { return (int)residuals.size(); }
  int NumParameterBlocks() const { return (int)params.size(); }
};
struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double max_solver_time_in_seconds = 1e9;
  };
  struct Summary {
    TerminationType termination_type = NO_CONVERGENCE;
    double initial_cost = 0, final_cost = 0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0;
    std::vector<int> iterations;
    std::string message;
    std::string BriefReport() const { return "stand-in: see the driver's hook"; }
    std::string FullReport() const { return BriefReport(); }
    bool IsSolutionUsable() const { return termination_type != FAILURE; }
  };
};
inline std::function<void(const Solver::Options&, Problem*, Solver::Summary*)>& solve_hook() {
  static std::function<void(const Solver::Options&, Problem*, Solver::Summary*)> h;
  return h;
}
inline void Solve(const Solver::Options& o, Problem* p, Solver::Summary* s) { if (solve_hook()) solve_hook()(o, p, s); }
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
