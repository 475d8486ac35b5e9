// cpu_ref.cpp — "reference-restated CPU path" (C++17, std::thread) of both hot paths.
//
// TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.  It is the second CPU oracle (independent of the numpy
// one: own Jacobi eigen-solver, own block LDL^T) and the CPU baseline that bench.py times beside the GPU
// (`cpu_baseline`, and `--impl reference`).  Nothing under global-lvba_b200/ links or calls it.
//
// PARITY: path A of this port is held against the reference's own source (tests/test_ref_pin.py: H, g, residual, damping_iter end
// poses vs tests/golden/ref_balm.npz; tools/ref_scale_check.py at configs B and C -> profiles/r02_ref_pin_scale_B.txt, _C.txt), the two Ceres
// cost functors likewise through the numpy oracle; the Ceres trust-region loop of path B stays a restatement of the published 2.1.0
// algorithm (no Ceres here).  The reference as a whole (ROS node, Eigen, Ceres, PCL) cannot be built here, so the
// timed CPU arm is this port, not the reference binary.  It keeps
// the reference's arithmetic and threading model and replaces only what cannot exist at the named sizes:
//   * dense per-voxel vector<PointCluster>(win_size) and dense 6W x 6W Hessians  ->  CSR slots and a
//     block-envelope (skyline) Hessian (the literal layout needs 41 GB + 22 GB at config C);
//   * Eigen::SelfAdjointEigenSolver<Matrix3d>  ->  cyclic Jacobi;   Eigen::SimplicialLDLT  ->  block LDL^T
//     without pivoting on the same lower triangle (bavoxel.hpp:695-710);
//   * Ceres DENSE_SCHUR  ->  the same Schur complement, stored block-sparse (a dense 12k x 12k Cholesky
//     would make this baseline far slower; the sparse solve is the generous choice).
// Threading as in the reference: Hessian / Jacobian build split over n_threads contiguous slices with
// private accumulators summed serially in thread order (bavoxel.hpp:614-633; Ceres num_threads,
// src/lvba_system.cpp:1575); residual-only pass and factorisation single-threaded (bavoxel.hpp:641-648, 706-710).
//
// Restated functions (reference file:line):
//   cluster_transform      PointCluster::transform                 include/BALM/tools.hpp:450-456
//   lidar_acc_evaluate2    VOX_HESS::acc_evaluate2                 include/BALM/bavoxel.hpp:68-174
//   lidar_only_residual    VOX_HESS::evaluate_only_residual        include/BALM/bavoxel.hpp:176-203
//   so3_exp                Exp                                     include/BALM/tools.hpp:62-77
//   ref_lidar_lm           BALM2::damping_iter + divide_thread     include/BALM/bavoxel.hpp:597-639, 662-767
//   reproj_eval            ReprojErrorWhitenedDistorted            include/utils.hpp:61-111
//   plane_eval             PointPlaneErrorWhitened                 include/utils.hpp:133-139
//   ref_visual_lm          Ceres block of optimizeCameraPoses      src/lvba_system.cpp:1571-1656
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

using std::vector;
double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ tiny dense helpers (row-major)
inline void mm3(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline void mm3bt(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
inline void mv3(const double* A, const double* x, double* y) { for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2]; }
inline void mtv3(const double* A, const double* x, double* y) { for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2]; }
inline void hat(const double* v, double* H) { H[0] = 0; H[1] = -v[2]; H[2] = v[1]; H[3] = v[2]; H[4] = 0; H[5] = -v[0]; H[6] = -v[1]; H[7] = v[0]; H[8] = 0; }
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// symmetric 3x3 eigen-decomposition, ascending, columns of U as u[k][.]
void eig3(const double* Cs /*xx xy xz yy yz zz*/, double lam[3], double u[3][3]) {
  double a[3][3] = {{Cs[0], Cs[1], Cs[2]}, {Cs[1], Cs[3], Cs[4]}, {Cs[2], Cs[4], Cs[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rot = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0 || std::fabs(a[p][q]) <= 1e-22 * (std::fabs(a[p][p]) + std::fabs(a[q][q]))) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = std::copysign(1.0, theta) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        const int r = 3 - p - q;
        a[p][p] -= t * a[p][q]; a[q][q] += t * a[p][q]; a[p][q] = a[q][p] = 0.0;
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = c * arp - s * arq;
        a[r][q] = a[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) { const double vp = v[k][p], vq = v[k][q]; v[k][p] = c * vp - s * vq; v[k][q] = s * vp + c * vq; }
        ++rot;
      }
    if (!rot) break;
  }
  int idx[3] = {0, 1, 2};
  std::sort(idx, idx + 3, [&](int x, int y) { return a[x][x] < a[y][y]; });
  for (int k = 0; k < 3; ++k) { lam[k] = a[idx[k]][idx[k]]; for (int i = 0; i < 3; ++i) u[k][i] = v[i][idx[k]]; }
}

void so3_exp(const double* w, double* R) {          // tools.hpp:62-77
  const double th = std::sqrt(dot3(w, w));
  if (th >= 1e-11) {
    const double a[3] = {w[0] / th, w[1] / th, w[2] / th};
    double K[9], KK[9];
    hat(a, K); mm3(K, K, KK);
    const double s = std::sin(th), oc = 1.0 - std::cos(th);
    for (int i = 0; i < 9; ++i) R[i] = s * K[i] + oc * KK[i];
    R[0] += 1; R[4] += 1; R[8] += 1;
  } else { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0); }
}

// ------------------------------------------------------------------ block envelope + LDL^T
struct Env {
  int n = 0;
  vector<int> first, last;
  vector<long long> rs;
  long long nb = 0;
  void build(vector<int> f) {
    n = (int)f.size(); first = f;
    for (int r = 0; r < n; ++r) first[r] = std::min(first[r], r);
    for (int r = n - 2; r >= 0; --r) first[r] = std::min(first[r], first[r + 1]);
    rs.assign(n + 1, 0);
    for (int r = 0; r < n; ++r) rs[r + 1] = rs[r] + (r - first[r] + 1);
    nb = rs[n]; last.assign(n, 0);
    int i = 0;
    for (int k = 0; k < n; ++k) { if (i < k) i = k; while (i + 1 < n && first[i + 1] <= k) ++i; last[k] = i; }
  }
  long long blk(int r, int c) const { return rs[r] + (c - first[r]); }
};

// 6x6 inverse through Gauss-Jordan without pivoting on the lower-mirrored block (same pivots as LDL^T)
bool inv6(const double* A, double* K) {
  double x[36];
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) x[6 * r + c] = (r >= c) ? A[6 * r + c] : A[6 * c + r];
  for (int p = 0; p < 6; ++p) {
    const double ip = 1.0 / x[7 * p];
    double n[36];
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) {
      if (r == p && c == p) n[6 * r + c] = ip;
      else if (r == p) n[6 * r + c] = x[6 * p + c] * ip;
      else if (c == p) n[6 * r + c] = -x[6 * r + p] * ip;
      else n[6 * r + c] = x[6 * r + c] - x[6 * r + p] * x[6 * p + c] * ip;
    }
    std::memcpy(x, n, sizeof x);
  }
  bool ok = true;
  for (int i = 0; i < 36; ++i) { K[i] = x[i]; ok = ok && std::isfinite(x[i]); }
  return ok;
}

// solves (H + diag(dadd)) x = b in place of L (copy of H); single-threaded like SimplicialLDLT
bool env_solve(const Env& e, vector<double>& L, const double* dadd, const double* b, double* x) {
  const int n = e.n;
  for (int r = 0; r < n; ++r) for (int a = 0; a < 6; ++a) L[e.blk(r, r) * 36 + 7 * a] += dadd[6 * r + a];
  vector<double> dinv((size_t)n * 36), z(b, b + 6 * (size_t)n), T;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    double* K = &dinv[(size_t)k * 36];
    ok = inv6(&L[e.blk(k, k) * 36], K) && ok;
    const int m = e.last[k] - k;
    T.resize((size_t)m * 36);
    for (int i = 0; i < m; ++i) {
      double* A = &L[e.blk(k + 1 + i, k) * 36];
      std::memcpy(&T[(size_t)i * 36], A, 36 * sizeof(double));
      double Ln[36];
      for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) { double s = 0; for (int q = 0; q < 6; ++q) s += A[6 * a + q] * K[6 * q + c]; Ln[6 * a + c] = s; }
      std::memcpy(A, Ln, sizeof Ln);
      for (int a = 0; a < 6; ++a) { double s = 0; for (int c = 0; c < 6; ++c) s += Ln[6 * a + c] * z[6 * k + c]; z[6 * (k + 1 + i) + a] -= s; }
    }
    for (int i = 0; i < m; ++i) {
      const double* Li = &L[e.blk(k + 1 + i, k) * 36];
      for (int j = 0; j <= i; ++j) {
        const double* Tj = &T[(size_t)j * 36];
        double* C = &L[e.blk(k + 1 + i, k + 1 + j) * 36];
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) {
          double s = 0; for (int q = 0; q < 6; ++q) s += Li[6 * a + q] * Tj[6 * c + q];
          C[6 * a + c] -= s;
        }
      }
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    const double* K = &dinv[(size_t)k * 36];
    double w[6];
    for (int a = 0; a < 6; ++a) { double s = 0; for (int c = 0; c < 6; ++c) s += K[6 * a + c] * z[6 * k + c]; w[a] = s; }
    for (int i = k + 1; i <= e.last[k]; ++i) {
      const double* B = &L[e.blk(i, k) * 36];
      for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) w[c] -= B[6 * a + c] * x[6 * i + a];
    }
    for (int a = 0; a < 6; ++a) { x[6 * k + a] = w[a]; ok = ok && std::isfinite(w[a]); }
  }
  return ok;
}

// ================================================================== path A
struct Lidar {
  int W; long long V;
  const int64_t* vp; const int32_t* pi; const double* cl;
  Env env;
};

inline void cluster_transform(const double* c, const double* pose, double* out /*P6 v3 N*/) {   // tools.hpp:450-456
  const double P[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
  const double* R = pose; const double* t = pose + 9; const double N = c[9];
  double RP[9], RPRt[9], Rv[3];
  mm3(R, P, RP); mm3bt(RP, R, RPRt); mv3(R, c + 6, Rv);
  out[0] = RPRt[0] + 2 * Rv[0] * t[0] + N * t[0] * t[0];
  out[1] = RPRt[1] + Rv[0] * t[1] + Rv[1] * t[0] + N * t[0] * t[1];
  out[2] = RPRt[2] + Rv[0] * t[2] + Rv[2] * t[0] + N * t[0] * t[2];
  out[3] = RPRt[4] + 2 * Rv[1] * t[1] + N * t[1] * t[1];
  out[4] = RPRt[5] + Rv[1] * t[2] + Rv[2] * t[1] + N * t[1] * t[2];
  out[5] = RPRt[8] + 2 * Rv[2] * t[2] + N * t[2] * t[2];
  out[6] = Rv[0] + N * t[0]; out[7] = Rv[1] + N * t[1]; out[8] = Rv[2] + N * t[2]; out[9] = N;
}

inline void voxel_cov(const Lidar& L, const double* poses, long long a, double cov[6], double vbar[3], double& Nsum) {
  double acc[10] = {0};
  for (int64_t s = L.vp[a]; s < L.vp[a + 1]; ++s) {
    double o[10];
    cluster_transform(L.cl + 10 * s, poses + 12 * (size_t)L.pi[s], o);
    for (int q = 0; q < 10; ++q) acc[q] += o[q];
  }
  Nsum = acc[9];
  for (int i = 0; i < 3; ++i) vbar[i] = acc[6 + i] / Nsum;
  cov[0] = acc[0] / Nsum - vbar[0] * vbar[0]; cov[1] = acc[1] / Nsum - vbar[0] * vbar[1]; cov[2] = acc[2] / Nsum - vbar[0] * vbar[2];
  cov[3] = acc[3] / Nsum - vbar[1] * vbar[1]; cov[4] = acc[4] / Nsum - vbar[1] * vbar[2]; cov[5] = acc[5] / Nsum - vbar[2] * vbar[2];
}

double lidar_only_residual(const Lidar& L, const double* poses) {    // bavoxel.hpp:176-203 (single thread)
  double res = 0;
  for (long long a = 0; a < L.V; ++a) {
    double cov[6], vbar[3], N, lam[3], u[3][3];
    voxel_cov(L, poses, a, cov, vbar, N);
    eig3(cov, lam, u);
    res += lam[0];
  }
  return res;
}

// bavoxel.hpp:68-174 on the voxel slice [head,end): H (envelope, lower), g, residual
void lidar_acc_evaluate2(const Lidar& L, const double* poses, long long head, long long end, double* H, double* g, double& residual) {
  residual = 0;
  vector<double> Auk, vr;   // per slot: Auk 18, viRiTuk 3
  for (long long a = head; a < end; ++a) {
    const int64_t s0 = L.vp[a], K = L.vp[a + 1] - s0;
    double cov[6], vbar[3], Nsum, lam[3], u[3][3];
    voxel_cov(L, poses, a, cov, vbar, Nsum);
    eig3(cov, lam, u);
    const double NN = (double)(int)Nsum;
    const double* uk = u[0];
    double umumT[9] = {0}, ukukT[9];
    for (int m = 1; m < 3; ++m) { const double w = 2.0 / (lam[0] - lam[m]); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) umumT[3 * i + j] += w * u[m][i] * u[m][j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ukukT[3 * i + j] = uk[i] * uk[j];
    Auk.resize((size_t)K * 18); vr.resize((size_t)K * 3);
    for (int64_t q = 0; q < K; ++q) {
      const double* c = L.cl + 10 * (s0 + q);
      const int pose = L.pi[s0 + q];
      const double* R = poses + 12 * (size_t)pose; const double* t = R + 9;
      const double Pi[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
      const double* vi = c + 6; const double ni = c[9];
      double vihat[9], RiTuk[3], RiTukhat[9], PiRiTuk[3], viRiTuk[3], tiv[3];
      hat(vi, vihat); mtv3(R, uk, RiTuk); hat(RiTuk, RiTukhat); mv3(Pi, RiTuk, PiRiTuk); mv3(vihat, RiTuk, viRiTuk);
      for (int i = 0; i < 3; ++i) { tiv[i] = t[i] - vbar[i]; vr[3 * q + i] = viRiTuk[i]; }
      const double ukTtiv = dot3(uk, tiv);
      double combo1[9], hp[9], combo2[3], Rvi[3];
      hat(PiRiTuk, hp);
      for (int i = 0; i < 9; ++i) combo1[i] = hp[i] + vihat[i] * ukTtiv;
      mv3(R, vi, Rvi);
      for (int i = 0; i < 3; ++i) combo2[i] = Rvi[i] + ni * tiv[i];
      double M1[9], A0[9], Rc1[9];
      mm3(R, Pi, M1);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M1[3 * i + j] += tiv[i] * vi[j];
      mm3(M1, RiTukhat, A0); mm3(R, combo1, Rc1);
      double* A = &Auk[(size_t)q * 18];
      const double c2u = dot3(combo2, uk);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        A[6 * i + j] = (A0[3 * i + j] - Rc1[3 * i + j]) / NN;
        A[6 * i + 3 + j] = (combo2[i] * uk[j] + (i == j ? c2u : 0.0)) / NN;
      }
      double jjt[6];
      for (int j = 0; j < 6; ++j) { jjt[j] = A[j] * uk[0] + A[6 + j] * uk[1] + A[12 + j] * uk[2]; g[6 * (size_t)pose + j] += jjt[j]; }
      double UA[18];   // umumT * Auk
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) UA[6 * i + j] = umumT[3 * i] * A[j] + umumT[3 * i + 1] * A[6 + j] + umumT[3 * i + 2] * A[12 + j];
      double Hb[36];
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Hb[6 * i + j] = A[i] * UA[j] + A[6 + i] * UA[6 + j] + A[12 + i] * UA[12 + j];
      double D0[9], haP[9], E[9], hj[9];
      mm3(RiTukhat, Pi, haP);
      for (int i = 0; i < 9; ++i) D0[i] = combo1[i] - haP[i];
      mm3(D0, RiTukhat, E); hat(jjt, hj);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        Hb[6 * i + j] += 2.0 / NN * E[3 * i + j] - 2.0 / NN / NN * viRiTuk[i] * viRiTuk[j] - 0.5 * hj[3 * i + j];
        const double hrt = 2.0 / NN * (1.0 - ni / NN) * viRiTuk[i] * uk[j];
        Hb[6 * i + 3 + j] += hrt; Hb[6 * (3 + j) + i] += hrt;
        Hb[6 * (3 + i) + 3 + j] += 2.0 / NN * (ni - ni * ni / NN) * ukukT[3 * i + j];
      }
      double* dst = H + L.env.blk(pose, pose) * 36;
      for (int i = 0; i < 36; ++i) dst[i] += Hb[i];
    }
    for (int64_t qi = 0; qi < K - 1; ++qi) {
      const double ni = L.cl[10 * (s0 + qi) + 9];
      const double* Ai = &Auk[(size_t)qi * 18];
      double UAi[18];   // Auk_i^T umumT  -> (6x3) stored as [j][c]
      for (int j = 0; j < 6; ++j) for (int c = 0; c < 3; ++c) UAi[3 * j + c] = Ai[j] * umumT[c] + Ai[6 + j] * umumT[3 + c] + Ai[12 + j] * umumT[6 + c];
      for (int64_t qj = qi + 1; qj < K; ++qj) {
        const double nj = L.cl[10 * (s0 + qj) + 9];
        const double* Aj = &Auk[(size_t)qj * 18];
        double Hb[36];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Hb[6 * i + j] = UAi[3 * i] * Aj[j] + UAi[3 * i + 1] * Aj[6 + j] + UAi[3 * i + 2] * Aj[12 + j];
        const double* wi = &vr[3 * qi]; const double* wj = &vr[3 * qj];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
          Hb[6 * i + j] += -2.0 / NN / NN * wi[i] * wj[j];
          Hb[6 * i + 3 + j] += -2.0 * nj / NN / NN * wi[i] * uk[j];
          Hb[6 * (3 + i) + j] += -2.0 * ni / NN / NN * uk[i] * wj[j];
          Hb[6 * (3 + i) + 3 + j] += -2.0 * ni * nj / NN / NN * ukukT[3 * i + j];
        }
        // lower block (row = pose_j, col = pose_i) = Hb^T   (bavoxel.hpp:171-173)
        double* dst = H + L.env.blk(L.pi[s0 + qj], L.pi[s0 + qi]) * 36;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) dst[6 * j + i] += Hb[6 * i + j];
      }
    }
    residual += lam[0];
  }
}

// divide_thread (bavoxel.hpp:597-639): private accumulators per thread, serial sum in thread order
double lidar_divide_thread(const Lidar& L, const double* poses, int nthreads, vector<double>& H, vector<double>& g, vector<vector<double>>& scratchH) {
  const size_t hs = (size_t)L.env.nb * 36, gs = (size_t)L.W * 6;
  int tn = nthreads;
  if (L.V < tn) tn = 1;
  if ((int)scratchH.size() < tn) scratchH.resize(tn);
  vector<vector<double>> gt(tn, vector<double>(gs, 0.0));
  vector<double> res(tn, 0.0);
  vector<std::thread> th;
  const double part = 1.0 * (double)L.V / tn;
  for (int i = 0; i < tn; ++i) {
    scratchH[i].assign(hs, 0.0);
    th.emplace_back([&, i] { lidar_acc_evaluate2(L, poses, (long long)(part * i), (long long)(part * (i + 1)), scratchH[i].data(), gt[i].data(), res[i]); });
  }
  H.assign(hs, 0.0); g.assign(gs, 0.0);
  double residual = 0;
  for (int i = 0; i < tn; ++i) {
    th[i].join();
    for (size_t q = 0; q < hs; ++q) H[q] += scratchH[i][q];
    for (size_t q = 0; q < gs; ++q) g[q] += gt[i][q];
    residual += res[i];
  }
  return residual;
}

void lidar_setup(Lidar& L, int W, long long V, const int64_t* vp, const int32_t* pi, const double* cl) {
  L.W = W; L.V = V; L.vp = vp; L.pi = pi; L.cl = cl;
  vector<int> f(W);
  for (int r = 0; r < W; ++r) f[r] = r;
  for (long long a = 0; a < V; ++a) { const int m = pi[vp[a]]; for (int64_t s = vp[a]; s < vp[a + 1]; ++s) f[pi[s]] = std::min(f[pi[s]], m); }
  L.env.build(f);
}

void lidar_retract(int W, const double* poses, const double* dx, double* out) {   // bavoxel.hpp:722-727
  for (int j = 0; j < W; ++j) {
    double E[9];
    so3_exp(dx + 6 * j, E);
    mm3(poses + 12 * j, E, out + 12 * j);
    for (int i = 0; i < 3; ++i) out[12 * j + 9 + i] = poses[12 * j + 9 + i] + dx[6 * j + 3 + i];
  }
}

// ================================================================== path B
struct Visual {
  int M; long long T;
  const double* plane; const int64_t* op; const int32_t* oc; const float* uv;
  double intr[8], isp, ispl;
  int fixed;
  vector<char> tv; vector<int> row_of_cam, cam_of_row; vector<long long> valid;
  Env env;
};

struct Obs { double r[2], Jc[12], JX[6]; };

void reproj_eval(const Visual& P, const double* qp, const double* tp, const double* X, const float* uv, bool jac, Obs& o) {  // utils.hpp:61-111
  double q0 = qp[0], q1 = qp[1], q2 = qp[2], q3 = qp[3];
  const double inv = 1.0 / std::sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 *= inv; q1 *= inv; q2 *= inv; q3 *= inv;
  const double v[3] = {q1, q2, q3};
  const double vxX[3] = {v[1] * X[2] - v[2] * X[1], v[2] * X[0] - v[0] * X[2], v[0] * X[1] - v[1] * X[0]};
  const double uvv[3] = {2 * vxX[0], 2 * vxX[1], 2 * vxX[2]};
  const double t2[3] = {v[1] * uvv[2] - v[2] * uvv[1], v[2] * uvv[0] - v[0] * uvv[2], v[0] * uvv[1] - v[1] * uvv[0]};
  const double Xc[3] = {X[0] + q0 * uvv[0] + t2[0] + tp[0], X[1] + q0 * uvv[1] + t2[1] + tp[1], X[2] + q0 * uvv[2] + t2[2] + tp[2]};
  std::memset(&o, 0, sizeof o);
  if (!(Xc[2] > 1e-8)) return;
  const double fx = P.intr[0], fy = P.intr[1], cx = P.intr[2], cy = P.intr[3], k1 = P.intr[4], k2 = P.intr[5], p1 = P.intr[6], p2 = P.intr[7];
  const double iz = 1.0 / Xc[2], xn = Xc[0] * iz, yn = Xc[1] * iz, r2 = xn * xn + yn * yn;
  const double rad = 1 + k1 * r2 + k2 * r2 * r2;
  const double xd = xn * rad + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn);
  const double yd = yn * rad + p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn;
  o.r[0] = (fx * xd + cx - (double)uv[0]) * P.isp;
  o.r[1] = (fy * yd + cy - (double)uv[1]) * P.isp;
  if (!jac) return;
  const double g = 2 * (k1 + 2 * k2 * r2);
  const double d00 = rad + xn * xn * g + 2 * p1 * yn + 6 * p2 * xn, d01 = xn * yn * g + 2 * p1 * xn + 2 * p2 * yn, d11 = rad + yn * yn * g + 6 * p1 * yn + 2 * p2 * xn;
  const double sx = fx * P.isp, sy = fy * P.isp;
  const double Jp[6] = {sx * d00 * iz, sx * d01 * iz, -sx * (d00 * xn + d01 * yn) * iz, sy * d01 * iz, sy * d11 * iz, -sy * (d01 * xn + d11 * yn) * iz};
  const double R[9] = {1 - 2 * (q2 * q2 + q3 * q3), 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2),
                       2 * (q1 * q2 + q0 * q3), 1 - 2 * (q1 * q1 + q3 * q3), 2 * (q2 * q3 - q0 * q1),
                       2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 1 - 2 * (q1 * q1 + q2 * q2)};
  for (int rho = 0; rho < 2; ++rho) for (int m = 0; m < 3; ++m) o.JX[3 * rho + m] = Jp[3 * rho] * R[m] + Jp[3 * rho + 1] * R[3 + m] + Jp[3 * rho + 2] * R[6 + m];
  double Ja[12], hX[9], hC[9];
  const double vX = dot3(v, X);
  hat(X, hX); hat(vxX, hC);
  for (int i = 0; i < 3; ++i) {
    Ja[4 * i] = 2 * vxX[i];
    for (int j = 0; j < 3; ++j) Ja[4 * i + 1 + j] = -2 * q0 * hX[3 * i + j] - 2 * hC[3 * i + j] - 2 * X[i] * v[j] + (i == j ? 2 * vX : 0.0);
  }
  const double PJ[12] = {q3, q2, -q1, -q2, q3, q0, q1, -q0, q3, -q0, -q1, -q2};   // EigenQuaternionManifold on (w,x,y,z) memory (Q9)
  double J3[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J3[3 * i + j] = Ja[4 * i] * PJ[j] + Ja[4 * i + 1] * PJ[3 + j] + Ja[4 * i + 2] * PJ[6 + j] + Ja[4 * i + 3] * PJ[9 + j];
  for (int rho = 0; rho < 2; ++rho) for (int j = 0; j < 3; ++j) {
    o.Jc[6 * rho + j] = Jp[3 * rho] * J3[j] + Jp[3 * rho + 1] * J3[3 + j] + Jp[3 * rho + 2] * J3[6 + j];
    o.Jc[6 * rho + 3 + j] = Jp[3 * rho + j];
  }
}

inline void plane_eval(const Visual& P, const double* pl, const double* X, double& r, double* J) {   // utils.hpp:133-139
  const double e = -(pl[0] * X[0] + pl[1] * X[1] + pl[2] * X[2] + pl[3]);
  const double root = std::sqrt(e * e + 1e-12);
  r = root * P.ispl;
  const double k = (e / root) * P.ispl;
  J[0] = -k * pl[0]; J[1] = -k * pl[1]; J[2] = -k * pl[2];
}

void visual_setup(Visual& P, int M, long long T, const double* plane, const int64_t* op, const int32_t* oc, const float* uv,
                  const double* intr, double sp, double spl, int fixed) {
  P.M = M; P.T = T; P.plane = plane; P.op = op; P.oc = oc; P.uv = uv; P.fixed = fixed;
  for (int i = 0; i < 8; ++i) P.intr[i] = intr[i];
  P.isp = 1.0 / sp; P.ispl = 1.0 / std::max(1e-9, spl);
  P.tv.assign(T, 0);
  vector<char> used(M, 0);
  for (long long i = 0; i < T; ++i) {
    const double* p = plane + 4 * i;
    bool ok = std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && std::isfinite(p[3]) &&
              (std::fabs(p[0]) > 1e-6 || std::fabs(p[1]) > 1e-6 || std::fabs(p[2]) > 1e-6);
    if (!ok) continue;
    P.tv[i] = 1; P.valid.push_back(i);
    for (int64_t s = op[i]; s < op[i + 1]; ++s) used[oc[s]] = 1;
  }
  if (fixed >= 0 && fixed < M) used[fixed] = 0;
  P.row_of_cam.assign(M, -1);
  for (int c = 0; c < M; ++c) if (used[c]) { P.row_of_cam[c] = (int)P.cam_of_row.size(); P.cam_of_row.push_back(c); }
  const int n = (int)P.cam_of_row.size();
  vector<int> f(n);
  for (int r = 0; r < n; ++r) f[r] = r;
  for (long long i : P.valid) {
    int m = INT32_MAX;
    for (int64_t s = op[i]; s < op[i + 1]; ++s) { const int r = P.row_of_cam[oc[s]]; if (r >= 0) m = std::min(m, r); }
    if (m == INT32_MAX) continue;
    for (int64_t s = op[i]; s < op[i + 1]; ++s) { const int r = P.row_of_cam[oc[s]]; if (r >= 0) f[r] = std::min(f[r], m); }
  }
  P.env.build(f);
}

double visual_cost(const Visual& P, const double* q, const double* t, const double* X) {
  double c = 0;
  for (long long i : P.valid) {
    for (int64_t s = P.op[i]; s < P.op[i + 1]; ++s) { Obs o; reproj_eval(P, q + 4 * P.oc[s], t + 3 * P.oc[s], X + 3 * i, P.uv + 2 * s, false, o); c += o.r[0] * o.r[0] + o.r[1] * o.r[1]; }
    double rp, J[3]; plane_eval(P, P.plane + 4 * i, X + 3 * i, rp, J); c += rp * rp;
  }
  return 0.5 * c;
}

struct VisualAcc { vector<double> S, rhs, colsq, grad; double cost = 0, gmax = 0; };

// linearise + Schur-eliminate the landmark slice [k0,k1) of P.valid into acc (private per thread)
void visual_build_slice(const Visual& P, const double* q, const double* t, const double* X, const double* s_cam, const double* s_pt,
                        double radius, double mind, double maxd, size_t k0, size_t k1, VisualAcc& acc) {
  vector<Obs> ob; vector<double> E, Y;
  for (size_t k = k0; k < k1; ++k) {
    const long long i = P.valid[k];
    const int64_t s0 = P.op[i], L = P.op[i + 1] - s0;
    ob.resize(L); E.assign((size_t)L * 18, 0.0); Y.assign((size_t)L * 18, 0.0);
    const double* sp = s_pt + 3 * i;
    double C[6] = {0}, gp[3] = {0};
    for (int64_t l = 0; l < L; ++l) {
      Obs& o = ob[l];
      const int cam = P.oc[s0 + l], row = P.row_of_cam[cam];
      reproj_eval(P, q + 4 * cam, t + 3 * cam, X + 3 * i, P.uv + 2 * (s0 + l), true, o);
      acc.cost += o.r[0] * o.r[0] + o.r[1] * o.r[1];
      for (int rho = 0; rho < 2; ++rho) for (int m = 0; m < 3; ++m) o.JX[3 * rho + m] *= sp[m];
      if (row >= 0) { for (int rho = 0; rho < 2; ++rho) for (int a = 0; a < 6; ++a) o.Jc[6 * rho + a] *= s_cam[6 * row + a]; }
      else std::memset(o.Jc, 0, sizeof o.Jc);
      C[0] += o.JX[0] * o.JX[0] + o.JX[3] * o.JX[3]; C[1] += o.JX[0] * o.JX[1] + o.JX[3] * o.JX[4]; C[2] += o.JX[0] * o.JX[2] + o.JX[3] * o.JX[5];
      C[3] += o.JX[1] * o.JX[1] + o.JX[4] * o.JX[4]; C[4] += o.JX[1] * o.JX[2] + o.JX[4] * o.JX[5]; C[5] += o.JX[2] * o.JX[2] + o.JX[5] * o.JX[5];
      for (int m = 0; m < 3; ++m) gp[m] += o.JX[m] * o.r[0] + o.JX[3 + m] * o.r[1];
    }
    double rp, Jp[3];
    plane_eval(P, P.plane + 4 * i, X + 3 * i, rp, Jp);
    acc.cost += rp * rp;
    for (int m = 0; m < 3; ++m) Jp[m] *= sp[m];
    C[0] += Jp[0] * Jp[0]; C[1] += Jp[0] * Jp[1]; C[2] += Jp[0] * Jp[2]; C[3] += Jp[1] * Jp[1]; C[4] += Jp[1] * Jp[2]; C[5] += Jp[2] * Jp[2];
    for (int m = 0; m < 3; ++m) { gp[m] += Jp[m] * rp; acc.gmax = std::max(acc.gmax, std::fabs(gp[m] / sp[m])); }
    C[0] += std::min(std::max(C[0], mind), maxd) / radius; C[3] += std::min(std::max(C[3], mind), maxd) / radius; C[5] += std::min(std::max(C[5], mind), maxd) / radius;
    const double A = C[3] * C[5] - C[4] * C[4], B = C[2] * C[4] - C[1] * C[5], Cc = C[1] * C[4] - C[2] * C[3];
    const double id = 1.0 / (C[0] * A + C[1] * B + C[2] * Cc);
    const double Ci[9] = {A * id, B * id, Cc * id, B * id, (C[0] * C[5] - C[2] * C[2]) * id, (C[1] * C[2] - C[0] * C[4]) * id,
                          Cc * id, (C[1] * C[2] - C[0] * C[4]) * id, (C[0] * C[3] - C[1] * C[1]) * id};
    double w[3]; mv3(Ci, gp, w);
    for (int64_t l = 0; l < L; ++l) {
      const Obs& o = ob[l];
      const int row = P.row_of_cam[P.oc[s0 + l]];
      if (row < 0) continue;
      double* e = &E[(size_t)l * 18]; double* y = &Y[(size_t)l * 18];
      for (int a = 0; a < 6; ++a) for (int m = 0; m < 3; ++m) e[3 * a + m] = o.Jc[a] * o.JX[m] + o.Jc[6 + a] * o.JX[3 + m];
      for (int a = 0; a < 6; ++a) for (int m = 0; m < 3; ++m) y[3 * a + m] = e[3 * a] * Ci[m] + e[3 * a + 1] * Ci[3 + m] + e[3 * a + 2] * Ci[6 + m];
      double* D = &acc.S[P.env.blk(row, row) * 36];
      for (int a = 0; a < 6; ++a) {
        const double gc = o.Jc[a] * o.r[0] + o.Jc[6 + a] * o.r[1];
        acc.rhs[6 * row + a] += -(gc - (e[3 * a] * w[0] + e[3 * a + 1] * w[1] + e[3 * a + 2] * w[2]));
        acc.colsq[6 * row + a] += o.Jc[a] * o.Jc[a] + o.Jc[6 + a] * o.Jc[6 + a];
        acc.grad[6 * row + a] += gc;
        for (int c = 0; c < 6; ++c) D[6 * a + c] += o.Jc[a] * o.Jc[c] + o.Jc[6 + a] * o.Jc[6 + c] - (y[3 * a] * e[3 * c] + y[3 * a + 1] * e[3 * c + 1] + y[3 * a + 2] * e[3 * c + 2]);
      }
    }
    for (int64_t l1 = 0; l1 < L; ++l1) for (int64_t l2 = l1 + 1; l2 < L; ++l2) {
      const int r1 = P.row_of_cam[P.oc[s0 + l1]], r2 = P.row_of_cam[P.oc[s0 + l2]];
      if (r1 < 0 || r2 < 0) continue;
      auto add = [&](int64_t hi, int64_t lo, int rh, int rl) {
        double* D = &acc.S[P.env.blk(rh, rl) * 36];
        const double* y = &Y[(size_t)hi * 18]; const double* e = &E[(size_t)lo * 18];
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) D[6 * a + c] -= y[3 * a] * e[3 * c] + y[3 * a + 1] * e[3 * c + 1] + y[3 * a + 2] * e[3 * c + 2];
      };
      if (r1 > r2) add(l1, l2, r1, r2); else if (r2 > r1) add(l2, l1, r2, r1); else { add(l1, l2, r1, r2); add(l2, l1, r2, r1); }
    }
  }
}

}  // namespace

// ====================================================================== C interface (ctypes)
extern "C" {

// out[0..]: iterations, accepted, builds, cost_first, cost_last, u_last, ms_total, ms_build, ms_solve, ms_resid
int ref_lidar_lm(int32_t W, int64_t V, const int64_t* vp, const int32_t* pi, const double* cl, double* poses,
                 double u0, double v0, int32_t max_iter, double rel_tol, int32_t nthreads, double* out) {
  Lidar L; lidar_setup(L, W, V, vp, pi, cl);
  const double t0 = now_ms();
  double u = u0, v = v0, residual1 = 0, residual2 = 0, q;
  bool is_calc_hess = true;
  vector<double> H, g, Lm, dx((size_t)6 * W), dadd((size_t)6 * W), mg((size_t)6 * W), trial((size_t)12 * W), cur(poses, poses + 12 * (size_t)W);
  vector<vector<double>> scratch;
  int iters = 0, acc = 0, builds = 0; double first = 0, lastc = 0, tb = 0, ts = 0, tr = 0;
  for (int i = 0; i < max_iter; ++i) {
    if (is_calc_hess) { const double a = now_ms(); residual1 = lidar_divide_thread(L, cur.data(), nthreads, H, g, scratch) / (double)V; tb += now_ms() - a; ++builds; if (i == 0) first = lastc = residual1; }
    const double a = now_ms();
    Lm = H;
    for (int r = 0; r < W; ++r) for (int c = 0; c < 6; ++c) { dadd[6 * r + c] = u * H[L.env.blk(r, r) * 36 + 7 * c]; mg[6 * r + c] = -g[6 * r + c]; }
    env_solve(L.env, Lm, dadd.data(), mg.data(), dx.data());
    ts += now_ms() - a;
    const double b = now_ms();
    lidar_retract(W, cur.data(), dx.data(), trial.data());
    double q1 = 0;
    for (int r = 0; r < 6 * W; ++r) q1 += dx[r] * (dadd[r] * dx[r] - g[r]);
    q1 = 0.5 * q1 / (double)V;
    residual2 = lidar_only_residual(L, trial.data()) / (double)V;
    tr += now_ms() - b;
    q = residual1 - residual2;
    ++iters;
    if (q > 0) { cur = trial; q = q / q1; v = 2; q = 1 - std::pow(2 * q - 1, 3); u *= (q < 1.0 / 3.0 ? 1.0 / 3.0 : q); is_calc_hess = true; ++acc; lastc = residual2; }
    else { u = u * v; v = 2 * v; is_calc_hess = false; }
    if (rel_tol >= 0 && std::fabs(residual1 - residual2) / residual1 < rel_tol) break;
  }
  std::memcpy(poses, cur.data(), sizeof(double) * 12 * (size_t)W);
  if (out) { out[0] = iters; out[1] = acc; out[2] = builds; out[3] = first; out[4] = lastc; out[5] = u; out[6] = now_ms() - t0; out[7] = tb; out[8] = ts; out[9] = tr; }
  return 0;
}

// one damped step at the given poses: (H + u diag(H)) dx = -g  (bavoxel.hpp:692-710); for the per-pose-update parity check
int ref_lidar_step(int32_t W, int64_t V, const int64_t* vp, const int32_t* pi, const double* cl, const double* poses, double u,
                   int32_t nthreads, double* dx_out, double* residual_sum) {
  Lidar L; lidar_setup(L, W, V, vp, pi, cl);
  vector<double> H, g, Lm, dadd((size_t)6 * W), mg((size_t)6 * W); vector<vector<double>> scratch;
  const double r = lidar_divide_thread(L, poses, nthreads, H, g, scratch);
  if (residual_sum) *residual_sum = r;
  Lm = H;
  for (int rr = 0; rr < W; ++rr) for (int c = 0; c < 6; ++c) { dadd[6 * rr + c] = u * H[L.env.blk(rr, rr) * 36 + 7 * c]; mg[6 * rr + c] = -g[6 * rr + c]; }
  env_solve(L.env, Lm, dadd.data(), mg.data(), dx_out);
  return 0;
}

// single phases for cross-checks against the numpy oracle
int ref_lidar_structure(int32_t W, int64_t V, const int64_t* vp, const int32_t* pi, int64_t* nblocks, int32_t* brow, int32_t* bcol) {
  Lidar L; lidar_setup(L, W, V, vp, pi, nullptr);
  *nblocks = L.env.nb;
  if (brow && bcol) for (int r = 0; r < W; ++r) for (int c = L.env.first[r]; c <= r; ++c) { brow[L.env.blk(r, c)] = r; bcol[L.env.blk(r, c)] = c; }
  return 0;
}
int ref_lidar_build(int32_t W, int64_t V, const int64_t* vp, const int32_t* pi, const double* cl, const double* poses, int32_t nthreads,
                    double* residual_sum, double* g_out, double* blocks_out) {
  Lidar L; lidar_setup(L, W, V, vp, pi, cl);
  vector<double> H, g; vector<vector<double>> scratch;
  *residual_sum = lidar_divide_thread(L, poses, nthreads, H, g, scratch);
  std::memcpy(g_out, g.data(), g.size() * sizeof(double));
  std::memcpy(blocks_out, H.data(), H.size() * sizeof(double));
  return 0;
}
double ref_lidar_residual(int32_t W, int64_t V, const int64_t* vp, const int32_t* pi, const double* cl, const double* poses) {
  Lidar L; lidar_setup(L, W, V, vp, pi, cl);
  return lidar_only_residual(L, poses);
}

// out: iterations, accepted, builds, cost_first, cost_last, radius, ms_total, ms_build, ms_solve, ms_resid, termination
int ref_visual_lm(int32_t M, int64_t T, double* q, double* t, double* X, const double* plane, const int64_t* op, const int32_t* oc,
                  const float* uv, const double* intr, double sp, double spl, int32_t fixed, int32_t max_iter, int32_t nthreads,
                  int32_t jacobi_scaling, double ftol, double* out) {
  Visual P; visual_setup(P, M, T, plane, op, oc, uv, intr, sp, spl, fixed);
  const double t0 = now_ms();
  const int n = P.env.n;
  const size_t nv = P.valid.size();
  vector<double> s_cam((size_t)6 * std::max(n, 1), 1.0), s_pt((size_t)3 * std::max<long long>(T, 1), 1.0);
  const double mind = 1e-6, maxd = 1e32;
  int tn = std::max(1, nthreads); if ((size_t)tn > nv) tn = 1;
  vector<VisualAcc> accs(tn);
  auto build = [&](double radius, VisualAcc& tot) {
    vector<std::thread> th;
    for (int i = 0; i < tn; ++i) {
      VisualAcc& a = accs[i];
      a.S.assign((size_t)P.env.nb * 36, 0.0); a.rhs.assign((size_t)6 * n, 0.0); a.colsq.assign((size_t)6 * n, 0.0); a.grad.assign((size_t)6 * n, 0.0); a.cost = 0; a.gmax = 0;
      const size_t k0 = nv * i / tn, k1 = nv * (i + 1) / tn;
      th.emplace_back([&, k0, k1, i] { visual_build_slice(P, q, t, X, s_cam.data(), s_pt.data(), radius, mind, maxd, k0, k1, accs[i]); });
    }
    tot.S.assign((size_t)P.env.nb * 36, 0.0); tot.rhs.assign((size_t)6 * n, 0.0); tot.colsq.assign((size_t)6 * n, 0.0); tot.grad.assign((size_t)6 * n, 0.0); tot.cost = 0; tot.gmax = 0;
    for (int i = 0; i < tn; ++i) {
      th[i].join();
      for (size_t k = 0; k < tot.S.size(); ++k) tot.S[k] += accs[i].S[k];
      for (size_t k = 0; k < tot.rhs.size(); ++k) { tot.rhs[k] += accs[i].rhs[k]; tot.colsq[k] += accs[i].colsq[k]; tot.grad[k] += accs[i].grad[k]; }
      tot.cost += accs[i].cost; tot.gmax = std::max(tot.gmax, accs[i].gmax);
    }
    tot.cost *= 0.5;
  };
  VisualAcc A;
  if (jacobi_scaling) {   // Jacobi scaling from the unscaled Jacobian at iteration 0
    build(1e300, A);      // radius -> infinity: S unused, only the column norms matter
    for (int r = 0; r < 6 * n; ++r) s_cam[r] = 1.0 / (1.0 + std::sqrt(A.colsq[r]));
    for (long long i : P.valid) {
      double c[3] = {0, 0, 0};
      for (int64_t s = op[i]; s < op[i + 1]; ++s) { Obs o; reproj_eval(P, q + 4 * oc[s], t + 3 * oc[s], X + 3 * i, uv + 2 * s, true, o); for (int m = 0; m < 3; ++m) c[m] += o.JX[m] * o.JX[m] + o.JX[3 + m] * o.JX[3 + m]; }
      double rp, J[3]; plane_eval(P, plane + 4 * i, X + 3 * i, rp, J);
      for (int m = 0; m < 3; ++m) s_pt[3 * i + m] = 1.0 / (1.0 + std::sqrt(c[m] + J[m] * J[m]));
    }
  }
  double radius = 1e4, nu = 2, cost = 0, first = 0, tb = 0, ts = 0, tr = 0;
  int iters = 0, accn = 0, builds = 0, invalid = 0, term = 0;
  vector<double> Lm, y((size_t)6 * std::max(n, 1)), dadd((size_t)6 * std::max(n, 1));
  vector<double> qc(q, q + 4 * (size_t)M), tc(t, t + 3 * (size_t)M), Xc(X, X + 3 * (size_t)T);
  for (int it = 0; it < max_iter; ++it) {
    double a = now_ms();
    build(radius, A); ++builds; tb += now_ms() - a;
    cost = A.cost; if (it == 0) first = cost;
    double gmax = A.gmax; for (int r = 0; r < 6 * n; ++r) gmax = std::max(gmax, std::fabs(A.grad[r] / s_cam[r]));
    if (gmax <= 1e-10) { term = 3; break; }
    ++iters;
    a = now_ms();
    for (int r = 0; r < 6 * n; ++r) dadd[r] = std::min(std::max(A.colsq[r], mind), maxd) / radius;
    Lm = A.S;
    const bool ok = env_solve(P.env, Lm, dadd.data(), A.rhs.data(), y.data());
    ts += now_ms() - a;
    a = now_ms();
    // back-substitution, model cost change, candidate
    double model = 0, step2 = 0, x2 = 0;
    for (long long i : P.valid) {
      const double* sp = &s_pt[3 * i];
      const int64_t s0 = op[i], L = op[i + 1] - s0;
      double C[6] = {0}, gp[3] = {0}, ety[3] = {0};
      vector<Obs> ob(L); vector<double> jcy((size_t)L * 2, 0.0);
      for (int64_t l = 0; l < L; ++l) {
        Obs& o = ob[l]; const int cam = oc[s0 + l], row = P.row_of_cam[cam];
        reproj_eval(P, q + 4 * cam, t + 3 * cam, X + 3 * i, uv + 2 * (s0 + l), true, o);
        for (int rho = 0; rho < 2; ++rho) for (int m = 0; m < 3; ++m) o.JX[3 * rho + m] *= sp[m];
        if (row >= 0) for (int rho = 0; rho < 2; ++rho) for (int c = 0; c < 6; ++c) { o.Jc[6 * rho + c] *= s_cam[6 * row + c]; jcy[2 * l + rho] += o.Jc[6 * rho + c] * y[6 * row + c]; }
        C[0] += o.JX[0] * o.JX[0] + o.JX[3] * o.JX[3]; C[1] += o.JX[0] * o.JX[1] + o.JX[3] * o.JX[4]; C[2] += o.JX[0] * o.JX[2] + o.JX[3] * o.JX[5];
        C[3] += o.JX[1] * o.JX[1] + o.JX[4] * o.JX[4]; C[4] += o.JX[1] * o.JX[2] + o.JX[4] * o.JX[5]; C[5] += o.JX[2] * o.JX[2] + o.JX[5] * o.JX[5];
        for (int m = 0; m < 3; ++m) { gp[m] += o.JX[m] * o.r[0] + o.JX[3 + m] * o.r[1]; ety[m] += o.JX[m] * jcy[2 * l] + o.JX[3 + m] * jcy[2 * l + 1]; }
      }
      double rp, Jp[3]; plane_eval(P, plane + 4 * i, X + 3 * i, rp, Jp);
      double Jps[3] = {Jp[0] * sp[0], Jp[1] * sp[1], Jp[2] * sp[2]};
      C[0] += Jps[0] * Jps[0]; C[1] += Jps[0] * Jps[1]; C[2] += Jps[0] * Jps[2]; C[3] += Jps[1] * Jps[1]; C[4] += Jps[1] * Jps[2]; C[5] += Jps[2] * Jps[2];
      for (int m = 0; m < 3; ++m) gp[m] += Jps[m] * rp;
      C[0] += std::min(std::max(C[0], mind), maxd) / radius; C[3] += std::min(std::max(C[3], mind), maxd) / radius; C[5] += std::min(std::max(C[5], mind), maxd) / radius;
      const double Aa = C[3] * C[5] - C[4] * C[4], B = C[2] * C[4] - C[1] * C[5], Cc = C[1] * C[4] - C[2] * C[3];
      const double id = 1.0 / (C[0] * Aa + C[1] * B + C[2] * Cc);
      const double Ci[9] = {Aa * id, B * id, Cc * id, B * id, (C[0] * C[5] - C[2] * C[2]) * id, (C[1] * C[2] - C[0] * C[4]) * id, Cc * id, (C[1] * C[2] - C[0] * C[4]) * id, (C[0] * C[3] - C[1] * C[1]) * id};
      const double rhsP[3] = {gp[0] + ety[0], gp[1] + ety[1], gp[2] + ety[2]};
      double yp[3]; mv3(Ci, rhsP, yp); for (int m = 0; m < 3; ++m) yp[m] = -yp[m];
      for (int64_t l = 0; l < L; ++l) for (int rho = 0; rho < 2; ++rho) {
        const double jy = jcy[2 * l + rho] + ob[l].JX[3 * rho] * yp[0] + ob[l].JX[3 * rho + 1] * yp[1] + ob[l].JX[3 * rho + 2] * yp[2];
        model -= jy * (ob[l].r[rho] + 0.5 * jy);
      }
      const double jy = Jps[0] * yp[0] + Jps[1] * yp[1] + Jps[2] * yp[2];
      model -= jy * (rp + 0.5 * jy);
      for (int m = 0; m < 3; ++m) { const double d = sp[m] * yp[m]; Xc[3 * i + m] = X[3 * i + m] + d; step2 += d * d; x2 += X[3 * i + m] * X[3 * i + m]; }
    }
    for (int r = 0; r < n; ++r) {
      const int c = P.cam_of_row[r];
      double d[6]; for (int k = 0; k < 6; ++k) d[k] = s_cam[6 * r + k] * y[6 * r + k];
      const double m0 = q[4 * c], m1 = q[4 * c + 1], m2 = q[4 * c + 2], m3 = q[4 * c + 3];
      const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      double n0 = m0, n1 = m1, n2 = m2, n3 = m3;
      if (nd > 0) {
        const double k = std::sin(nd) / nd, cs = std::cos(nd), s0 = k * d[0], s1 = k * d[1], s2 = k * d[2];
        n3 = cs * m3 - (s0 * m0 + s1 * m1 + s2 * m2);
        n0 = cs * m0 + m3 * s0 + (s1 * m2 - s2 * m1); n1 = cs * m1 + m3 * s1 + (s2 * m0 - s0 * m2); n2 = cs * m2 + m3 * s2 + (s0 * m1 - s1 * m0);
      }
      qc[4 * c] = n0; qc[4 * c + 1] = n1; qc[4 * c + 2] = n2; qc[4 * c + 3] = n3;
      for (int k = 0; k < 3; ++k) { tc[3 * c + k] = t[3 * c + k] + d[3 + k]; step2 += d[3 + k] * d[3 + k]; x2 += t[3 * c + k] * t[3 * c + k]; }
      step2 += (n0 - m0) * (n0 - m0) + (n1 - m1) * (n1 - m1) + (n2 - m2) * (n2 - m2) + (n3 - m3) * (n3 - m3);
      x2 += m0 * m0 + m1 * m1 + m2 * m2 + m3 * m3;
    }
    const double cand = visual_cost(P, qc.data(), tc.data(), Xc.data());
    tr += now_ms() - a;
    if (!ok || !std::isfinite(model) || !(model > 0)) { ++invalid; radius *= 0.5; if (invalid >= 5) { term = 5; break; } continue; }
    invalid = 0;
    const double rho = (cost - cand) / model, step_norm = std::sqrt(step2), x_norm = std::sqrt(x2);
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { term = 2; break; }
    if (ftol >= 0 && std::fabs(cost - cand) <= ftol * cost) { term = 1; break; }
    if (std::isfinite(cand) && rho > 1e-3) {
      std::memcpy(q, qc.data(), sizeof(double) * 4 * (size_t)M); std::memcpy(t, tc.data(), sizeof(double) * 3 * (size_t)M); std::memcpy(X, Xc.data(), sizeof(double) * 3 * (size_t)T);
      cost = cand; ++accn;
      radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2 * rho - 1, 3))); nu = 2;
    } else { radius /= nu; nu *= 2; if (radius < 1e-32) { term = 4; break; } }
  }
  if (out) { out[0] = iters; out[1] = accn; out[2] = builds; out[3] = first; out[4] = cost; out[5] = radius; out[6] = now_ms() - t0; out[7] = tb; out[8] = ts; out[9] = tr; out[10] = term; }
  return 0;
}

int ref_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
