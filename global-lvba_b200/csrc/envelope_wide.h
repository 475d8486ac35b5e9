// envelope_wide.h — block LDL^T of the pose system for envelopes of ANY width.
//
// The fast solvers (factor_la.cuh: columns of <= 30 blocks in registers; envelope.cuh: <= 320 blocks in shared memory)
// cover trajectories whose couplings stay inside a band.  One loop closure — a voxel seen from pose 10 and pose 1500 —
// makes the monotone envelope of every row in between reach back to column 10, and the reference's
// Eigen::SimplicialLDLT (include/BALM/bavoxel.hpp:704-707) still solves that system.  This is the general path that keeps
// the library a drop-in there: the same right-looking factorisation, one pivot column at a time, but with every column
// step spread over the whole device instead of one CTA:
//
//   for k = 0 .. n-1:   pivot   D_k^-1 (Gauss-Jordan without pivoting, the pivots of the scalar LDL^T the reference runs)
//                       scale   rows i below k:  T_i = A_ik,  L_ik = A_ik D_k^-1,  z_i -= L_ik z_k
//                       update  pairs i >= j below k:  A_ij -= L_ik T_j^T
//   w = D^-1 z ;  for i = n-1 .. 1:  w_k -= L_ik^T w_i  for every k in first[i] .. i-1          (x = w)
//
// Each pass is a functor over an index range (as in voxel_pipeline.h), launched by whatever `Launch` the caller passes:
// a grid-stride kernel on the device (runtime.cuh), a plain loop in the CPU test of the arithmetic (tests/emu/).
// 4 launches per block row: a fallback that is correct at any sparsity, not a fast path — DESIGN.md §4.1.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LVBA_WHD __host__ __device__ __forceinline__
#else
#define LVBA_WHD inline
#endif

namespace lvba {
namespace wide {

struct View {                 // same arrays as EnvView
  int n;
  const int* first;
  const long long* row_start;
};
LVBA_WHD long long block(const View& e, int r, int c) { return e.row_start[r] + (c - e.first[r]); }

// lower-triangular linear index t -> (a, b), b <= a
LVBA_WHD void tri_decode(int64_t t, int& a, int& b) {
  int64_t i = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  while (i * (i + 1) / 2 > t) --i;
  a = (int)i;
  b = (int)(t - i * (i + 1) / 2);
}

struct PivotF {               // one item
  View e; int k; double* L; double* dinv; int* status;
  LVBA_WHD void operator()(int64_t) const {
    const double* A = L + block(e, k, k) * 36;
    double K[36];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) K[6 * r + c] = (r >= c) ? A[6 * r + c] : A[6 * c + r];      // lower triangle mirrored
    for (int p = 0; p < 6; ++p) {                                                              // in-place Gauss-Jordan, no pivoting
      const double ip = 1.0 / K[7 * p];
      double row[6], col[6];
      for (int q = 0; q < 6; ++q) { row[q] = K[6 * p + q]; col[q] = K[6 * q + p]; }
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double v;
          if (r == p && c == p) v = ip;
          else if (r == p) v = row[c] * ip;
          else if (c == p) v = -col[r] * ip;
          else v = K[6 * r + c] - col[r] * row[c] * ip;
          K[6 * r + c] = v;
        }
    }
    double chk = 0.0;
    for (int q = 0; q < 36; ++q) { dinv[(long long)k * 36 + q] = K[q]; chk += K[q]; }
    if (!(fabs(chk) <= 1.79769313486231570e308)) status[0] = 1;
  }
};

struct ScaleF {               // items: rows k+1 .. k+n_k
  View e; int k; double* L; const double* dinv; double* z; double* colT;
  LVBA_WHD void operator()(int64_t r) const {
    const int i = k + 1 + (int)r;
    double* A = L + block(e, i, k) * 36;
    const double* K = dinv + (long long)k * 36;
    double T[36], Lk[36];
    for (int q = 0; q < 36; ++q) { T[q] = A[q]; colT[36 * r + q] = T[q]; }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) {
        double s = 0.0;
        for (int c = 0; c < 6; ++c) s += T[6 * a + c] * K[6 * c + b];
        Lk[6 * a + b] = s;
      }
    for (int q = 0; q < 36; ++q) A[q] = Lk[q];
    for (int a = 0; a < 6; ++a) {                          // forward substitution of the right-hand side
      double s = 0.0;
      for (int c = 0; c < 6; ++c) s += Lk[6 * a + c] * z[6 * k + c];
      z[6 * i + a] -= s;
    }
  }
};

struct UpdateF {              // items: pairs (a, b), b <= a, of the rows below k
  View e; int k; double* L; const double* colT;
  LVBA_WHD void operator()(int64_t t) const {
    int a, b;
    tri_decode(t, a, b);
    const int i = k + 1 + a, j = k + 1 + b;
    const double* Li = L + block(e, i, k) * 36;
    const double* Tj = colT + 36 * (long long)b;
    double* G = L + block(e, i, j) * 36;
    for (int x = 0; x < 6; ++x)
      for (int y = 0; y < 6; ++y) {
        double s = 0.0;
        for (int c = 0; c < 6; ++c) s += Li[6 * x + c] * Tj[6 * y + c];
        G[6 * x + y] -= s;
      }
  }
};

struct ApplyF {               // w = D^-1 z, items: 6 n scalars
  const double* dinv; const double* z; double* w;
  LVBA_WHD void operator()(int64_t i) const {
    const int64_t k = i / 6, r = i - 6 * k;
    double s = 0.0;
    for (int q = 0; q < 6; ++q) s += dinv[k * 36 + r * 6 + q] * z[6 * k + q];
    w[i] = s;
  }
};

struct BackF {                // items: columns first[i] .. i-1 of row i (w_i is final)
  View e; int i; const double* L; double* w;
  LVBA_WHD void operator()(int64_t r) const {
    const int k = e.first[i] + (int)r;
    const double* B = L + block(e, i, k) * 36;
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
      for (int a = 0; a < 6; ++a) s += B[6 * a + c] * w[6 * i + a];
      w[6 * k + c] -= s;
    }
  }
};

// h_first / h_last: host copies of the envelope arrays.  L enters holding H + damping; z the right-hand side; colT is a
// scratch of max_col * 36 doubles; x receives the solution.  launch(n_items, functor) runs one pass.
template <class Launch>
inline void factor_and_solve(Launch&& launch, const View& e, const int* h_first, const int* h_last, double* L, double* dinv, double* z,
                             double* colT, double* x, int* status) {
  for (int k = 0; k < e.n; ++k) {
    const int64_t nk = h_last[k] - k;
    launch((int64_t)1, PivotF{e, k, L, dinv, status});
    if (nk > 0) {
      launch(nk, ScaleF{e, k, L, dinv, z, colT});
      launch(nk * (nk + 1) / 2, UpdateF{e, k, L, colT});
    }
  }
  launch((int64_t)6 * e.n, ApplyF{dinv, z, x});
  for (int i = e.n - 1; i >= 1; --i) {
    const int64_t cnt = i - h_first[i];
    if (cnt > 0) launch(cnt, BackF{e, i, L, x});
  }
}

}  // namespace wide
}  // namespace lvba
