// envelope.cuh — block-envelope (skyline) storage of the symmetric 6x6-block pose system and its
// LDL^T solve on the device.
//
// Replaces, for hot path A, the dense->triplet->Eigen::SimplicialLDLT sequence of
// BALM2::damping_iter (reference include/BALM/bavoxel.hpp:692-710; SimplicialLDLT reads the lower
// triangle only — SURVEY.md Q4 — and performs LDL^T without pivoting, which is required because the
// BALM2 Newton Hessian is indefinite — Q5), and for hot path B the DENSE_SCHUR Cholesky of the reduced
// camera system inside ceres::Solve (src/lvba_system.cpp:1574,1643).
//
// Storage: lower block triangle by rows.  Row r keeps the 6x6 blocks of columns first[r]..r
// contiguously: block (r,c) lives at (row_start[r] + c - first[r]) * 36, row-major, element [a][b] =
// M[6r+a, 6c+b].  first[] is made monotone non-decreasing on the host so that the row set below a
// pivot column k is the contiguous range k+1..last[k]; all LDL^T fill stays inside the envelope.
#pragma once
#include "common.cuh"

namespace lvba {

struct EnvView {
  int n;                       // block rows
  const int* first;            // [n]
  const long long* row_start;  // [n+1] in blocks
  const int* last;             // [n]   last[k] = max row i with first[i] <= k
  long long nblocks;
};

LVBA_DEV long long env_block(const EnvView& e, int r, int c) { return e.row_start[r] + (c - e.first[r]); }

constexpr int kEnvMaxCol = 320;     // max rows below one pivot column handled by the factor kernel
constexpr int kFactorThreads = 1024;

// L <- H + diag(dadd)   (dadd: [6n] added on the scalar diagonal).  Grid-stride over doubles.
__global__ void env_copy_damped_kernel(EnvView e, const double* __restrict__ H, const double* __restrict__ dadd,
                                       double* __restrict__ L) {
  const long long total = e.nblocks * 36;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    L[i] = H[i];
}
__global__ void env_add_diag_kernel(EnvView e, const double* __restrict__ dadd, double* __restrict__ L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * e.n) {
    const int r = i / 6, a = i % 6;
    L[env_block(e, r, r) * 36 + a * 7] += dadd[i];
  }
}
// diag[6r+a] = H[(r,r)][a][a]
__global__ void env_get_diag_kernel(EnvView e, const double* __restrict__ H, double* __restrict__ diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * e.n) {
    const int r = i / 6, a = i % 6;
    diag[i] = H[env_block(e, r, r) * 36 + a * 7];
  }
}

// decode lower-triangular linear index t -> (i, j), j <= i
LVBA_DEV void tri_decode(int t, int& i, int& j) {
  int ii = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
  while (ii * (ii + 1) / 2 > t) --ii;
  i = ii;
  j = t - ii * (ii + 1) / 2;
}

// Right-looking block LDL^T, in place on L (which enters holding H + damping), with the forward
// substitution of the right-hand side fused in.  One CTA walks the pivot columns in order; the
// parallelism is inside one column step (<= n(n+1)/2 trailing 6x6 blocks, n = last[k]-k).
//   after return:  block (i,k), i>k  holds  L_ik = A_ik D_k^-1
//                  dinv[k] (36)      holds  D_k^-1   (D_k = Schur-updated diagonal block)
//                  z[6k..]           holds  (L^-1 b)_k
// status[0] = 1 if a non-finite pivot inverse appeared.
__global__ void __launch_bounds__(kFactorThreads, 1)
env_factor_kernel(EnvView e, double* __restrict__ L, double* __restrict__ dinv, double* __restrict__ z,
                  int* __restrict__ status) {
  extern __shared__ double smem[];
  double* sT = smem;                       // [kEnvMaxCol][36]  A_ik before scaling
  double* sL = sT + kEnvMaxCol * 36;       // [kEnvMaxCol][36]  L_ik = A_ik D^-1
  double* sK = sL + kEnvMaxCol * 36;       // [36] D_k^-1
  double* sZ = sK + 36;                    // [6]  z_k
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bad = 0;

  for (int k = 0; k < e.n; ++k) {
    const int n = e.last[k] - k;           // rows k+1 .. k+n below the pivot
    // ---- phase 1: stage the pivot block (lower triangle mirrored, as SimplicialLDLT reads it) and
    //      column k into shared memory
    if (tid < 36) {
      const long long bk = env_block(e, k, k) * 36;
      const int r = tid / 6, c = tid % 6;
      sK[tid] = (r >= c) ? L[bk + r * 6 + c] : L[bk + c * 6 + r];
    }
    for (int idx = tid; idx < n * 36; idx += kFactorThreads) {
      const int i = idx / 36, el = idx % 36;
      sT[idx] = L[env_block(e, k + 1 + i, k) * 36 + el];
    }
    __syncthreads();
    if (warp == 0) {
      // D_k^-1 by in-place Gauss-Jordan WITHOUT pivoting (same pivots d_p as the scalar LDL^T the
      // reference runs): one warp, 36 elements over 32 lanes with a second slot on lanes 0..3.
      const int e0 = lane, e1 = 32 + lane;           // e1 valid for lane < 4
      const int r0 = e0 / 6, c0 = e0 % 6, r1 = (e1 < 36) ? e1 / 6 : 0, c1 = (e1 < 36) ? e1 % 6 : 0;
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        const double piv = sK[p * 7];
        const double ip = 1.0 / piv;
        const double x0 = sK[e0], xr0 = sK[r0 * 6 + p], xc0 = sK[p * 6 + c0];
        const double x1 = (e1 < 36) ? sK[e1] : 0.0, xr1 = sK[r1 * 6 + p], xc1 = sK[p * 6 + c1];
        __syncwarp();
        double n0, n1;
        if (r0 == p && c0 == p) n0 = ip; else if (r0 == p) n0 = xc0 * ip; else if (c0 == p) n0 = -xr0 * ip; else n0 = x0 - xr0 * xc0 * ip;
        if (r1 == p && c1 == p) n1 = ip; else if (r1 == p) n1 = xc1 * ip; else if (c1 == p) n1 = -xr1 * ip; else n1 = x1 - xr1 * xc1 * ip;
        sK[e0] = n0;
        if (e1 < 36) sK[e1] = n1;
        __syncwarp();
      }
      double chk = sK[lane] + ((lane < 4) ? sK[32 + lane] : 0.0);
      if (!isfinite(chk)) bad = 1;
      dinv[(long long)k * 36 + lane] = sK[lane];
      if (lane < 4) dinv[(long long)k * 36 + 32 + lane] = sK[32 + lane];
    }
    if (tid >= 32 && tid < 38) sZ[tid - 32] = z[6 * k + (tid - 32)];
    __syncthreads();
    // ---- phase 2: L_ik = A_ik D^-1, written to smem and back to global
    for (int idx = tid; idx < n * 36; idx += kFactorThreads) {
      const int i = idx / 36, el = idx % 36, a = el / 6, b = el % 6;
      const double* t = sT + i * 36 + a * 6;
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) s += t[c] * sK[c * 6 + b];
      sL[idx] = s;
      L[env_block(e, k + 1 + i, k) * 36 + el] = s;
    }
    __syncthreads();
    // ---- phase 3: trailing update A_ij -= L_ik A_jk^T for k < j <= i <= k+n; half a block (3 rows)
    //      per thread.  Fused forward substitution: z_i -= L_ik z_k.
    const int nhalf = n * (n + 1);            // (n(n+1)/2 blocks) * 2 halves
    for (int h = tid; h < nhalf; h += kFactorThreads) {
      int i, j;
      tri_decode(h >> 1, i, j);
      const int a0 = (h & 1) * 3;
      const double* li = sL + i * 36 + a0 * 6;
      const double* tj = sT + j * 36;
      double* dst = L + env_block(e, k + 1 + i, k + 1 + j) * 36 + a0 * 6;
      double acc[18];
#pragma unroll
      for (int q = 0; q < 18; ++q) acc[q] = dst[q];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < 6; ++c) s += li[a * 6 + c] * tj[b * 6 + c];
          acc[a * 6 + b] -= s;
        }
#pragma unroll
      for (int q = 0; q < 18; ++q) dst[q] = acc[q];
    }
    for (int idx = tid; idx < n * 6; idx += kFactorThreads) {
      const int i = idx / 6, a = idx % 6;
      const double* li = sL + i * 36 + a * 6;
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) s += li[c] * sZ[c];
      z[6 * (k + 1 + i) + a] -= s;
    }
    __syncthreads();
  }
  if (bad) status[0] = 1;
}

// x = L^-T D^-1 z  (z from env_factor_kernel).  One warp walks the columns backwards; lanes split
// the rows below the pivot, partial 6-vectors are shuffle-reduced.
__global__ void __launch_bounds__(32, 1)
env_backsolve_kernel(EnvView e, const double* __restrict__ L, const double* __restrict__ dinv,
                     const double* __restrict__ z, double* __restrict__ x) {
  const int lane = threadIdx.x;
  for (int k = e.n - 1; k >= 0; --k) {
    const int n = e.last[k] - k;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += 32) {
      const double* b = L + env_block(e, k + 1 + i, k) * 36;
      const double* xi = x + 6 * (k + 1 + i);
      double xv[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) xv[a] = xi[a];
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < 6; ++c) s[c] += b[a * 6 + c] * xv[a];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = warp_sum(s[c]);
    if (lane < 6) {
      const double* K = dinv + (long long)k * 36 + lane * 6;
      double w = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) w += K[c] * z[6 * k + c];
      // select s[lane] without dynamic register indexing
      double sl = (lane == 0) ? s[0] : (lane == 1) ? s[1] : (lane == 2) ? s[2] : (lane == 3) ? s[3] : (lane == 4) ? s[4] : s[5];
      x[6 * k + lane] = w - sl;
    }
    __syncwarp();
    __threadfence_block();
  }
}


// =====================================================================================================
// v2: register-resident sliding-window factorisation for envelopes whose column height is < P (P <= 32).
//
// The trailing window of a banded LDL^T holds P(P+1)/2 live 6x6 blocks (rows/cols k..k+P-1).  A 6x6x6
// block update costs 216 DFMA but would move 144 doubles if the target lived in shared memory or L2, so
// one SM's shared-memory (128 B/clk) and L2 bandwidth cap the step at ~2x the FP64 time.  Here every
// live block lives in the REGISTERS of one thread for its whole life:
//   * thread t <-> unordered slot pair {a,b}, a >= b, a,b in [0,P): at step k it owns block (i,j) with
//     {i mod P, j mod P} = {a,b} and k <= j <= i < k+P.  Exactly one of (a,b)/(b,a) is live at a time, and
//     when column k retires, block (i,k)'s registers are re-used for the entering block (k+P, i);
//   * per step:  P1a column threads publish A_ik (transposed) to shared memory and load their entering
//                block from global (its latency hides under P3);
//                P1b four warps form L_ik = A_ik D_k^-1, store it to global + shared;
//                P3  every trailing thread does C -= L_i T_j^T from shared-memory operands (LDS.128,
//                one operand broadcast per warp), while a dedicated pivot warp looks ahead: it finishes
//                block (k+1,k+1), inverts it (Gauss-Jordan, no pivoting), applies the fused forward
//                substitution z_i -= L_ik z_k and fetches the row metadata of the next entering row.
//   Three block barriers per pivot column; no global-memory round trip on the critical path.
template <int P>
struct RegCfg {
  static constexpr int kPairs = P * (P + 1) / 2;
  static constexpr int kPairWarps = (kPairs + 31) / 32;
  static constexpr int kThreads = kPairWarps * 32 + 32;          // + the pivot warp
  static constexpr int kScaleThreads = (kPairWarps < 4 ? kPairWarps : 4) * 32;
  static constexpr int kStride = 38;                             // doubles per transposed block in smem (16B aligned, conflict-free)
};

LVBA_DEV void warp_gj_inverse36(double* K, int lane, int& bad) {
  // in-place Gauss-Jordan without pivoting on K (36 doubles in shared memory), one warp
  const int e0 = lane, e1 = 32 + lane;
  const int r0 = e0 / 6, c0 = e0 % 6, r1 = (e1 < 36) ? e1 / 6 : 0, c1 = (e1 < 36) ? e1 % 6 : 0;
  // mirror the lower triangle (the reference's SimplicialLDLT reads only the lower triangle)
  const double m0 = (r0 >= c0) ? K[e0] : K[c0 * 6 + r0];
  const double m1 = (e1 < 36) ? ((r1 >= c1) ? K[e1] : K[c1 * 6 + r1]) : 0.0;
  __syncwarp();
  K[e0] = m0;
  if (e1 < 36) K[e1] = m1;
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    const double ip = 1.0 / K[p * 7];
    const double x0 = K[e0], xr0 = K[r0 * 6 + p], xc0 = K[p * 6 + c0];
    const double x1 = (e1 < 36) ? K[e1] : 0.0, xr1 = K[r1 * 6 + p], xc1 = K[p * 6 + c1];
    __syncwarp();
    double n0, n1;
    if (r0 == p && c0 == p) n0 = ip; else if (r0 == p) n0 = xc0 * ip; else if (c0 == p) n0 = -xr0 * ip; else n0 = x0 - xr0 * xc0 * ip;
    if (r1 == p && c1 == p) n1 = ip; else if (r1 == p) n1 = xc1 * ip; else if (c1 == p) n1 = -xr1 * ip; else n1 = x1 - xr1 * xc1 * ip;
    K[e0] = n0;
    if (e1 < 36) K[e1] = n1;
    __syncwarp();
  }
  const double chk = K[lane] + ((lane < 4) ? K[32 + lane] : 0.0);
  if (!isfinite(chk)) bad = 1;
}

template <int P>
__global__ void __launch_bounds__(RegCfg<P>::kThreads, 1) __maxnreg__(P == 32 ? 112 : 160)
env_factor_reg_kernel(EnvView e, double* __restrict__ L, double* __restrict__ dinv, double* __restrict__ z,
                      int* __restrict__ status) {
  using Cfg = RegCfg<P>;
  constexpr int S = Cfg::kStride;
  __shared__ __align__(16) double sTt[P * S];   // A_ik transposed: [slot][q*6 + a] = A[a][q]
  __shared__ __align__(16) double sLt[P * S];   // L_ik transposed
  __shared__ double sK[2][36];                  // D_k^-1 (parity k&1)
  __shared__ double sDg[2][36];                 // diagonal block handed to the look-ahead (parity of its row)
  __shared__ double sZ[P * 6];                  // forward-substitution window of z, by row slot
  __shared__ long long sRS[P];                  // row_start of the row living in each slot
  __shared__ int sFirst[P];                     // first    "
  __shared__ int sNk[2];                        // last[k]-k (parity k&1)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_pair = tid < Cfg::kPairs;
  const bool is_pivot_warp = warp == Cfg::kPairWarps;
  const int n = e.n;
  int a = 0, b = 0;
  if (is_pair) tri_decode(tid, a, b);           // a >= b
  double C[36];
  int bad = 0;

  auto load_block = [&](int r, int col, int slot) {
    if (r < n && col >= sFirst[slot]) {
      const double2* src = reinterpret_cast<const double2*>(L + (sRS[slot] + (col - sFirst[slot])) * 36);
#pragma unroll
      for (int q = 0; q < 18; ++q) { const double2 v = src[q]; C[2 * q] = v.x; C[2 * q + 1] = v.y; }
    } else {
#pragma unroll
      for (int q = 0; q < 36; ++q) C[q] = 0.0;
    }
  };

  // ---------------- prologue: rows 0..P-1
  for (int r = tid; r < P; r += Cfg::kThreads) {
    if (r < n) { sFirst[r] = e.first[r]; sRS[r] = e.row_start[r]; }
    else { sFirst[r] = 0x7fffffff; sRS[r] = 0; }
#pragma unroll
    for (int q = 0; q < 6; ++q) sZ[r * 6 + q] = (r < n) ? z[6 * r + q] : 0.0;
  }
  __syncthreads();
  if (is_pair) load_block(a, b, a);
  __syncthreads();
  if (is_pivot_warp) {
    // D_0^-1 ; A_11 for the first look-ahead ; n_0 ; slot 0 <- row P
    const long long b0 = sRS[0] * 36;             // block (0,0) is the first block of row 0
    sK[0][lane] = L[b0 + lane];
    if (lane < 4) sK[0][32 + lane] = L[b0 + 32 + lane];
    if (n > 1) {
      const long long b1 = (sRS[1] + (1 - sFirst[1])) * 36;
      sDg[1][lane] = L[b1 + lane];
      if (lane < 4) sDg[1][32 + lane] = L[b1 + 32 + lane];
    }
    __syncwarp();
    warp_gj_inverse36(sK[0], lane, bad);
    dinv[lane] = sK[0][lane];
    if (lane < 4) dinv[32 + lane] = sK[0][32 + lane];
    if (lane == 0) {
      sNk[0] = e.last[0];
      if (P < n) { sFirst[0] = e.first[P]; sRS[0] = e.row_start[P]; } else { sFirst[0] = 0x7fffffff; sRS[0] = 0; }
    }
  }
  __syncthreads();

  int c = 0;                                     // k mod P
  for (int k = 0; k < n; ++k) {
    const int nk = sNk[k & 1];
    int da = a - c; if (da < 0) da += P;
    int db = b - c; if (db < 0) db += P;
    const int hi = da > db ? da : db, lo = da > db ? db : da;
    const int islot = da >= db ? a : b, jslot = da >= db ? b : a;     // slots of the block's row / column
    // ---- P1a: publish the pivot column, fetch entering blocks
    if (is_pair && lo == 0) {
      if (hi >= 1 && hi <= nk) {
        double* dst = sTt + islot * S;
#pragma unroll
        for (int q = 0; q < 6; ++q)
#pragma unroll
          for (int x = 0; x < 6; ++x) dst[q * 6 + x] = C[x * 6 + q];
      }
      load_block(k + P, (hi == 0) ? k + P : k + hi, c);
    }
    __syncthreads();
    // ---- P1b: L_ik = A_ik D_k^-1
    if (tid < Cfg::kScaleThreads) {
      const double* K = sK[k & 1];
      for (int o = tid; o < nk * 36; o += Cfg::kScaleThreads) {
        const int h = 1 + o / 36, el = o - (h - 1) * 36, x = el / 6, cc = el - x * 6;
        int slot = c + h; if (slot >= P) slot -= P;
        const double* t = sTt + slot * S;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) v += t[q * 6 + x] * K[q * 6 + cc];
        sLt[slot * S + cc * 6 + x] = v;
        L[(sRS[slot] + (k - sFirst[slot])) * 36 + el] = v;
      }
    }
    __syncthreads();
    // ---- P3: trailing update (pair threads) | look-ahead (pivot warp)
    if (is_pair) {
      if (lo >= 1 && hi <= nk) {
        const double2* lp = reinterpret_cast<const double2*>(sLt + islot * S);
        const double2* tp = reinterpret_cast<const double2*>(sTt + jslot * S);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double2 l0 = lp[3 * q], l1 = lp[3 * q + 1], l2 = lp[3 * q + 2];
          const double2 t0 = tp[3 * q], t1 = tp[3 * q + 1], t2 = tp[3 * q + 2];
          const double lv[6] = {l0.x, l0.y, l1.x, l1.y, l2.x, l2.y};
          const double tv[6] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y};
#pragma unroll
          for (int x = 0; x < 6; ++x)
#pragma unroll
            for (int y = 0; y < 6; ++y) C[x * 6 + y] -= lv[x] * tv[y];
        }
      }
      if (da == 2 % P && db == 2 % P) {            // block (k+2,k+2): hand it to the look-ahead of step k+1
        double* dst = sDg[(k + 2) & 1];
#pragma unroll
        for (int q = 0; q < 36; ++q) dst[q] = C[q];
      }
    } else if (is_pivot_warp) {
      // forward substitution with the final z_k
      double zk[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) zk[q] = sZ[c * 6 + q];
      if (lane < 6) z[6 * (long long)k + lane] = sZ[c * 6 + lane];
      for (int o = lane; o < nk * 6; o += 32) {
        const int h = 1 + o / 6, x = o - (h - 1) * 6;
        int slot = c + h; if (slot >= P) slot -= P;
        const double* lt = sLt + slot * S;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) v += lt[q * 6 + x] * zk[q];
        sZ[slot * 6 + x] -= v;
      }
      __syncwarp();
      if (lane < 6) sZ[c * 6 + lane] = (k + P < n) ? z[6 * (long long)(k + P) + lane] : 0.0;
      // look-ahead: D_{k+1} = A_{k+1,k+1} - L_{k+1,k} T_{k+1,k}^T, then invert
      if (k + 1 < n) {
        double* Kn = sK[(k + 1) & 1];
        const double* dg = sDg[(k + 1) & 1];
        int s1 = c + 1; if (s1 >= P) s1 -= P;
        const double* lt = sLt + s1 * S;
        const double* tt = sTt + s1 * S;
        const int e0 = lane, e1 = 32 + lane;
        double v0 = dg[e0], v1 = (e1 < 36) ? dg[e1] : 0.0;
        if (nk >= 1) {
          const int x0 = e0 / 6, y0 = e0 % 6, x1 = (e1 < 36) ? e1 / 6 : 0, y1 = (e1 < 36) ? e1 % 6 : 0;
#pragma unroll
          for (int q = 0; q < 6; ++q) { v0 -= lt[q * 6 + x0] * tt[q * 6 + y0]; v1 -= lt[q * 6 + x1] * tt[q * 6 + y1]; }
        }
        Kn[e0] = v0;
        if (e1 < 36) Kn[e1] = v1;
        __syncwarp();
        warp_gj_inverse36(Kn, lane, bad);
        dinv[(long long)(k + 1) * 36 + lane] = Kn[lane];
        if (lane < 4) dinv[(long long)(k + 1) * 36 + 32 + lane] = Kn[32 + lane];
        if (lane == 0) {
          sNk[(k + 1) & 1] = e.last[k + 1] - (k + 1);
          const int r = k + 1 + P;
          if (r < n) { sFirst[s1] = e.first[r]; sRS[s1] = e.row_start[r]; } else { sFirst[s1] = 0x7fffffff; sRS[s1] = 0; }
        }
      }
    }
    __syncthreads();
    if (++c == P) c = 0;
  }
  if (bad) status[0] = 1;
}

// x = D^-1 z  (block diagonal solve, fully parallel)
__global__ void env_dinv_apply_kernel(int n, const double* __restrict__ dinv, const double* __restrict__ z, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * n) return;
  const int k = i / 6, r = i - 6 * k;
  const double* K = dinv + (long long)k * 36 + r * 6;
  const double* zz = z + 6 * (long long)k;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 6; ++q) s += K[q] * zz[q];
  x[i] = s;
}

// Backward substitution x <- L^-T x, row oriented: once x_i is final, row i of L (contiguous in the
// envelope) updates every pending x_j, j in [first[i], i).  One warp; the next row's blocks are
// prefetched into registers while the current row is applied.  Requires row length < 32 blocks.
__global__ void __launch_bounds__(32, 1)
env_backsolve_row_kernel(EnvView e, const double* __restrict__ L, double* __restrict__ x) {
  constexpr int W = 32;
  __shared__ double sX[W * 6];
  const int lane = threadIdx.x, n = e.n;
  // window rows (i-W, i]; slot = row % W
  for (int r = n - 1 - lane; r >= 0 && r > n - 1 - W; r -= 32)
#pragma unroll
    for (int q = 0; q < 6; ++q) sX[(r % W) * 6 + q] = x[6 * (long long)r + q];
  __syncwarp();
  double cur[36], nxt[36];
  auto fetch = [&](int i, double* buf) {
    if (i < 0) return;
    const int f = e.first[i], cnt = i - f;
    const double* row = L + e.row_start[i] * 36;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int o = lane + 32 * m;
      if (o < cnt * 6) {
        const int jr = o / 6, cc = o - jr * 6;
#pragma unroll
        for (int q = 0; q < 6; ++q) buf[m * 6 + q] = row[jr * 36 + q * 6 + cc];
      }
    }
  };
  fetch(n - 1, cur);
  for (int i = n - 1; i >= 0; --i) {
    fetch(i - 1, nxt);
    const int f = e.first[i], cnt = i - f;
    double xi[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) xi[q] = sX[(i % W) * 6 + q];
    if (lane < 6) x[6 * (long long)i + lane] = sX[(i % W) * 6 + lane];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int o = lane + 32 * m;
      if (o < cnt * 6) {
        const int jr = o / 6, cc = o - jr * 6;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) v += cur[m * 6 + q] * xi[q];
        sX[((f + jr) % W) * 6 + cc] -= v;
      }
    }
    // entering row i-W takes the slot row i just left
    if (lane < 6 && i - W >= 0) sX[(i % W) * 6 + lane] = x[6 * (long long)(i - W) + lane];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 36; ++q) cur[q] = nxt[q];
  }
}

}  // namespace lvba
