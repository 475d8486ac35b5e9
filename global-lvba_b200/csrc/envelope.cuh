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
  static constexpr int kPairGroups = (kPairs + 127) / 128;       // warpgroups (4 warps) of pair threads
  static constexpr int kPairThreads = kPairGroups * 128;
  static constexpr int kThreads = kPairThreads + 128;            // + one look-ahead warpgroup
  static constexpr int kStride = 38;                             // doubles per transposed block in smem (16B aligned, conflict-free)
  // register re-allocation (setmaxnreg): only needed when 5 warps share an SMSP (P = 31)
  static constexpr bool kRealloc = kThreads > 512;
  static constexpr int kStaggerCycles = 0;   // tools/ubench/p3_pipe.cu: staggering the pair warps does not help (LDS and DFMA already overlap)
  static constexpr int kPairRegs = 104, kAheadRegs = 56;   // 512*104 + 128*56 <= 640*96 (the CTA pool only holds what the CTA owns)
  static constexpr size_t kSmem = sizeof(double) * (size_t)(2 * P * kStride + P * kStride + 2 * P * 36 + 2 * 36 + 2 * 36 + P * 6) +
                                  sizeof(long long) * P + sizeof(int) * (P + 4) + 32;
};

template <int N> LVBA_DEV void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N> LVBA_DEV void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }

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

// 6x6 symmetric inverse by the symmetric sweep operator (Goodnight 1979), no pivoting, entirely in
// registers + warp shuffles: lane l < 21 owns the lower-triangle element (r,c), l = r(r+1)/2 + c.
// Sweeping pivot p:  x_pp <- -1/x_pp ; x_ip <- x_ip / x_pp ; x_ij <- x_ij - x_ip x_jp / x_pp.
// After the six sweeps the matrix holds -A^-1.  The pivots are the d_p of the unpivoted LDL^T that the
// reference's SimplicialLDLT computes on the same (lower-triangle) data.  Writes K (36, row-major).
LVBA_DEV void warp_sym_inverse6(const double* Ain /*36, lower triangle read*/, double* K, int lane, int& bad) {
  const int l = lane < 21 ? lane : 0;
  const int r = (l >= 15) ? 5 : (l >= 10) ? 4 : (l >= 6) ? 3 : (l >= 3) ? 2 : (l >= 1) ? 1 : 0;
  const int c = l - r * (r + 1) / 2;
  double x = Ain[r * 6 + c];
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    const int lpp = p * (p + 1) / 2 + p;
    const int lrp = (r >= p) ? r * (r + 1) / 2 + p : p * (p + 1) / 2 + r;
    const int lcp = (c >= p) ? c * (c + 1) / 2 + p : p * (p + 1) / 2 + c;
    const double d = __shfl_sync(0xffffffffu, x, lpp);
    const double xrp = __shfl_sync(0xffffffffu, x, lrp);
    const double xcp = __shfl_sync(0xffffffffu, x, lcp);
    const double ip = __drcp_rn(d);
    if (r == p && c == p) x = -ip;
    else if (c == p) x = xrp * ip;        // (r,p), r > p
    else if (r == p) x = xcp * ip;        // (p,c), c < p
    else x = x - xrp * xcp * ip;
  }
  if (lane < 21) {
    const double v = -x;
    if (!isfinite(v)) bad = 1;
    K[r * 6 + c] = v;
    K[c * 6 + r] = v;
  }
  __syncwarp();
}

// 6x6 LDL^T without pivoting, in registers (every lane of the look-ahead warp holds the whole lower
// triangle, 21 doubles: no shuffles / shared memory on the dependent chain).  x[i(i+1)/2 + j], i >= j.
// On return x holds the unit-lower factor below the diagonal and 1/d_p on the diagonal — the scalar LDL^T
// the reference's SimplicialLDLT computes on these six rows.
#define LVBA_T(i, j) ((i) * ((i) + 1) / 2 + (j))
LVBA_DEV void sym6_ldlt(double (&x)[21]) {
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    const double ip = __drcp_rn(x[LVBA_T(p, p)]);
    double li[6];
#pragma unroll
    for (int i = p + 1; i < 6; ++i) li[i] = x[LVBA_T(i, p)] * ip;
#pragma unroll
    for (int i = p + 1; i < 6; ++i)
#pragma unroll
      for (int j = p + 1; j <= i; ++j) x[LVBA_T(i, j)] -= li[i] * x[LVBA_T(j, p)];
#pragma unroll
    for (int i = p + 1; i < 6; ++i) x[LVBA_T(i, p)] = li[i];
    x[LVBA_T(p, p)] = ip;
  }
}
// r = t D^-1 with D = L diag(1/f_pp) L^T given the 21 packed factors f (shared or global memory, broadcast reads)
LVBA_DEV void ldlt_solve6(const double* __restrict__ f, double (&t)[6]) {
#pragma unroll
  for (int i = 1; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) t[i] -= f[LVBA_T(i, j)] * t[j];
#pragma unroll
  for (int i = 0; i < 6; ++i) t[i] *= f[LVBA_T(i, i)];
#pragma unroll
  for (int i = 4; i >= 0; --i)
#pragma unroll
    for (int j = i + 1; j < 6; ++j) t[i] -= f[LVBA_T(j, i)] * t[j];
}

// One factorisation instance.  The twisted solve (runtime.cuh) runs two at once (gridDim.x = 2): the top half of
// the pose system in natural order and the bottom half in REVERSED order, each on its own SM; both stop at the
// separator (n_stop < n) and dump their Schur-updated trailing window + forward-substituted rhs.
struct FactorJob {
  EnvView e;
  double* L;       // in: matrix (H + damping) in envelope layout; out: L_ik below the pivots
  double* dinv;    // out: packed LDL^T factors of every pivot block (21 of 36 doubles used)
  double* z;       // in: rhs ; out: forward-substituted rhs of the pivots
  int n_stop;      // number of pivots to eliminate (== e.n for a complete factorisation)
  double* wdump;   // [bs*bs*36] trailing window at n_stop, block (i,j) at ((i-n_stop)*bs + (j-n_stop))*36, bs = e.n - n_stop
  double* zdump;   // [bs*6]
};
struct FactorJobs { FactorJob j[2]; };

template <int P>
__global__ void __launch_bounds__(RegCfg<P>::kThreads, 1)
env_factor_reg_kernel(FactorJobs jobs, const unsigned short* __restrict__ pair_map, int* __restrict__ status,
                      long long* __restrict__ dbg_all) {
  using Cfg = RegCfg<P>;
  const FactorJob& J = jobs.j[blockIdx.x];
  const EnvView e = J.e;
  double* __restrict__ L = J.L;
  double* __restrict__ dinv = J.dinv;
  double* __restrict__ z = J.z;
  const int n_stop = J.n_stop;
  long long* dbg = (blockIdx.x == 0) ? dbg_all : nullptr;
  // optional phase timing (LVBA_FACTOR_TIMING=1): dbg[(k*8 + role)*4 + stamp], role 0..3 = pair warps 0..3, 4..7 = look-ahead warps
#define LVBA_STAMP(role, stamp) do { if (dbg && lane == 0) dbg[((long long)k * 8 + (role)) * 4 + (stamp)] = clock64(); } while (0)
  constexpr int S = Cfg::kStride;
  extern __shared__ __align__(16) double smem_reg[];
  double* sTt0 = smem_reg;                       // [2][P][S] A_ik transposed ([q*6+a] = A[a][q]); parity = pivot column & 1
  double* sLt = sTt0 + 2 * P * S;                // [P][S]    L_ik transposed (current pivot column)
  double* sEnter0 = sLt + P * S;                 // [2][P][36] entering row, by column slot; parity = retiring column & 1
  double* sK0 = sEnter0 + 2 * P * 36;            // [2][36]   D_k^-1
  double* sDg0 = sK0 + 72;                       // [2][36]   diagonal block handed to the look-ahead
  double* sZ = sDg0 + 72;                        // [P][6]
  long long* sRS = reinterpret_cast<long long*>(sZ + P * 6);   // [P]
  int* sFirst = reinterpret_cast<int*>(sRS + P);               // [P]
  int* sNk = sFirst + P;                                       // [4] ring: sNk[k & 3] = last[k] - k
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = e.n;
  const bool is_ahead = tid >= Cfg::kPairThreads;               // look-ahead warpgroup (warp-uniform)

  // L_ik = A_ik D_k^-1 for rows k+1..k+nk: one (row, x) item per thread = row x of L_ik, obtained by the two
  // triangular solves with the LDL^T factors of D_k (sK0: 21 doubles, broadcast reads)
  auto scale_column = [&](int k, int c, int nk) {
    const double* F = sK0 + (k & 1) * 36;
    const double* tb = sTt0 + (k & 1) * P * S;
    // executed by the look-ahead warpgroup only: the pair threads keep all their registers for the live block
    for (int o = tid - Cfg::kPairThreads; o < nk * 6; o += 128) {
      const int h = 1 + o / 6, x = o - (h - 1) * 6;
      int slot = c + h; if (slot >= P) slot -= P;
      const double* t = tb + slot * S;
      double v[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = t[q * 6 + x];
      ldlt_solve6(F, v);
      double* lt = sLt + slot * S;
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) lt[cc * 6 + x] = v[cc];
      double2* g = reinterpret_cast<double2*>(L + (sRS[slot] + (k - sFirst[slot])) * 36 + x * 6);
      g[0] = make_double2(v[0], v[1]); g[1] = make_double2(v[2], v[3]); g[2] = make_double2(v[4], v[5]);
    }
  };

  // ---------------- prologue (all threads, launch register budget)
  for (int r = tid; r < P; r += Cfg::kThreads) {
    if (r < n) { sFirst[r] = e.first[r]; sRS[r] = e.row_start[r]; } else { sFirst[r] = 0x7fffffff; sRS[r] = 0; }
#pragma unroll
    for (int q = 0; q < 6; ++q) sZ[r * 6 + q] = (r < n) ? z[6 * r + q] : 0.0;
  }
  if (tid == 0) { sNk[0] = e.last[0]; sNk[1] = (n > 1) ? e.last[1] - 1 : 0; sNk[2] = (n > 2) ? e.last[2] - 2 : 0; sNk[3] = 0; }
  __syncthreads();

  if (!is_ahead) {
    // =================================================== pair threads: one live 6x6 block in registers
    if (Cfg::kRealloc) reg_alloc<Cfg::kPairRegs>();
    const unsigned short pm = pair_map[tid];       // host-built map: lanes of a warp share few distinct slots
    const bool is_pair = pm != 0xffff;
    const int a = pm & 0xff, b = (pm >> 8) & 0xff; // a >= b
    double C[36];
    auto publish_T = [&](double* dst) {            // dst[q*6+x] = C[x][q]
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double2* d2 = reinterpret_cast<double2*>(dst + q * 6);
        d2[0] = make_double2(C[q], C[6 + q]);
        d2[1] = make_double2(C[12 + q], C[18 + q]);
        d2[2] = make_double2(C[24 + q], C[30 + q]);
      }
    };
    if (is_pair) {
      if (a < n && b >= sFirst[a]) {
        const double2* src = reinterpret_cast<const double2*>(L + (sRS[a] + (b - sFirst[a])) * 36);
#pragma unroll
        for (int q = 0; q < 18; ++q) { const double2 v = src[q]; C[2 * q] = v.x; C[2 * q + 1] = v.y; }
      } else {
#pragma unroll
        for (int q = 0; q < 36; ++q) C[q] = 0.0;
      }
      if (b == 0 && a >= 1) publish_T(sTt0 + a * S);            // column 0
      if (a == 1 && b == 1) {
#pragma unroll
        for (int q = 0; q < 36; ++q) sDg0[36 + q] = C[q];
      }
      if (a == 0 && b == 0) {
#pragma unroll
        for (int q = 0; q < 36; ++q) sK0[q] = C[q];
      }
    }
    __syncthreads();     // (A) column 0 published
    __syncthreads();     // (B) look-ahead group finished D_0^-1 and the first entering rows
    int c = 0;
    for (int k = 0; k < n_stop; ++k) {
      const int cur = k & 1;
      const int nk = sNk[k & 3];
      if (tid < 128) LVBA_STAMP(tid >> 5, 0);
      if (tid < 128) LVBA_STAMP(tid >> 5, 1);
      __syncthreads();                             // L_ik of this pivot column is ready (look-ahead group)
      if (tid < 128) LVBA_STAMP(tid >> 5, 2);
      // Two of the four pair warps on every SMSP start half an operand-period late: the shared-memory pipe
      // (the binding resource: 576 B of operands per thread and step) then serves one half while the other
      // half runs its DFMAs, instead of all 16 warps alternating between the two in lockstep.
      if (Cfg::kStaggerCycles > 0 && (tid & 128)) {
        const unsigned t_0 = (unsigned)clock();
        while ((unsigned)clock() - t_0 < (unsigned)Cfg::kStaggerCycles) { }
      }
      if (is_pair) {
        int da = a - c; if (da < 0) da += P;
        int db = b - c; if (db < 0) db += P;
        const int hi = da > db ? da : db, lo = da > db ? db : da;
        const int islot = da >= db ? a : b, jslot = da >= db ? b : a;
        double* sTn = sTt0 + (cur ^ 1) * P * S;
        if (lo == 0) {                             // column-k block is dead: take the entering block (k+P, .)
          const double2* src = reinterpret_cast<const double2*>(sEnter0 + (cur * P + islot) * 36);
#pragma unroll
          for (int q = 0; q < 18; ++q) { const double2 v = src[q]; C[2 * q] = v.x; C[2 * q + 1] = v.y; }
          if (hi == 1) publish_T(sTn + c * S);     // block (k+P, k+1): last row of column k+1
        } else {
          if (hi <= nk) {
            const double2* lp = reinterpret_cast<const double2*>(sLt + islot * S);
            const double2* tp = reinterpret_cast<const double2*>(sTt0 + (cur * P + jslot) * S);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              const double2 t0 = tp[3 * q], t1 = tp[3 * q + 1], t2 = tp[3 * q + 2];
#pragma unroll
              for (int xx = 0; xx < 3; ++xx) {
                const double2 l = lp[3 * q + xx];
                double* c0 = C + (2 * xx) * 6;
                double* c1 = C + (2 * xx + 1) * 6;
                c0[0] -= l.x * t0.x; c0[1] -= l.x * t0.y; c0[2] -= l.x * t1.x; c0[3] -= l.x * t1.y; c0[4] -= l.x * t2.x; c0[5] -= l.x * t2.y;
                c1[0] -= l.y * t0.x; c1[1] -= l.y * t0.y; c1[2] -= l.y * t1.x; c1[3] -= l.y * t1.y; c1[4] -= l.y * t2.x; c1[5] -= l.y * t2.y;
              }
            }
          }
          if (lo == 1 && hi >= 2) publish_T(sTn + islot * S);    // column k+1, rows k+2..k+P-1
          if (da == 2 % P && db == 2 % P) {          // block (k+2,k+2) for the look-ahead of step k+1
#pragma unroll
            for (int q = 0; q < 36; ++q) sDg0[cur * 36 + q] = C[q];   // (k+2)&1 == k&1
          }
        }
      }
      if (tid < 128) LVBA_STAMP(tid >> 5, 3);
      __syncthreads();
      if (++c == P) c = 0;
    }
    // partial factorisation: hand the Schur-updated trailing window (rows/cols n_stop..n-1) to the separator solve
    if (n_stop < n && J.wdump && is_pair) {
      int da = a - c; if (da < 0) da += P;
      int db = b - c; if (db < 0) db += P;
      const int hi = da > db ? da : db, lo = da > db ? db : da;
      const int bs = n - n_stop;
      if (hi < bs) {
        double2* dst = reinterpret_cast<double2*>(J.wdump + ((long long)hi * bs + lo) * 36);
#pragma unroll
        for (int q = 0; q < 18; ++q) dst[q] = make_double2(C[2 * q], C[2 * q + 1]);
      }
    }
  } else {
    // =================================================== look-ahead warpgroup (4 warps, one per SMSP)
    if (Cfg::kRealloc) reg_dealloc<Cfg::kAheadRegs>();
    const int aw = (tid - Cfg::kPairThreads) >> 5;              // 0: pivot LDL^T + labels, 1: forward substitution, 2,3: row prefetch
    const int pl = tid - Cfg::kPairThreads - 64;                // prefetch lane id over warps 2..3 (0..63), negative otherwise
    int bad = 0;
    constexpr int kPf = (P * 18 + 63) / 64;                     // double2 per prefetch lane per row
    double2 buf[kPf];                                           // row loaded during the previous step
    double zin = 0.0;
    int pf_first = 0x7fffffff; long long pf_rs = 0;             // label of the row to be loaded THIS step (fetched a step earlier)
    auto issue_row = [&](int kc, int rf, long long rrs) {       // row kc+P -> buf (registers; latency hides under the step)
      const int r = kc + P, ck = kc % P;
#pragma unroll
      for (int m = 0; m < kPf; ++m) {
        const int o = pl + 64 * m;
        double2 v = make_double2(0.0, 0.0);
        if (o < P * 18) {
          const int cs = o / 18, w = o - cs * 18;
          int dcol = cs - ck; if (dcol <= 0) dcol += P;          // col = kc + dcol ; dcol == P <=> col == r
          const int col = kc + dcol;
          if (r < n && col >= rf) v = reinterpret_cast<const double2*>(L + (rrs + (col - rf)) * 36)[w];
        }
        buf[m] = v;
      }
    };
    auto retire_row = [&](int kc) {                             // buf -> sEnter[kc & 1]
#pragma unroll
      for (int m = 0; m < kPf; ++m) {
        const int o = pl + 64 * m;
        if (o < P * 18) reinterpret_cast<double2*>(sEnter0 + ((kc & 1) * P) * 36)[o] = buf[m];
      }
    };
    // warp 0: D (36 row-major in `src`, lower triangle read) -> LDL^T factors (21) in `dst` (shared) and in dinv[kc]
    auto factor_pivot = [&](const double* src, double* dst, int kc) {
      double x[21];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) x[LVBA_T(i, j)] = src[i * 6 + j];
      __syncwarp();
      sym6_ldlt(x);
      double chk = 0.0;
#pragma unroll
      for (int q = 0; q < 21; ++q) { dst[q] = x[q]; chk += x[q]; }
      if (!isfinite(chk)) bad = 1;
      __syncwarp();
      if (lane < 21) dinv[(long long)kc * 36 + lane] = dst[lane];
    };
    __syncthreads();     // (A)
    if (aw == 0) {
      factor_pivot(sK0, sK0, 0);
      if (lane == 0) {                                          // slot 0 now describes row P (row 0's label is dead)
        sFirst[0] = (P < n) ? e.first[P] : 0x7fffffff;
        sRS[0] = (P < n) ? e.row_start[P] : 0;
      }
    } else if (aw == 1) {
      if (lane < 6) zin = (P < n) ? z[6 * (long long)P + lane] : 0.0;      // z of row P, stored at step 0
    } else {
      // rows P (needed at step 0) and P+1 (stored at step 0); label of row P+2
      issue_row(0, (P < n) ? e.first[P] : 0x7fffffff, (P < n) ? e.row_start[P] : 0);
      retire_row(0);
      issue_row(1, (P + 1 < n) ? e.first[P + 1] : 0x7fffffff, (P + 1 < n) ? e.row_start[P + 1] : 0);
      pf_first = (P + 2 < n) ? e.first[P + 2] : 0x7fffffff;
      pf_rs = (P + 2 < n) ? e.row_start[P + 2] : 0;
    }
    __syncthreads();     // (B)
    int c = 0;
    for (int k = 0; k < n_stop; ++k) {
      const int cur = k & 1;
      const int nk = sNk[k & 3];
      LVBA_STAMP(4 + aw, 0);
      scale_column(k, c, nk);
      LVBA_STAMP(4 + aw, 1);
      __syncthreads();
      LVBA_STAMP(4 + aw, 2);
      int s1 = c + 1; if (s1 >= P) s1 -= P;
      if (aw == 0) {
        // label of the entering row k+1+P and n_{k+3}: loads issued first, consumed after the factorisation
        int m_first = 0x7fffffff, m_last = 0; long long m_rs = 0;
        if (lane == 0) {
          const int r1 = k + 1 + P;
          m_first = (r1 < n) ? e.first[r1] : 0x7fffffff;
          m_rs = (r1 < n) ? e.row_start[r1] : 0;
          m_last = (k + 3 < n) ? e.last[k + 3] - (k + 3) : 0;
        }
        // look-ahead: D_{k+1} = A_{k+1,k+1} - L_{k+1,k} T_{k+1,k}^T (lane <-> lower-triangle element), then LDL^T
        if (k + 1 < n) {
          double* Kn = sK0 + (cur ^ 1) * 36;
          const double* dg = sDg0 + (cur ^ 1) * 36;
          const double* lt = sLt + s1 * S;
          const double* tt = sTt0 + (cur * P + s1) * S;
          const int l21 = lane < 21 ? lane : 0;
          const int i = (l21 >= 15) ? 5 : (l21 >= 10) ? 4 : (l21 >= 6) ? 3 : (l21 >= 3) ? 2 : (l21 >= 1) ? 1 : 0;
          const int j = l21 - i * (i + 1) / 2;
          double v = dg[i * 6 + j];
          if (nk >= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) v -= lt[q * 6 + i] * tt[q * 6 + j];
          }
          __syncwarp();
          if (lane < 21) Kn[i * 6 + j] = v;                       // lower triangle, row-major 6x6 scratch
          __syncwarp();
          factor_pivot(Kn, Kn, k + 1);
        }
        if (lane == 0) {
          sFirst[s1] = m_first; sRS[s1] = m_rs;                   // slot of row k+1 now describes row k+1+P
          sNk[(k + 3) & 3] = m_last;                              // n_{k+3} (slot last used by n_{k-1})
        }
      } else if (aw == 1) {
        // forward substitution with the final z_k : lane <-> row k+1+lane
        double zk[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) zk[q] = sZ[c * 6 + q];
        if (lane < 6) z[6 * (long long)k + lane] = sZ[c * 6 + lane];
        for (int h = 1 + lane; h <= nk; h += 32) {
          int slot = c + h; if (slot >= P) slot -= P;
          const double2* lt2 = reinterpret_cast<const double2*>(sLt + slot * S);
          double acc[6];
#pragma unroll
          for (int x = 0; x < 6; ++x) acc[x] = 0.0;
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const double2 l0 = lt2[3 * q], l1 = lt2[3 * q + 1], l2 = lt2[3 * q + 2];
            acc[0] += l0.x * zk[q]; acc[1] += l0.y * zk[q]; acc[2] += l1.x * zk[q]; acc[3] += l1.y * zk[q]; acc[4] += l2.x * zk[q]; acc[5] += l2.y * zk[q];
          }
#pragma unroll
          for (int x = 0; x < 6; ++x) sZ[slot * 6 + x] -= acc[x];
        }
        __syncwarp();
        if (lane < 6) {
          sZ[c * 6 + lane] = zin;                               // row k+P takes slot c (loaded one step earlier)
          zin = (k + 1 + P < n) ? z[6 * (long long)(k + 1 + P) + lane] : 0.0;
        }
      } else {
        // ---- warps 2,3: retire the row loaded last step (row k+1+P -> sEnter[(k+1)&1]), issue row k+2+P
        retire_row(k + 1);
        issue_row(k + 2, pf_first, pf_rs);
        const int r3 = k + 3 + P;
        pf_first = (r3 < n) ? e.first[r3] : 0x7fffffff;
        pf_rs = (r3 < n) ? e.row_start[r3] : 0;
      }
      LVBA_STAMP(4 + aw, 3);
      __syncthreads();
      if (++c == P) c = 0;
    }
    if (n_stop < n && J.zdump && aw == 1) {                     // forward-substituted rhs of the separator rows
      for (int o = lane; o < (n - n_stop) * 6; o += 32) {
        const int i = n_stop + o / 6;
        J.zdump[o] = sZ[(i % P) * 6 + o % 6];
      }
    }
    if (bad) status[0] = 1;
  }
#undef LVBA_STAMP
}


// x = D^-1 z with the packed LDL^T factors of every pivot block (register-window path)
__global__ void env_ldl_apply_kernel(int n, const double* __restrict__ dinv, const double* __restrict__ z, double* __restrict__ x) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double t[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) t[q] = z[6 * (long long)k + q];
  ldlt_solve6(dinv + (long long)k * 36, t);
#pragma unroll
  for (int q = 0; q < 6; ++q) x[6 * (long long)k + q] = t[q];
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

// Backward substitution, row oriented, with the rows of L streamed through a 4-deep cp.async ring in
// shared memory (one warp).  Requires row length (blocks left of the diagonal) <= 31.
LVBA_DEV void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
LVBA_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> LVBA_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Block of 7 warps: warps 0..5 own one output (row jr, component cc) of the row being applied, warp 6 is the
// producer that streams future rows of L through a cp.async ring and prefetches their labels and the x entries
// that enter the 32-row window.  Two block barriers per row; no global-memory latency on the dependent chain.
constexpr int kBsThreads = 224, kBsDepth = 5;
struct BacksolveJob {
  EnvView e;
  const double* L;
  double* x;        // in: D^-1 z for the pivots (rows < n_given) and the FINAL solution for rows >= n_given ; out: solution
  int n_given;      // rows >= n_given are given (separator of the twisted solve); == e.n for a plain solve
};
struct BacksolveJobs { BacksolveJob j[2]; };

__global__ void __launch_bounds__(kBsThreads, 1)
env_backsolve_ring_kernel(BacksolveJobs jobs) {
  constexpr int W = 32, D = kBsDepth, ROWMAX = 31 * 36;
  __shared__ __align__(16) double ring[D][ROWMAX];
  __shared__ double sX[W * 6];
  __shared__ int sF[D], sCnt[D];
  const BacksolveJob& J = jobs.j[blockIdx.x];
  const EnvView e = J.e;
  const double* __restrict__ L = J.L;
  double* __restrict__ x = J.x;
  const int n_given = J.n_given;
  const int tid = threadIdx.x, lane = tid & 31, n = e.n;
  const bool producer = tid >= 192;
  for (int idx = tid; idx < W * 6; idx += kBsThreads) {
    const int r = n - 1 - idx / 6;                       // rows n-1 .. n-32
    if (r >= 0) sX[(r % W) * 6 + idx % 6] = x[6 * (long long)r + idx % 6];
  }
  int nf = 0; long long nrs = 0; double xin = 0.0;
  auto issue = [&](int i) {                               // stream row i (label in nf/nrs) into ring[i % D]
    if (i >= 0) {
      const int upto = i < n_given ? i : n_given;         // given rows only act on the pivots' columns
      const int cnt = upto > nf ? upto - nf : 0;
      if (lane == 0) { sF[i % D] = nf; sCnt[i % D] = cnt; }
      const double* row = L + nrs * 36;
      for (int o = lane; o < cnt * 18; o += 32) cp_async16(&ring[i % D][2 * o], row + 2 * o);
    }
    cp_async_commit();
  };
  if (producer) {
    for (int d = 0; d < D - 1; ++d) {
      const int i = n - 1 - d;
      if (i >= 0) { nf = e.first[i]; nrs = e.row_start[i]; }
      issue(i);
    }
    const int inext = n - 1 - (D - 1);
    if (inext >= 0) { nf = e.first[inext]; nrs = e.row_start[inext]; }
    if (lane < 6 && n - 1 - W >= 0) xin = x[6 * (long long)(n - 1 - W) + lane];
  }
  __syncthreads();
  const int jr = tid / 6, cc = tid - jr * 6;              // consumer output (tid < 192 -> jr < 32)
  for (int i = n - 1; i >= 0; --i) {
    if (producer) {
      issue(i - (D - 1));
      const int i2 = i - D;
      if (i2 >= 0) { nf = e.first[i2]; nrs = e.row_start[i2]; }
      cp_async_wait<D - 1>();
    }
    __syncthreads();                                      // (A) row i visible; x_i final
    if (!producer) {
      const int f = sF[i % D], cnt = sCnt[i % D];
      const double* cur = ring[i % D];
      if (tid < 6) x[6 * (long long)i + tid] = sX[(i % W) * 6 + tid];
      if (jr < cnt) {
        const double* xi = sX + (i % W) * 6;
        const double* lb = cur + jr * 36 + cc;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) v += lb[q * 6] * xi[q];
        sX[((f + jr) % W) * 6 + cc] -= v;
      }
    }
    __syncthreads();                                      // (B) updates applied; slot of row i is free
    if (producer && lane < 6) {
      if (i - W >= 0) sX[(i % W) * 6 + lane] = xin;
      if (i - 1 - W >= 0) xin = x[6 * (long long)(i - 1 - W) + lane];
    }
  }
}

// ---- twisted solve helpers -------------------------------------------------------------------------------
// Reversed copy of the bottom part: row r' of the reversed matrix = original row n-1-r'; its lower blocks are the
// transposes of the original column's blocks.  One CTA per reversed row.
__global__ void env_reverse_gather_kernel(EnvView eo, EnvView eb, const double* __restrict__ Lo, double* __restrict__ Lb,
                                          const double* __restrict__ zo, double* __restrict__ zb) {
  const int n = eo.n;
  for (int rp = blockIdx.x; rp < eb.n; rp += gridDim.x) {
    const int ir = n - 1 - rp;                            // original index of this reversed row
    const int f = eb.first[rp];
    const long long base = eb.row_start[rp] * 36;
    const int nblk = rp - f + 1;
    for (int o = threadIdx.x; o < nblk * 36; o += blockDim.x) {
      const int cb = o / 36, el = o - cb * 36, a = el / 6, b2 = el - a * 6;
      const int ic = n - 1 - (f + cb);                    // original row of the coupled pose (ic >= ir)
      // reversed block (rp, f+cb)[a][b2] = original block (ic, ir)[b2][a]
      Lb[base + o] = Lo[(eo.row_start[ic] + (ir - eo.first[ic])) * 36 + b2 * 6 + a];
    }
    if (threadIdx.x < 6) zb[6 * (long long)rp + threadIdx.x] = zo[6 * (long long)ir + threadIdx.x];
  }
}

// Separator system: S = W_top + W_bot^T(reversed) - A_sep ; z_sep = z_top + z_bot(reversed) - z_orig
__global__ void env_twist_combine_kernel(EnvView eo, int m, int bs, const double* __restrict__ Lo, const double* __restrict__ zo,
                                         const double* __restrict__ wtop, const double* __restrict__ wbot,
                                         const double* __restrict__ ztop, const double* __restrict__ zbot,
                                         double* __restrict__ Lsep, double* __restrict__ zsep) {
  const int nblk = bs * (bs + 1) / 2;
  for (int o = threadIdx.x; o < nblk * 36; o += blockDim.x) {
    const int blk = o / 36, el = o - blk * 36, a = el / 6, b2 = el - a * 6;
    int si, sj;
    tri_decode(blk, si, sj);                              // si >= sj, dense lower envelope: block index si(si+1)/2 + sj
    const int i = m + si, j = m + sj;
    const double orig = (j >= eo.first[i]) ? Lo[(eo.row_start[i] + (j - eo.first[i])) * 36 + el] : 0.0;
    const int ur = bs - 1 - sj, uc = bs - 1 - si;         // reversed-local indices: (ur >= uc)
    const double top = wtop[((long long)si * bs + sj) * 36 + el];
    const double bot = wbot[((long long)ur * bs + uc) * 36 + b2 * 6 + a];
    Lsep[o] = top + bot - orig;
  }
  for (int o = threadIdx.x; o < bs * 6; o += blockDim.x) {
    const int si = o / 6, q = o - si * 6;
    zsep[o] = ztop[o] + zbot[(bs - 1 - si) * 6 + q] - zo[6 * (long long)(m + si) + q];
  }
}

// x_top[m..m+bs) = x_sep ; x_bot tail (reversed) = x_sep
__global__ void env_twist_place_sep_kernel(int m, int bs, int nb_stop, const double* __restrict__ xsep, double* __restrict__ xtop, double* __restrict__ xbot) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= bs * 6) return;
  const int si = o / 6, q = o - si * 6;
  xtop[6 * (long long)(m + si) + q] = xsep[o];
  xbot[6 * (long long)(nb_stop + (bs - 1 - si)) + q] = xsep[o];
}
// x[orig i] = x_bot[n-1-i] for the bottom pivots
__global__ void env_twist_scatter_kernel(int n, int nb_stop, const double* __restrict__ xbot, double* __restrict__ x) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nb_stop * 6) return;
  const int rp = o / 6, q = o - rp * 6;
  x[6 * (long long)(n - 1 - rp) + q] = xbot[o];
}

}  // namespace lvba
