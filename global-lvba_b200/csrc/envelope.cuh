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
#include "env_types.h"

namespace lvba {

LVBA_DEV long long env_block(const EnvView& e, int r, int c) { return e.row_start[r] + (c - e.first[r]); }

constexpr int kEnvMaxCol = 320;     // max rows below one pivot column handled by the factor kernel
constexpr int kFactorThreads = 1024;

// L += diag(dadd)   (dadd: [6n] added on the scalar diagonal of the copy of H that the factorisation overwrites)
__global__ void env_add_diag_kernel(EnvView e, const double* __restrict__ dadd, double* __restrict__ L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * e.n) {
    const int r = i / 6, a = i % 6;
    L[env_block(e, r, r) * 36 + a * 7] += dadd[i];
  }
}
// y += x
__global__ void env_axpy_kernel(long long n, const double* __restrict__ x, double* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
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


// register re-allocation between warpgroups (setmaxnreg): used by the register-window factorisation (factor_la.cuh)
template <int N> LVBA_DEV void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N> LVBA_DEV void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
#define LVBA_T(i, j) ((i) * ((i) + 1) / 2 + (j))

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

// status[0] |= status[1..n-1]  (the twisted solve has one flag per factorisation instance)
__global__ void env_status_or_kernel(int* __restrict__ status, int n) {
  if (threadIdx.x == 0) { int v = 0; for (int i = 0; i < n; ++i) v |= status[i]; status[0] = v; }
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
      // reversed block (rp, f+cb)[a][b2] = original block (ic, ir)[b2][a]; a diagonal block is read through its lower
      // triangle only (its upper triangle is unspecified)
      const int ra = (ic == ir && a > b2) ? a : b2, rb = (ic == ir && a > b2) ? b2 : a;
      Lb[base + o] = Lo[(eo.row_start[ic] + (ir - eo.first[ic])) * 36 + ra * 6 + rb];
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
