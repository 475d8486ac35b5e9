// nd_kernels.cuh — device kernels of the substructured block LDL^T (nd_plan.h / nd_passes.h): the multi-right-hand-side
// forward substitution ("spike") through a finished block factor and the symmetric rank update it feeds.
//
// Both serve the solve the reference gets from Eigen::SimplicialLDLT (include/BALM/bavoxel.hpp:695-710) / Ceres
// DENSE_SCHUR (src/lvba_system.cpp:1573-1575).  Neither is on the pivot chain of a factorisation: they only read
// finished columns of L, so they run on SMs the factorising CTAs leave idle (one CTA per group of right-hand sides / per
// tile of the update), FP64 FMA + shared memory, no HBM traffic to speak of (everything is L2 resident).
#pragma once
#include "envelope.cuh"
#include "factor_la.cuh"
#include "nd_passes.h"

namespace lvba {

// grid-stride item pass (functors of nd_passes.h)
template <class F>
__global__ void __launch_bounds__(256) nd_pass_kernel(long long n, F f) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) f(i);
}

// ---------------------------------------------------------------------------------------------------------------
// Spike: Z = L^-1 E for KS right-hand sides (SpikeJob, nd_passes.h).  blockIdx.x = group of kSpikeCols right-hand sides,
// blockIdx.y = job.  Thread (x, col) owns component x of block row k for one right-hand side; row k of L (its blocks are
// contiguous in the envelope) is staged by cp.async one row ahead, the last 32 block rows of Z live in shared memory
// (column height <= 30).  Per row and CTA: ~30 x (3 LDS.128 + 6 LDS.64 + 6 DFMA) per thread, two block barriers.
constexpr int kSpikeCols = 10;           // right-hand sides per CTA
constexpr int kSpikeThreads = 64;        // 6 x 10 outputs per row (+ 4 idle)
constexpr size_t kSpikeSmem = sizeof(double) * (2 * 31 * 36 + 32 * 6 * kSpikeCols);

__global__ void __launch_bounds__(kSpikeThreads)
nd_spike_kernel(const nd::SpikeJob* __restrict__ jobs) {
  constexpr int CW = kSpikeCols;
  extern __shared__ __align__(16) double smem_spike[];
  double (*sRow)[31 * 36] = reinterpret_cast<double (*)[31 * 36]>(smem_spike);            // [2] blocks of row k (double buffered)
  double (*sZ)[6][CW] = reinterpret_cast<double (*)[6][CW]>(smem_spike + 2 * 31 * 36);   // [32] the last 32 rows of Z of this group
  const nd::SpikeJob J = jobs[blockIdx.y];
  const EnvView e = J.e;
  const int c0 = blockIdx.x * CW;
  if (c0 >= J.KS) return;
  const int tid = threadIdx.x;
  const int x = tid / CW, col = tid - x * CW;               // output (component x of the row, right-hand side c0+col)
  const bool act = tid < 6 * CW && c0 + col < J.KS;
  auto stage_row = [&](int k) {                             // blocks (k, f .. min(k, n_stop)-1) -> sRow[k & 1]
    if (k >= e.n) return;
    const int f = e.first[k];
    const int jend = k < J.n_stop ? k : J.n_stop;
    const int nb = jend > f ? jend - f : 0;
    const double* src = J.L + e.row_start[k] * 36;
    for (int o = tid; o < nb * 18; o += kSpikeThreads) cp_async16_zfill(&sRow[k & 1][2 * o], src + 2 * o, true);
  };
  stage_row(0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int k = 0; k < e.n; ++k) {
    stage_row(k + 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                        // row k staged; Z rows < k complete
    const int f = e.first[k];
    const int jend = k < J.n_stop ? k : J.n_stop;
    if (act) {
      double acc = (k < J.nE) ? J.E[((long long)k * 6 + x) * J.KS + c0 + col] : 0.0;
      double acc2 = 0.0;
      for (int j = f; j < jend; ++j) {
        const double* b = &sRow[k & 1][(j - f) * 36 + x * 6];
        const double (*zj)[CW] = sZ[j & 31];
        acc -= b[0] * zj[0][col] + b[2] * zj[2][col] + b[4] * zj[4][col];
        acc2 += b[1] * zj[1][col] + b[3] * zj[3][col] + b[5] * zj[5][col];
      }
      acc -= acc2;
      J.Z[((long long)k * 6 + x) * J.KS + c0 + col] = acc;
      sZ[k & 31][x][col] = acc;                             // row k-32 is no longer needed (column height <= 30)
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SYRK: U -= sum_k Z_k^T K_k Z_k, u -= sum_k Z_k^T K_k w_k over the pivot rows of a node (SyrkSeg, nd_passes.h).
// grid = (tiles of 30 x 30 scalars of the lower triangle, chunks of kSyrkRows rows, segments); partial sums of the row
// chunks meet in U by RED.ADD.F64.
constexpr int kSyrkTile = 30;            // scalar columns per tile side (5 blocks)
constexpr int kSyrkRows = 32;            // rows of Z per CTA

__global__ void __launch_bounds__(256)
nd_syrk_kernel(const nd::SyrkSeg* __restrict__ segs) {
  constexpr int TW = kSyrkTile;
  __shared__ double sKZ[6][TW];          // (K_k Z_k)[p][a] for the tile's row side
  __shared__ double sZb[6][TW];          // Z_k[p][b] for the tile's column side
  __shared__ double sKw[6];
  const nd::SyrkSeg G = segs[blockIdx.z];
  int ti = 0, tj = 0;                    // tile (ti, tj), tj <= ti, from the linear index
  { int t = blockIdx.x; while ((ti + 1) * (ti + 2) / 2 <= t) ++ti; tj = t - ti * (ti + 1) / 2; }
  if (ti * TW >= G.KS) return;
  const int k0 = blockIdx.y * kSyrkRows;
  if (k0 >= G.rows) return;
  const int k1 = min(G.rows, k0 + kSyrkRows);
  const int tid = threadIdx.x;
  double acc[4] = {0, 0, 0, 0};          // outputs (a, b) of the TW x TW tile: 900 over 256 threads
  double racc = 0.0;
  for (int k = k0; k < k1; ++k) {
    __syncthreads();
    if (tid < 6 * TW) {
      const int p = tid / TW, a = tid - p * TW;
      const double* Kk = G.K + (long long)k * 36 + p * 6;
      const double* Zk = G.Z + (long long)k * 6 * G.KS;
      double s = 0.0;
      const bool ina = ti * TW + a < G.KS, inb = tj * TW + a < G.KS;
      if (ina) {
#pragma unroll
        for (int q = 0; q < 6; ++q) s += Kk[q] * Zk[(long long)q * G.KS + ti * TW + a];
      }
      sKZ[p][a] = s;
      sZb[p][a] = inb ? Zk[(long long)p * G.KS + tj * TW + a] : 0.0;
    } else if (tid < 6 * TW + 6) {
      const int p = tid - 6 * TW;
      const double* Kk = G.K + (long long)k * 36 + p * 6;
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) s += Kk[q] * G.w[(long long)k * 6 + q];
      sKw[p] = s;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = tid + 256 * u;
      if (o < TW * TW) {
        const int a = o / TW, b = o - a * TW;
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < 6; ++p) s += sKZ[p][a] * sZb[p][b];
        acc[u] += s;
      }
    }
    if (tj == 0 && tid < TW && ti * TW + tid < G.KS) {           // u rows of tile row ti: sum_p Z_k[p][a] (K w)[p]
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < 6; ++p) s += G.Z[((long long)k * 6 + p) * G.KS + ti * TW + tid] * sKw[p];
      racc += s;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = tid + 256 * u;
    if (o < TW * TW) {
      const int a = o / TW, b = o - a * TW;
      const int ga = ti * TW + a, gb = tj * TW + b;             // scalar row / column inside the boundary
      const int bi = ga / 6, bj = gb / 6;
      if (bj <= bi && ga < G.KS && gb < G.KS) atomicAdd(&G.U[((long long)bi * (bi + 1) / 2 + bj) * 36 + (ga - 6 * bi) * 6 + (gb - 6 * bj)], -acc[u]);
    }
  }
  if (tj == 0 && tid < TW && ti * TW + tid < G.KS) atomicAdd(&G.u[ti * TW + tid], -racc);
}

}  // namespace lvba
