// four_chunk.cuh — EXPERIMENTAL (development harness only; not part of liblvba_b200.so).
//
// Next step of the block-envelope LDL^T solve (algebra checked in tools/proto/chunked_ldl.py):
// the pose system is split at a middle separator S (<= 30 block rows) into a top part T and a bottom part B; T and B
// are factorised independently, each by the existing twisted scheme (four chunk factorisations run as four CTAs of ONE
// launch), and S is eliminated last:
//       T x_T + F_T^T x_S = r_T          F_T = A[S, T]  (non-zero only in T's last <= 30 block columns)
//       F_T x_T + S x_S + F_B^T x_B = r_S
//       B x_B + F_B x_S = r_B            F_B = A[B, S]  (non-zero only in B's first <= 30 block rows)
//   S' = S - Z_T^T D_T^-1 Z_T - Z_B^T D_B^-1 Z_B,   Z_X = L_X^-1 P_X F_X^T   (the "spike", forward substitution only)
//   r_S' = r_S - Z_T^T D_T^-1 w_T - Z_B^T D_B^-1 w_B, w_X = L_X^-1 P_X r_X    (already produced by the factor kernels)
//   x_S = S'^-1 r_S' ;  w_X <- w_X - Z_X x_S ;  x_X = P_X^T L_X^-T D_X^-1 w_X  (the existing backward substitutions)
// In the twisted order of T the rows next to S are eliminated FIRST by T's reversed chunk, so Z_T is non-zero on that
// chunk and on T's inner separator only (mirror image for B: its natural-order chunk).  The spike has 6*|S| = 180
// right-hand sides: twice the flops of the trailing update per pivot column, but it needs nothing but finished columns
// of L, so it runs on other SMs (one CTA per 30 right-hand sides).
#pragma once
#include "../../global-lvba_b200/csrc/runtime.cuh"

namespace lvba {

// First measurement (profiles/r01_four_chunk_lab.txt) ran with 30 right-hand sides per CTA (6 CTAs per job): every output
// issues 12 LDS per 6 FMA, ~2 000 warp-LDS per row and CTA, i.e. the spike alone took ~0.6 ms.  Ten right-hand sides per CTA
// (18 CTAs per job, 36 SMs for the two spikes) cut the per-CTA shared-memory work by three; UNMEASURED so far.
constexpr int kSpikeCols = 10;           // right-hand sides per spike CTA
constexpr int kSpikeThreads = 64;        // 6 x 10 outputs per row (+ 4 idle)
constexpr size_t kSpikeSmem = sizeof(double) * (2 * 31 * 36 + 32 * 6 * kSpikeCols);

// Forward substitution of KS right-hand sides through the unit-lower block factor of one factorisation instance.
//   rows k < n_stop (pivots):      Z_k = E_k - sum_{j<k} L_kj Z_j
//   rows k >= n_stop (trailing):   Z_k = E_k - sum_{j<n_stop} L_kj Z_j      (what the separator inherits)
// E has nE block rows ([nE][6][KS], zero beyond); Z is [e.n][6][KS].  blockIdx.x = column group, blockIdx.y = job.
struct SpikeJob {
  EnvView e;
  const double* L;
  int n_stop;
  const double* E;
  int nE;
  double* Z;
  int KS;
};

__global__ void __launch_bounds__(kSpikeThreads, 1)
env_spike_kernel(const SpikeJob* __restrict__ jobs) {
  constexpr int CW = kSpikeCols;
  extern __shared__ __align__(16) double smem_spike[];
  double (*sRow)[31 * 36] = reinterpret_cast<double (*)[31 * 36]>(smem_spike);            // [2] blocks of row k (double buffered)
  double (*sZ)[6][CW] = reinterpret_cast<double (*)[6][CW]>(smem_spike + 2 * 31 * 36);   // [32] the last 32 rows of Z for this column group
  const SpikeJob J = jobs[blockIdx.y];
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

// S' -= sum_k Z_k^T K_k Z_k  and  r' -= sum_k Z_k^T K_k w_k  over a segment of rows (K_k = D_k^-1, 36 doubles each).
// Output: lower block triangle of the separator in envelope layout (dense: block (i,j), j <= i, at (i(i+1)/2+j)*36),
// accumulated with atomics.  grid = (tiles of 5x5 blocks of the lower triangle, row chunks, segments).
struct SyrkSeg {
  const double* Z;     // [rows][6][KS]
  const double* K;     // [rows][36]
  const double* w;     // [rows][6]
  int rows;
  int KS;
};
constexpr int kSyrkTile = 30;            // scalar columns per tile side (5 blocks)
constexpr int kSyrkRows = 32;            // rows of Z per CTA

__global__ void __launch_bounds__(256, 1)
env_syrk_kernel(const SyrkSeg* __restrict__ segs, int ntile, double* __restrict__ Ssep, double* __restrict__ rsep) {
  constexpr int TW = kSyrkTile;
  __shared__ double sKZ[6][TW];          // (K_k Z_k)[p][a] for the tile's row side
  __shared__ double sZb[6][TW];          // Z_k[p][b] for the tile's column side
  __shared__ double sKw[6];
  const SyrkSeg G = segs[blockIdx.z];
  int ti = 0, tj = 0;                    // tile (ti, tj), tj <= ti, from the linear index
  { int t = blockIdx.x; while ((ti + 1) * (ti + 2) / 2 <= t) ++ti; tj = t - ti * (ti + 1) / 2; }
  (void)ntile;
  const int k0 = blockIdx.y * kSyrkRows;
  if (k0 >= G.rows) return;
  const int k1 = min(G.rows, k0 + kSyrkRows);
  const int tid = threadIdx.x;
  // outputs of this thread: (a, b) pairs of the TW x TW tile, 900 outputs over 256 threads
  double acc[4] = {0, 0, 0, 0};
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
    if (tj == 0 && tid < TW && ti * TW + tid < G.KS) {           // r' rows of tile row ti: sum_p Z_k[p][a] (K w)[p]
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
      const int ga = ti * TW + a, gb = tj * TW + b;             // scalar row / column inside the separator
      const int bi = ga / 6, bj = gb / 6;
      if (bj <= bi && ga < G.KS && gb < G.KS) atomicAdd(&Ssep[((long long)bi * (bi + 1) / 2 + bj) * 36 + (ga - 6 * bi) * 6 + (gb - 6 * bj)], -acc[u]);
    }
  }
  if (tj == 0 && tid < TW && ti * TW + tid < G.KS) atomicAdd(&rsep[ti * TW + tid], -racc);
}

// w_k -= Z_k x_S over a segment
__global__ void env_spike_correct_kernel(const double* __restrict__ Z, int rows, int KS, const double* __restrict__ xs, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // (row, component)
  if (i >= rows * 6) return;
  const double* zr = Z + (long long)i * KS;
  double s = 0.0;
  for (int c = 0; c < KS; ++c) s += zr[c] * xs[c];
  w[i] -= s;
}

// sub-matrix rows/cols >= off of the envelope `eo` -> envelope `eb` (first[] clipped and re-based)
__global__ void env_gather_sub_kernel(EnvView eo, EnvView eb, int off, const double* __restrict__ Ho, double* __restrict__ Hb) {
  for (int rp = blockIdx.x; rp < eb.n; rp += gridDim.x) {
    const int r = rp + off, f = eb.first[rp];
    const long long dst = eb.row_start[rp] * 36;
    const long long src = (eo.row_start[r] + (f + off - eo.first[r])) * 36;
    for (int o = threadIdx.x; o < (rp - f + 1) * 36; o += blockDim.x) Hb[dst + o] = Ho[src + o];
  }
}

// E_T[rp][x][6 si + y] = H(m+si, m-1-rp)[y][x]   (top part: reversed chunk, rp < nE)
// E_B[r'][x][6 si + y] = H(m+s+r', m+si)[x][y]   (bottom part: natural-order chunk, r' < nE)
__global__ void env_spike_rhs_kernel(EnvView eo, int m, int s, int nE, const double* __restrict__ H, double* __restrict__ ET, double* __restrict__ EB) {
  const int KS = 6 * s;
  const int total = nE * 6 * KS;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < 2 * total; o += gridDim.x * blockDim.x) {
    const bool bot = o >= total;
    const int oo = bot ? o - total : o;
    const int r = oo / (6 * KS), rem = oo - r * 6 * KS, x = rem / KS, c = rem - x * KS, si = c / 6, y = c - 6 * si;
    double v = 0.0;
    if (!bot) {
      const int i = m + si, j = m - 1 - r;
      if (j >= 0 && j >= eo.first[i]) v = H[(eo.row_start[i] + (j - eo.first[i])) * 36 + y * 6 + x];
      ET[oo] = v;
    } else {
      const int i = m + s + r, j = m + si;
      if (i < eo.n && j >= eo.first[i]) v = H[(eo.row_start[i] + (j - eo.first[i])) * 36 + x * 6 + y];
      EB[oo] = v;
    }
  }
}

// outer separator: lower dense envelope of A[S,S] (+ damping on the scalar diagonal) and its right-hand side
__global__ void env_sep_assemble_kernel(EnvView eo, int m, int s, const double* __restrict__ H, const double* __restrict__ dadd,
                                        const double* __restrict__ rhs, double* __restrict__ Ssep, double* __restrict__ rsep) {
  const int nblk = s * (s + 1) / 2;
  for (int o = threadIdx.x; o < nblk * 36; o += blockDim.x) {
    const int blk = o / 36, el = o - blk * 36;
    int si, sj;
    tri_decode(blk, si, sj);
    const int i = m + si, j = m + sj;
    double v = (j >= eo.first[i]) ? H[(eo.row_start[i] + (j - eo.first[i])) * 36 + el] : 0.0;
    if (si == sj && el / 6 == el % 6) v += dadd[6 * i + el / 6];
    Ssep[o] = v;
  }
  for (int o = threadIdx.x; o < s * 6; o += blockDim.x) rsep[o] = rhs[6 * m + o];
}

// inner separator right-hand sides of the spike: T: rows of the reversed chunk's trailing part, reversed again;
// B: rows of the natural-order chunk's trailing part as they are
__global__ void env_spike_sep_rhs_kernel(const double* __restrict__ Zchunk, int n_stop, int bs, int KS, int reversed, double* __restrict__ Esep) {
  const int total = bs * 6 * KS;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
    const int si = o / (6 * KS), rem = o - si * 6 * KS;
    const int srow = reversed ? (bs - 1 - si) : si;
    Esep[o] = Zchunk[((long long)(n_stop + srow) * 6) * KS + rem];
  }
}

struct ChunkedSolver {
  int n = 0, m = 0, s = 0, nB = 0, KS = 0;
  Envelope envT, envB, envS;
  EnvSolver solT, solB, solS;
  DevBuf<double> HB, ET, EB, ZT, ZB, ZTs, ZBs, ETs, EBs, Ssep, xS, xTb;
  DevBuf<FactorJob> fj4, fjs2;
  DevBuf<BacksolveJob> bj4, bjs2;
  DevBuf<SpikeJob> sj2, sjs2;
  DevBuf<SyrkSeg> segs;
  bool ready = false;
  const double* jobs_x = nullptr;

  // returns LVBA_ERR_UNSUPPORTED when the system is too small / too wide for the four-chunk split
  int prepare(const Envelope& env, cudaStream_t st) {
    n = env.n;
    if (env.max_col > 30 || n < 1024) return fail(LVBA_ERR_UNSUPPORTED, "four-chunk split needs n >= 1024 and column height <= 30");
    m = n / 2;
    const int send = env.last[m - 1] + 1;
    s = send - m;
    nB = n - send;
    if (s < 3 || s > 30 || nB < 256 || m < 256) return fail(LVBA_ERR_UNSUPPORTED, "no usable middle separator");
    KS = 6 * s;
    int64_t dummy = 0;
    std::vector<int> fT(env.first.begin(), env.first.begin() + m);
    LVBA_TRY(envT.build(fT, st, &dummy));
    std::vector<int> fB((size_t)nB);
    for (int r = 0; r < nB; ++r) fB[r] = std::max(env.first[send + r], send) - send;
    LVBA_TRY(envB.build(fB, st, &dummy));
    std::vector<int> fS((size_t)s, 0);
    LVBA_TRY(envS.build(fS, st, &dummy));
    LVBA_TRY(solT.prepare(envT, st));
    LVBA_TRY(solB.prepare(envB, st));
    LVBA_TRY(solS.prepare(envS, st));
    if (!solT.tw || !solB.tw) return fail(LVBA_ERR_UNSUPPORTED, "inner twisted split not available");
    LVBA_TRY(HB.alloc((size_t)envB.nblocks * 36));
    const int nE = 30;
    LVBA_TRY(ET.alloc((size_t)nE * 6 * KS)); LVBA_TRY(EB.alloc((size_t)nE * 6 * KS));
    LVBA_TRY(ZT.alloc((size_t)solT.tw_nb * 6 * KS)); LVBA_TRY(ZB.alloc((size_t)solB.tw_send * 6 * KS));
    LVBA_TRY(ZTs.alloc((size_t)solT.tw_bs * 6 * KS)); LVBA_TRY(ZBs.alloc((size_t)solB.tw_bs * 6 * KS));
    LVBA_TRY(ETs.alloc((size_t)solT.tw_bs * 6 * KS)); LVBA_TRY(EBs.alloc((size_t)solB.tw_bs * 6 * KS));
    LVBA_TRY(Ssep.alloc((size_t)envS.nblocks * 36)); LVBA_TRY(xS.alloc((size_t)KS));
    LVBA_CUDA(cudaFuncSetAttribute(env_spike_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpikeSmem));
    LVBA_CUDA(cudaStreamSynchronize(st));
    ready = true; jobs_x = nullptr;
    return LVBA_OK;
  }

  int build_tables(double* x, cudaStream_t st) {
    if (jobs_x == x) return LVBA_OK;
    LVBA_TRY(solT.build_jobs(envT, x, st));
    LVBA_TRY(solB.build_jobs(envB, x + 6 * (size_t)(m + s), st));
    LVBA_TRY(fj4.alloc(4)); LVBA_TRY(fjs2.alloc(2)); LVBA_TRY(bj4.alloc(4)); LVBA_TRY(bjs2.alloc(2));
    auto d2d = [&](void* d, const void* s_, size_t b) { return cudaMemcpyAsync(d, s_, b, cudaMemcpyDeviceToDevice, st); };
    LVBA_CUDA(d2d(fj4.p, solT.d_fjobs.p, 2 * sizeof(FactorJob)));
    LVBA_CUDA(d2d(fj4.p + 2, solB.d_fjobs.p, 2 * sizeof(FactorJob)));
    LVBA_CUDA(d2d(fjs2.p, solT.d_fjobs.p + 2, sizeof(FactorJob)));
    LVBA_CUDA(d2d(fjs2.p + 1, solB.d_fjobs.p + 2, sizeof(FactorJob)));
    LVBA_CUDA(d2d(bjs2.p, solT.d_bjobs.p, sizeof(BacksolveJob)));
    LVBA_CUDA(d2d(bjs2.p + 1, solB.d_bjobs.p, sizeof(BacksolveJob)));
    LVBA_CUDA(d2d(bj4.p, solT.d_bjobs.p + 1, 2 * sizeof(BacksolveJob)));
    LVBA_CUDA(d2d(bj4.p + 2, solB.d_bjobs.p + 1, 2 * sizeof(BacksolveJob)));
    // spike through T's reversed chunk and B's natural-order chunk, then through the two inner separators
    EnvView vBt = envB.view(); vBt.n = solB.tw_send;
    std::vector<SpikeJob> sj = {SpikeJob{solT.env_bot.view(), solT.Lbot.p, solT.tw_nbstop, ET.p, 30, ZT.p, KS},
                                SpikeJob{vBt, solB.L.p, solB.tw_m, EB.p, 30, ZB.p, KS}};
    std::vector<SpikeJob> sjs = {SpikeJob{solT.env_sep.view(), solT.Lsep.p, solT.tw_bs, ETs.p, solT.tw_bs, ZTs.p, KS},
                                 SpikeJob{solB.env_sep.view(), solB.Lsep.p, solB.tw_bs, EBs.p, solB.tw_bs, ZBs.p, KS}};
    std::vector<SyrkSeg> sg = {SyrkSeg{ZT.p, solT.dinv_bot.p, solT.zbot.p, solT.tw_nbstop, KS}, SyrkSeg{ZTs.p, solT.dinv_sep.p, solT.zsep.p, solT.tw_bs, KS},
                               SyrkSeg{ZB.p, solB.dinv.p, solB.z.p, solB.tw_m, KS}, SyrkSeg{ZBs.p, solB.dinv_sep.p, solB.zsep.p, solB.tw_bs, KS}};
    LVBA_TRY(sj2.upload(sj, st)); LVBA_TRY(sjs2.upload(sjs, st)); LVBA_TRY(segs.upload(sg, st));
    LVBA_CUDA(cudaStreamSynchronize(st));
    jobs_x = x;
    return LVBA_OK;
  }

  // Solves (H + diag(dadd)) x = rhs.  rhs: [6n] device vector (left untouched).
  int solve(const Envelope& env, const double* H, const double* dadd, const double* rhs, double* x, cudaStream_t st, int64_t* launches) {
    if (!ready) return fail(LVBA_ERR_INVALID_ARG, "ChunkedSolver not prepared");
    LVBA_TRY(build_tables(x, st));
    const EnvView v = env.view(), vT = envT.view(), vB = envB.view(), vS = envS.view();
    const int send = m + s;
    // ---- matrices and right-hand sides of the two parts
    LVBA_CUDA(cudaMemcpyAsync(solT.L.p, H, (size_t)envT.nblocks * 36 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    env_gather_sub_kernel<<<std::min(nB, 2048), 128, 0, st>>>(v, vB, send, H, solB.L.p);
    env_add_diag_kernel<<<(6 * m + 255) / 256, 256, 0, st>>>(vT, dadd, solT.L.p);
    env_add_diag_kernel<<<(6 * nB + 255) / 256, 256, 0, st>>>(vB, dadd + 6 * (size_t)send, solB.L.p);
    LVBA_CUDA(cudaMemcpyAsync(solT.z.p, rhs, (size_t)6 * m * sizeof(double), cudaMemcpyDeviceToDevice, st));
    LVBA_CUDA(cudaMemcpyAsync(solB.z.p, rhs + 6 * (size_t)send, (size_t)6 * nB * sizeof(double), cudaMemcpyDeviceToDevice, st));
    LVBA_CUDA(cudaMemsetAsync(solT.status.p, 0, solT.status.n * sizeof(int), st));
    LVBA_CUDA(cudaMemsetAsync(solB.status.p, 0, solB.status.n * sizeof(int), st));
    // ---- four chunk factorisations in one launch, then the two inner separators
    env_reverse_gather_kernel<<<std::min(solT.tw_nb, 2048), 128, 0, st>>>(vT, solT.env_bot.view(), solT.L.p, solT.Lbot.p, solT.z.p, solT.zbot.p);
    env_reverse_gather_kernel<<<std::min(solB.tw_nb, 2048), 128, 0, st>>>(vB, solB.env_bot.view(), solB.L.p, solB.Lbot.p, solB.z.p, solB.zbot.p);
    const int mc = std::max(std::max(envT.max_col, solT.env_bot.max_col), std::max(envB.max_col, solB.env_bot.max_col));
    LVBA_TRY(solT.launch_factor(EnvSolver::pid(mc), 4, fj4.p, st, launches));
    env_twist_combine_kernel<<<1, 1024, 0, st>>>(vT, solT.tw_m, solT.tw_bs, solT.L.p, solT.z.p, solT.wtop.p, solT.wbot.p, solT.ztopd.p, solT.zbotd.p, solT.Lsep.p, solT.zsep.p);
    env_twist_combine_kernel<<<1, 1024, 0, st>>>(vB, solB.tw_m, solB.tw_bs, solB.L.p, solB.z.p, solB.wtop.p, solB.wbot.p, solB.ztopd.p, solB.zbotd.p, solB.Lsep.p, solB.zsep.p);
    LVBA_TRY(solT.launch_factor(EnvSolver::pid(std::max(solT.env_sep.max_col, solB.env_sep.max_col)), 2, fjs2.p, st, launches));
    // ---- spike
    env_spike_rhs_kernel<<<64, 256, 0, st>>>(v, m, s, 30, H, ET.p, EB.p);
    const dim3 gsp((KS + kSpikeCols - 1) / kSpikeCols, 2);
    env_spike_kernel<<<gsp, kSpikeThreads, kSpikeSmem, st>>>(sj2.p);
    env_spike_sep_rhs_kernel<<<32, 256, 0, st>>>(ZT.p, solT.tw_nbstop, solT.tw_bs, KS, 1, ETs.p);
    env_spike_sep_rhs_kernel<<<32, 256, 0, st>>>(ZB.p, solB.tw_m, solB.tw_bs, KS, 0, EBs.p);
    env_spike_kernel<<<gsp, kSpikeThreads, kSpikeSmem, st>>>(sjs2.p);
    // ---- outer separator
    env_sep_assemble_kernel<<<1, 1024, 0, st>>>(v, m, s, H, dadd, rhs, Ssep.p, solS.z.p);
    const int ntile1 = (KS + kSyrkTile - 1) / kSyrkTile, ntile = ntile1 * (ntile1 + 1) / 2;
    const int maxrows = std::max(std::max(solT.tw_nbstop, solB.tw_m), std::max(solT.tw_bs, solB.tw_bs));
    env_syrk_kernel<<<dim3(ntile, (maxrows + kSyrkRows - 1) / kSyrkRows, 4), 256, 0, st>>>(segs.p, ntile, Ssep.p, solS.z.p);
    // no damping on the Schur complement
    LVBA_TRY(xS.zero(st));
    if (!zero_dadd.p) { LVBA_TRY(zero_dadd.alloc((size_t)KS)); LVBA_TRY(zero_dadd.zero(st)); }
    LVBA_TRY(solS.solve(envS, Ssep.p, zero_dadd.p, xS.p, st, launches));
    // ---- corrections of the forward-substituted right-hand sides, then the existing backward substitutions
    env_spike_correct_kernel<<<(solT.tw_nbstop * 6 + 127) / 128, 128, 0, st>>>(ZT.p, solT.tw_nbstop, KS, xS.p, solT.zbot.p);
    env_spike_correct_kernel<<<(solT.tw_bs * 6 + 127) / 128, 128, 0, st>>>(ZTs.p, solT.tw_bs, KS, xS.p, solT.zsep.p);
    env_spike_correct_kernel<<<(solB.tw_m * 6 + 127) / 128, 128, 0, st>>>(ZB.p, solB.tw_m, KS, xS.p, solB.z.p);
    env_spike_correct_kernel<<<(solB.tw_bs * 6 + 127) / 128, 128, 0, st>>>(ZBs.p, solB.tw_bs, KS, xS.p, solB.zsep.p);
    double* xT = x;
    double* xB = x + 6 * (size_t)send;
    solT.launch_apply(solT.tw_bs, solT.dinv_sep.p, solT.zsep.p, solT.xsep.p, st);
    solB.launch_apply(solB.tw_bs, solB.dinv_sep.p, solB.zsep.p, solB.xsep.p, st);
    solT.launch_backsolve(2, bjs2.p, st);
    solT.launch_apply(solT.tw_m, solT.dinv.p, solT.z.p, xT, st);
    solT.launch_apply(solT.tw_nbstop, solT.dinv_bot.p, solT.zbot.p, solT.xbot.p, st);
    solB.launch_apply(solB.tw_m, solB.dinv.p, solB.z.p, xB, st);
    solB.launch_apply(solB.tw_nbstop, solB.dinv_bot.p, solB.zbot.p, solB.xbot.p, st);
    env_twist_place_sep_kernel<<<(solT.tw_bs * 6 + 127) / 128, 128, 0, st>>>(solT.tw_m, solT.tw_bs, solT.tw_nbstop, solT.xsep.p, xT, solT.xbot.p);
    env_twist_place_sep_kernel<<<(solB.tw_bs * 6 + 127) / 128, 128, 0, st>>>(solB.tw_m, solB.tw_bs, solB.tw_nbstop, solB.xsep.p, xB, solB.xbot.p);
    solT.launch_backsolve(4, bj4.p, st);
    env_twist_scatter_kernel<<<(solT.tw_nbstop * 6 + 255) / 256, 256, 0, st>>>(m, solT.tw_nbstop, solT.xbot.p, xT);
    env_twist_scatter_kernel<<<(solB.tw_nbstop * 6 + 255) / 256, 256, 0, st>>>(nB, solB.tw_nbstop, solB.xbot.p, xB);
    LVBA_CUDA(cudaMemcpyAsync(x + 6 * (size_t)m, xS.p, (size_t)KS * sizeof(double), cudaMemcpyDeviceToDevice, st));
    *launches += 30;
    LVBA_CUDA(cudaGetLastError());
    return LVBA_OK;
  }
  DevBuf<double> zero_dadd;
};

}  // namespace lvba
