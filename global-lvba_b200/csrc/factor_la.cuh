// factor_la.cuh — register-window block LDL^T with a panel look-ahead group ("v7").
//
// Same job as SimplicialLDLT in BALM2::damping_iter (reference include/BALM/bavoxel.hpp:695-710, lower
// triangle, no pivoting — SURVEY.md Q4/Q5) and as the DENSE_SCHUR Cholesky of the reduced camera system
// inside ceres::Solve (src/lvba_system.cpp:1574,1643), for a block envelope whose column height is < P.
//
// One CTA walks the pivot columns.  Every live 6x6 block of the trailing window (rows/cols k..k+P-1) lives in
// the registers of one PAIR thread for its whole life (thread <-> unordered slot pair {a,b}, slot = row mod P).
// What is new against the first register-window kernel:
//   * symmetric update form.  The stored block G always has its rows on slot a and its columns on slot b,
//     whichever of the two is the lower row: A_ab -= L_a D L_b^T = L_a T_b^T holds for both orientations, so a
//     thread reads the SAME two shared-memory operands (L of slot a, T of slot b) at every step and the 32 lanes
//     of a warp (an 8x4 tile of the slot triangle) touch 8 + 4 distinct operands: one wavefront per LDS.128;
//   * the look-ahead group owns the NEXT pivot column.  During step k it applies column k to column k+1 itself
//     (30 block updates), factorises D_{k+1} (6x6 LDL^T in registers), scales the column and publishes
//     L_{.,k+1} / T_{.,k+1} for step k+1, all overlapped with the pair threads' trailing update of step k.  The
//     pair threads therefore never wait for a scale phase: ONE block barrier per pivot column (was two, with the
//     scale serialised between them);
//   * the entering row is streamed global -> shared with cp.async (no staging registers).
#pragma once
#include "envelope.cuh"

namespace lvba {

template <int P>
struct LaCfg {
  static constexpr int kPairs = P * (P + 1) / 2;
  static constexpr int kPairGroups = (kPairs + 127) / 128;       // warpgroups (4 warps) of pair threads
  static constexpr int kPairThreads = kPairGroups * 128;
  static constexpr int kLaThreads = 128;                         // one look-ahead warpgroup
  static constexpr int kThreads = kPairThreads + kLaThreads;
  static constexpr int kItemThreads = 96;                        // look-ahead warps 1..3
  static constexpr int kItems = (P - 1) * 6;                     // (row of the next column, block row x)
  static constexpr int kRounds = (kItems + kItemThreads - 1) / kItemThreads;
  static constexpr int S = 38;                                   // doubles per transposed block (16 B aligned, conflict-free over 8 slots)
  static constexpr bool kRealloc = kThreads > 512;               // P = 31: 640 threads launch at 96 registers
  static constexpr int kPairRegs = 104, kLaRegs = 64;            // 512*104 + 128*64 == 640*96
  static constexpr int kDoubles = 6 * P * S + 2 * P * 36 + 2 * 36 + 36 + 2 * 24 + P * 6;
  static constexpr size_t kSmem = sizeof(double) * (size_t)kDoubles + sizeof(long long) * P + sizeof(int) * (P + 4) + 32;
};

LVBA_DEV void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
LVBA_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
LVBA_DEV void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

#ifdef LVBA_LAB
__device__ int g_la_mode = 0;      // solver_lab only: 1 = pair threads skip their update, 2 = look-ahead group skips its math
#endif

template <int P, bool kLaFirst>
__global__ void __launch_bounds__(LaCfg<P>::kThreads, 1)
env_factor_la_kernel(FactorJobs jobs, const unsigned short* __restrict__ pair_map, int* __restrict__ status,
                     long long* __restrict__ dbg_all) {
  using Cfg = LaCfg<P>;
  constexpr int S = Cfg::S;
  const FactorJob& J = jobs.j[blockIdx.x];
  const EnvView e = J.e;
  double* __restrict__ L = J.L;
  double* __restrict__ dinv = J.dinv;
  double* __restrict__ z = J.z;
  const int n_stop = J.n_stop;
  long long* dbg = (blockIdx.x == 0) ? dbg_all : nullptr;
  // optional phase clocks (LVBA_FACTOR_TIMING=1): dbg[(k*8 + role)*4 + stamp]
#define LVBA_STAMP(role, stamp) do { if (dbg && lane == 0) dbg[((long long)k * 8 + (role)) * 4 + (stamp)] = clock64(); } while (0)

  extern __shared__ __align__(16) double smem_la[];
  double* sL = smem_la;                          // [2][P][S] L_ik transposed ([q*6+x] = L[x][q]); parity = pivot column & 1
  double* sT = sL + 2 * P * S;                   // [2][P][S] T_ik = A_ik (updated, unscaled), same layout
  double* sA = sT + 2 * P * S;                   // [2][P][S] next-next column handed to the look-ahead group, same layout
  double* sEnter = sA + 2 * P * S;               // [2][P][36] entering row by column slot, row-major blocks
  double* sDg = sEnter + 2 * P * 36;             // [2][36]   diagonal block of the next pivot (before column k's update)
  double* sDu = sDg + 72;                        // [36]      updated diagonal block of the next pivot (kept for the window dump)
  double* sF = sDu + 36;                         // [2][24]   packed LDL^T factors of the pivot block
  double* sZ = sF + 48;                          // [P][6]
  long long* sRS = reinterpret_cast<long long*>(sZ + P * 6);   // [P]
  int* sFirst = reinterpret_cast<int*>(sRS + P);               // [P]
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = e.n;
  // warp-uniform roles.  kLaFirst puts the look-ahead warpgroup on the LOWEST warp ids (the warp scheduler's
  // oldest-first tie break then favours its dependent chain over the pair warps' long independent DFMA runs)
  const bool is_la = kLaFirst ? tid < Cfg::kLaThreads : tid >= Cfg::kPairThreads;
  const int pt = kLaFirst ? tid - Cfg::kLaThreads : tid;       // pair thread index
#ifdef LVBA_LAB
  const int lab_mode = g_la_mode;
#endif

  // ---------------- prologue: labels and rhs of the first P rows
  for (int r = tid; r < P; r += Cfg::kThreads) {
    if (r < n) { sFirst[r] = e.first[r]; sRS[r] = e.row_start[r]; } else { sFirst[r] = 0x7fffffff; sRS[r] = 0; }
#pragma unroll
    for (int q = 0; q < 6; ++q) sZ[r * 6 + q] = (r < n) ? z[6 * r + q] : 0.0;
  }
  for (int o = tid; o < 6 * P * S; o += Cfg::kThreads) sL[o] = 0.0;      // sL, sT, sA (padding included)
  __syncthreads();

  if (!is_la) {
    // =================================================== pair threads: one live 6x6 block in registers
    if (Cfg::kRealloc) reg_alloc<Cfg::kPairRegs>();
    const unsigned short pm = pair_map[pt];
    const bool is_pair = pm != 0xffff;
    const int a = pm & 0xff, b = (pm >> 8) & 0xff;             // a >= b
    double G[36];                                              // rows <-> slot a, columns <-> slot b
    auto publish_T = [&](double* dst) {                        // dst[q*6+x] = G[x][q]
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double2* d2 = reinterpret_cast<double2*>(dst + q * 6);
        d2[0] = make_double2(G[q], G[6 + q]);
        d2[1] = make_double2(G[12 + q], G[18 + q]);
        d2[2] = make_double2(G[24 + q], G[30 + q]);
      }
    };
    auto publish_N = [&](double* dst) {                        // dst[q*6+x] = G[q][x]
      double2* d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
      for (int q = 0; q < 18; ++q) d2[q] = make_double2(G[2 * q], G[2 * q + 1]);
    };
    if (is_pair) {
      if (a < n && b >= sFirst[a]) {
        const double2* src = reinterpret_cast<const double2*>(L + (sRS[a] + (b - sFirst[a])) * 36);
#pragma unroll
        for (int q = 0; q < 18; ++q) { const double2 v = src[q]; G[2 * q] = v.x; G[2 * q + 1] = v.y; }
      } else {
#pragma unroll
        for (int q = 0; q < 36; ++q) G[q] = 0.0;
      }
      if (b == 0 && a >= 1) publish_T(sT + a * S);                       // column 0 -> sT[0]
      if (a == 0 && b == 0) publish_N(sDg + 36);                         // pivot 0 (scratch: sDg[1])
      if (P > 1 && b == 1 % P && a >= 2) publish_T(sA + a * S);          // column 1 -> sA[0]
      if (a == 1 % P && b == 1 % P) publish_N(sDg);                      // block (1,1) -> sDg[0]
    }
    __syncthreads();     // (A) columns 0, 1 published
    __syncthreads();     // (B) look-ahead group: F_0, L_{.,0}, entering row P
    int c = 0;
    for (int k = 0; k < n_stop; ++k) {
      const int cur = k & 1;
      const int prole = (pt < 32) ? 0 : (pt >= Cfg::kPairThreads - 32) ? 1 : -1;
      if (prole >= 0) LVBA_STAMP(prole, 0);
      if (is_pair) {
        int da = a - c; if (da < 0) da += P;
        int db = b - c; if (db < 0) db += P;
        const int lo = da < db ? da : db, hi = da < db ? db : da;
        if (lo == 0) {
          // the column-k block is dead: take the entering block (k+P, k+hi) (hi == 0: the diagonal (k+P,k+P))
          const int col_slot = (da == 0) ? b : a;
          const double2* src = reinterpret_cast<const double2*>(sEnter + (cur * P + col_slot) * 36);
          if (da == 0) {                          // slot a is the entering row: G = E
#pragma unroll
            for (int q = 0; q < 18; ++q) { const double2 v = src[q]; G[2 * q] = v.x; G[2 * q + 1] = v.y; }
          } else {                                // slot b is the entering row: G = E^T
#pragma unroll
            for (int x = 0; x < 6; ++x)
#pragma unroll
              for (int y2 = 0; y2 < 3; ++y2) { const double2 v = src[x * 3 + y2]; G[(2 * y2) * 6 + x] = v.x; G[(2 * y2 + 1) * 6 + x] = v.y; }
          }
        } else if (lo >= 2
#ifdef LVBA_LAB
                   && lab_mode != 1
#endif
        ) {
          const double2* lp = reinterpret_cast<const double2*>(sL + (cur * P + a) * S);
          const double2* tp = reinterpret_cast<const double2*>(sT + (cur * P + b) * S);
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const double2 t0 = tp[3 * q], t1 = tp[3 * q + 1], t2 = tp[3 * q + 2];
#pragma unroll
            for (int xx = 0; xx < 3; ++xx) {
              const double2 l = lp[3 * q + xx];
              double* c0 = G + (2 * xx) * 6;
              double* c1 = G + (2 * xx + 1) * 6;
              c0[0] -= l.x * t0.x; c0[1] -= l.x * t0.y; c0[2] -= l.x * t1.x; c0[3] -= l.x * t1.y; c0[4] -= l.x * t2.x; c0[5] -= l.x * t2.y;
              c1[0] -= l.y * t0.x; c1[1] -= l.y * t0.y; c1[2] -= l.y * t1.x; c1[3] -= l.y * t1.y; c1[4] -= l.y * t2.x; c1[5] -= l.y * t2.y;
            }
          }
        }
        if (prole >= 0) LVBA_STAMP(prole, 1);
        // hand column k+2 (the look-ahead group's column of the NEXT step) over: blocks (k+hi, k+2)
        if (lo != 1 && P > 2) {
          double* nA = sA + ((cur ^ 1) * P) * S;
          if (da == 2 && db == 2) publish_N(sDg + (cur ^ 1) * 36);
          else if (db == 2) publish_T(nA + a * S);           // slot a is the row
          else if (da == 2) publish_N(nA + b * S);           // slot b is the row: G = A^T
        }
      }
      if (prole >= 0) LVBA_STAMP(prole, 2);
      __syncthreads();
      if (prole >= 0) LVBA_STAMP(prole, 3);
      if (++c == P) c = 0;
    }
    // partial factorisation: hand the Schur-updated trailing window (rows/cols n_stop..n-1) to the separator solve
    if (n_stop < n && J.wdump && is_pair) {
      int da = a - c; if (da < 0) da += P;
      int db = b - c; if (db < 0) db += P;
      const int lo = da < db ? da : db, hi = da < db ? db : da;
      const int bs = n - n_stop;
      if (hi < bs) {
        double* dst = J.wdump + ((long long)hi * bs + lo) * 36;      // block (n_stop+hi, n_stop+lo), rows on the hi row
        const int fin = n_stop & 1;
        if (lo == 0 && hi == 0) {
          for (int q = 0; q < 36; ++q) { const int i = q / 6, j = q % 6; dst[q] = (i >= j) ? sDu[i * 6 + j] : sDu[j * 6 + i]; }
        } else if (lo == 0) {
          // column n_stop was the look-ahead group's: T_{i,n_stop} sits in sT[fin][slot of the hi row], [q*6+x] = T[x][q]
          const int row_slot = (da == 0) ? b : a;
          const double* t = sT + (fin * P + row_slot) * S;
          for (int q = 0; q < 36; ++q) { const int x = q / 6, y = q % 6; dst[q] = t[y * 6 + x]; }
        } else if (da >= db) {
#pragma unroll
          for (int q = 0; q < 36; ++q) dst[q] = G[q];
        } else {
#pragma unroll
          for (int x = 0; x < 6; ++x)
#pragma unroll
            for (int y = 0; y < 6; ++y) dst[x * 6 + y] = G[y * 6 + x];
        }
      }
    }
  } else {
    // =================================================== look-ahead warpgroup (4 warps, one per SMSP)
    if (Cfg::kRealloc) reg_dealloc<Cfg::kLaRegs>();
    const int lt = kLaFirst ? tid : tid - Cfg::kPairThreads;    // 0..127
    const int aw = lt >> 5;                                     // 0: pivot chain + forward substitution + labels; 1..3: column items + row prefetch
    const int it = lt - 32;                                     // item thread id (0..95), negative on warp 0
    int bad = 0;
    double zin = 0.0;
    int pf_first = 0x7fffffff; long long pf_rs = 0;             // label of the row this thread streams THIS step (fetched a step earlier)

    // row kc+P -> sEnter[kc & 1], by column slot (cp.async, zero-filled outside the envelope / past the last row)
    auto stream_row = [&](int kc, int rf, long long rrs) {
      const int r = kc + P, ck = kc % P;
      double* dstb = sEnter + ((kc & 1) * P) * 36;
      for (int o = it; o < P * 18; o += Cfg::kItemThreads) {
        const int cs = o / 18, w = o - cs * 18;
        int dcol = cs - ck; if (dcol <= 0) dcol += P;            // col = kc + dcol ; dcol == P <=> col == r
        const int col = kc + dcol;
        const bool valid = r < n && col >= rf;
        const double* src = valid ? L + (rrs + (col - rf)) * 36 + 2 * w : L;
        cp_async16_zfill(dstb + 2 * o, src, valid);
      }
    };
    // warp 0: packed factors of the pivot block.  src: 36 row-major, lower triangle read.
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
      for (int q = 0; q < 21; ++q) chk += x[q];
      if (!isfinite(chk)) bad = 1;
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 21; ++q) dst[q] = x[q];
      }
      __syncwarp();
      if (lane < 21 && kc < n_stop) dinv[(long long)kc * 36 + lane] = dst[lane];
    };
    // column item: row slot `slot`, block row x.  t = T_{i,col}[x][.] (already updated); writes L_{i,col}[x][.] = t D^-1
    auto scale_item = [&](const double* F, double (&t)[6], int slot, int x, int col, int par) {
      ldlt_solve6(F, t);
      double* lt_ = sL + (par * P + slot) * S;
#pragma unroll
      for (int q = 0; q < 6; ++q) lt_[q * 6 + x] = t[q];
      const int rf = sFirst[slot];
      if (col >= rf && col < n_stop) {      // a partial factorisation leaves column n_stop (the separator's first) untouched in memory
        double2* g = reinterpret_cast<double2*>(L + (sRS[slot] + (col - rf)) * 36 + x * 6);
        g[0] = make_double2(t[0], t[1]); g[1] = make_double2(t[2], t[3]); g[2] = make_double2(t[4], t[5]);
      }
    };

    __syncthreads();     // (A)
    // ---- column 0: F_0, L_{.,0}; entering row P; labels
    if (aw == 0) {
      factor_pivot(sDg + 36, sF, 0);
      if (lane < 6) zin = (P < n) ? z[6 * (long long)P + lane] : 0.0;
    } else {
      stream_row(0, (P < n) ? e.first[P] : 0x7fffffff, (P < n) ? e.row_start[P] : 0);
      pf_first = (P + 1 < n) ? e.first[P + 1] : 0x7fffffff;
      pf_rs = (P + 1 < n) ? e.row_start[P + 1] : 0;
    }
    named_bar_sync(1, Cfg::kLaThreads);
    if (aw != 0) {
      for (int o = it; o < Cfg::kItems; o += Cfg::kItemThreads) {
        const int h = 1 + o / 6, x = o - (h - 1) * 6;              // row h of column 0, slot h
        double t[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) t[q] = sT[h * S + q * 6 + x];
        scale_item(sF, t, h, x, 0, 0);
      }
      cp_async_wait_all();
    }
    named_bar_sync(1, Cfg::kLaThreads);
    if (aw == 0 && lane == 0) {                                    // slot 0 now describes row P
      sFirst[0] = (P < n) ? e.first[P] : 0x7fffffff;
      sRS[0] = (P < n) ? e.row_start[P] : 0;
    }
    __syncthreads();     // (B)

    int c = 0;
    for (int k = 0; k < n_stop; ++k) {
      const int cur = k & 1, nxt = cur ^ 1;
      int s1 = c + 1; if (s1 >= P) s1 -= P;
      const double* Lc = sL + cur * P * S;
      const double* Tc = sT + cur * P * S;
      const double* T1 = Tc + s1 * S;                               // T_{k+1,k}: [r*6+q] = T[q][r]
      LVBA_STAMP(4 + aw, 0);
      if (aw == 0) {
        // ---- pivot chain: D_{k+1} = A_{k+1,k+1} - L_{k+1,k} T_{k+1,k}^T (lane <-> lower-triangle element), LDL^T
        int m_first = 0x7fffffff; long long m_rs = 0;
        if (lane == 0) {                                            // label of row k+1+P (consumed at the end of the step)
          const int r1 = k + 1 + P;
          m_first = (r1 < n) ? e.first[r1] : 0x7fffffff;
          m_rs = (r1 < n) ? e.row_start[r1] : 0;
        }
        if (k + 1 < n
#ifdef LVBA_LAB
            && lab_mode != 2
#endif
        ) {
          const double* dg = sDg + cur * 36;
          const double* l1 = Lc + s1 * S;
          const int l21 = lane < 21 ? lane : 0;
          const int i = (l21 >= 15) ? 5 : (l21 >= 10) ? 4 : (l21 >= 6) ? 3 : (l21 >= 3) ? 2 : (l21 >= 1) ? 1 : 0;
          const int j = l21 - i * (i + 1) / 2;
          double v = dg[i * 6 + j];
#pragma unroll
          for (int q = 0; q < 6; ++q) v -= l1[q * 6 + i] * T1[q * 6 + j];
          __syncwarp();
          if (lane < 21) sDu[i * 6 + j] = v;
          __syncwarp();
          factor_pivot(sDu, sF + nxt * 24, k + 1);
        }
        LVBA_STAMP(4, 1);
        named_bar_sync(1, Cfg::kLaThreads);                         // F_{k+1} visible to the item warps
        LVBA_STAMP(4, 2);
        // ---- forward substitution with the final z_k : lane <-> row k+1+lane
        double zk[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) zk[q] = sZ[c * 6 + q];
        if (lane < 6) z[6 * (long long)k + lane] = sZ[c * 6 + lane];
        __syncwarp();
        for (int h = 1 + lane; h < P; h += 32) {
          int slot = c + h; if (slot >= P) slot -= P;
          const double2* lt2 = reinterpret_cast<const double2*>(Lc + slot * S);
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
          sZ[c * 6 + lane] = zin;                                   // row k+P takes slot c
          zin = (k + 1 + P < n) ? z[6 * (long long)(k + 1 + P) + lane] : 0.0;
        }
        if (lane == 0) { sFirst[s1] = m_first; sRS[s1] = m_rs; }   // slot of row k+1 now describes row k+1+P
        LVBA_STAMP(4, 3);
      } else {
        // ---- column items: T_{i,k+1} = A_{i,k+1} - L_{i,k} T_{k+1,k}^T for rows i = k+2 .. k+P, then L = T D_{k+1}^-1
        stream_row(k + 1, pf_first, pf_rs);                         // row k+1+P -> sEnter[nxt]
        {
          const int r2 = k + 2 + P;
          pf_first = (r2 < n) ? e.first[r2] : 0x7fffffff;
          pf_rs = (r2 < n) ? e.row_start[r2] : 0;
        }
        double t[Cfg::kRounds][6];
        int slot_[Cfg::kRounds], x_[Cfg::kRounds];
#ifdef LVBA_LAB
        if (lab_mode == 2) {
#pragma unroll
          for (int rd = 0; rd < Cfg::kRounds; ++rd) { slot_[rd] = 0; x_[rd] = 0;
#pragma unroll
            for (int q = 0; q < 6; ++q) t[rd][q] = 0.0; }
        } else
#endif
#pragma unroll
        for (int rd = 0; rd < Cfg::kRounds; ++rd) {
          const int o = it + rd * Cfg::kItemThreads;
          const int oo = o < Cfg::kItems ? o : 0;
          const int h = 2 + oo / 6;                                 // row k+h, h = 2..P
          const int x = oo - (h - 2) * 6;
          int slot = c + h; if (slot >= P) slot -= P;               // h == P -> slot c (the entering row k+P)
          slot_[rd] = slot; x_[rd] = x;
          if (h == P) {
            const double* en = sEnter + (cur * P + s1) * 36 + x * 6;   // block (k+P, k+1), row-major
#pragma unroll
            for (int q = 0; q < 6; ++q) t[rd][q] = en[q];
          } else {
            const double* as = sA + (cur * P + slot) * S;
#pragma unroll
            for (int q = 0; q < 6; ++q) t[rd][q] = as[q * 6 + x];
            const double* lr = Lc + slot * S;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              const double lv = lr[r * 6 + x];
              const double2 u0 = reinterpret_cast<const double2*>(T1 + r * 6)[0];
              const double2 u1 = reinterpret_cast<const double2*>(T1 + r * 6)[1];
              const double2 u2 = reinterpret_cast<const double2*>(T1 + r * 6)[2];
              t[rd][0] -= lv * u0.x; t[rd][1] -= lv * u0.y; t[rd][2] -= lv * u1.x; t[rd][3] -= lv * u1.y; t[rd][4] -= lv * u2.x; t[rd][5] -= lv * u2.y;
            }
          }
          if (o < Cfg::kItems) {
            double* tn = sT + (nxt * P + slot) * S;
#pragma unroll
            for (int q = 0; q < 6; ++q) tn[q * 6 + x] = t[rd][q];
          }
        }
        LVBA_STAMP(4 + aw, 1);
        named_bar_sync(1, Cfg::kLaThreads);                         // F_{k+1} ready
        LVBA_STAMP(4 + aw, 2);
        const double* F = sF + nxt * 24;
#pragma unroll
        for (int rd = 0; rd < Cfg::kRounds; ++rd) {
          const int o = it + rd * Cfg::kItemThreads;
          if (o < Cfg::kItems
#ifdef LVBA_LAB
              && lab_mode != 2
#endif
          ) scale_item(F, t[rd], slot_[rd], x_[rd], k + 1, nxt);
        }
        cp_async_wait_all();
        LVBA_STAMP(4 + aw, 3);
      }
      __syncthreads();
      if (++c == P) c = 0;
    }
    if (n_stop < n && J.zdump && aw == 0) {                         // forward-substituted rhs of the separator rows
      for (int o = lane; o < (n - n_stop) * 6; o += 32) {
        const int i = n_stop + o / 6;
        J.zdump[o] = sZ[(i % P) * 6 + o % 6];
      }
    }
    if (bad) status[0] = 1;
  }
#undef LVBA_STAMP
}

// =====================================================================================================
// Backward substitution  x <- L^-T x  on ONE consumer warp, no block barriers.
//
// Row-oriented: when x_i is final, every block L_ij of row i (columns first[i]..i-1, contiguous in memory) sends
// x_j -= L_ij^T x_i.  Lane s of the consumer warp OWNS x_j for the row j == s (mod 32) of the live 32-row window and
// keeps it in registers; x_i reaches the other lanes by shuffles, so the dependent chain per row is one shuffle +
// a short FMA tree (~100 cycles) instead of two block barriers.  A producer warp streams L through a ring of
// shared-memory stages guarded by full/empty mbarriers: rows are contiguous in memory, so ONE cp.async.bulk brings a
// group of kBsGroup consecutive rows.  Row labels travel in registers (32 rows per coalesced load, one chunk ahead),
// the x entries that enter the window are fetched kBsXDist rows ahead, the next row's block is loaded from shared
// memory while the current row is applied (software pipeline), and the two quarter-warp halves read their blocks in
// an XOR-swizzled 16-byte order so that the 288-byte block stride does not collide on the banks.
constexpr int kBsStages = 4;
constexpr int kBsGroup = 4;                                  // rows per stage (divides 32)
constexpr int kBsStageDoubles = kBsGroup * 31 * 36;
constexpr int kBsXDist = 8;                                  // prefetch distance (rows) of the entering x entries
constexpr size_t kBsSmem = sizeof(double) * (size_t)kBsStages * kBsStageDoubles + 2 * kBsStages * sizeof(unsigned long long) +
                           kBsStages * sizeof(long long) + 64;

LVBA_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
LVBA_DEV void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
LVBA_DEV void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
LVBA_DEV void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LVBA_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
LVBA_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(64, 1)
env_backsolve_warp_kernel(BacksolveJobs jobs) {
  constexpr int NS = kBsStages, R = kBsGroup;
  extern __shared__ __align__(128) double smem_bs[];
  double* ring = smem_bs;                                                       // [NS][kBsStageDoubles]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + NS * kBsStageDoubles);   // [NS]
  unsigned long long* empty = full + NS;                                        // [NS]
  long long* sBase = reinterpret_cast<long long*>(empty + NS);                  // [NS] row_start of the lowest staged row
  const BacksolveJob& J = jobs.j[blockIdx.x];
  const EnvView e = J.e;
  const double* __restrict__ L = J.L;
  double* __restrict__ x = J.x;
  const int n_given = J.n_given, n = e.n;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    for (int d = 0; d < NS; ++d) { mbar_init(full + d, 1); mbar_init(empty + d, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_groups = (n + R - 1) / R;
  if (tid >= 32) {
    // ------------------------------------------------ producer warp: one bulk copy per group of R rows
    int lab_f = 0; long long lab_rs = 0;
    for (int g = 0; g < n_groups; ++g) {
      const int itn = g * R;                         // rows i_hi = n-1-itn down to i_lo
      const int i_hi = n - 1 - itn;
      const int i_lo = (i_hi - R + 1 > 0) ? i_hi - R + 1 : 0;
      if ((itn & 31) == 0) {                         // labels of rows i_hi, i_hi-1, .., i_hi-31 : lane l holds row i_hi-l
        const int r = i_hi - lane;
        lab_f = (r >= 0) ? e.first[r] : 0;
        lab_rs = (r >= 0) ? e.row_start[r] : 0;
      }
      const int f_hi = __shfl_sync(0xffffffffu, lab_f, itn & 31);
      const long long rs_hi = __shfl_sync(0xffffffffu, lab_rs, itn & 31);
      const long long rs_lo = __shfl_sync(0xffffffffu, lab_rs, (itn & 31) + (i_hi - i_lo));
      if (lane == 0) {
        const int st = g % NS;
        const unsigned ph = (unsigned)((g / NS) & 1);
        mbar_wait(empty + st, ph ^ 1u);
        const long long nblk = rs_hi + (i_hi - f_hi + 1) - rs_lo;      // blocks of rows i_lo..i_hi, diagonal blocks included
        sBase[st] = rs_lo;
        mbar_arrive_expect_tx(full + st, (unsigned)(nblk * 288));
        bulk_g2s(ring + st * kBsStageDoubles, L + rs_lo * 36, (unsigned)(nblk * 288), full + st);
      }
    }
  } else {
    // ------------------------------------------------ consumer warp
    double xs[6], xn[6];
    {
      const int r = (n - 1) - (((n - 1) - lane) % 32 + 32) % 32;   // the row == lane (mod 32) inside [n-32, n-1]
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        xs[q] = (r >= 0) ? x[6 * (long long)r + q] : 0.0;
        xn[q] = (r - 32 >= 0) ? x[6 * (long long)(r - 32) + q] : 0.0;
      }
    }
    // row labels: lane l holds row (chunk top - l); the next chunk is fetched one chunk ahead
    int lab_f, lab_f2 = 0; long long lab_rs, lab_rs2 = 0;
    {
      const int r = n - 1 - lane;
      lab_f = (r >= 0) ? e.first[r] : 0;
      lab_rs = (r >= 0) ? e.row_start[r] : 0;
    }
    double2 blk[18];                                   // block of the row being applied (this lane's column)
    bool blk_on = false;
    auto load_block = [&](int i, int itn) {            // stage the block of row i for this lane into blk
      const int f = __shfl_sync(0xffffffffu, lab_f, itn & 31);
      const long long rs = __shfl_sync(0xffffffffu, lab_rs, itn & 31);
      const int g = itn / R, st = g % NS;
      if (itn % R == 0) mbar_wait(full + st, (unsigned)((g / NS) & 1));
      const int upto = i < n_given ? i : n_given;      // given rows only act on the pivots' columns
      const int cnt = upto > f ? upto - f : 0;
      const int jo = (lane - f) & 31;                  // column f + jo is the one congruent to this lane
      blk_on = jo < cnt;
      if (blk_on) {
        const double2* b2 = reinterpret_cast<const double2*>(ring + st * kBsStageDoubles + (rs - sBase[st]) * 36 + jo * 36);
        const int sw = (jo >> 2) & 1;
#pragma unroll
        for (int t = 0; t < 18; ++t) blk[t] = b2[t ^ sw];
        if (sw) {
#pragma unroll
          for (int t = 0; t < 18; t += 2) { const double2 tmp = blk[t]; blk[t] = blk[t + 1]; blk[t + 1] = tmp; }
        }
      }
    };
    if (n > 0) load_block(n - 1, 0);
    for (int i = n - 1; i >= 0; --i) {
      const int itn = n - 1 - i;
      const int owner = i & 31;
      double xi[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) xi[q] = __shfl_sync(0xffffffffu, xs[q], owner);
      if (lane == owner) {
        double2* xo = reinterpret_cast<double2*>(x + 6 * (long long)i);
        xo[0] = make_double2(xs[0], xs[1]); xo[1] = make_double2(xs[2], xs[3]); xo[2] = make_double2(xs[4], xs[5]);
      }
      // apply row i with the block loaded one iteration earlier
      double v0[6], v1[6];
      const bool on = blk_on;
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) { v0[cc] = 0.0; v1[cc] = 0.0; }
      if (on) {
#pragma unroll
        for (int q = 0; q < 6; q += 2) {
          const double2 a0 = blk[3 * q], a1 = blk[3 * q + 1], a2 = blk[3 * q + 2];
          const double2 c0 = blk[3 * q + 3], c1 = blk[3 * q + 4], c2 = blk[3 * q + 5];
          v0[0] += a0.x * xi[q]; v0[1] += a0.y * xi[q]; v0[2] += a1.x * xi[q]; v0[3] += a1.y * xi[q]; v0[4] += a2.x * xi[q]; v0[5] += a2.y * xi[q];
          v1[0] += c0.x * xi[q + 1]; v1[1] += c0.y * xi[q + 1]; v1[2] += c1.x * xi[q + 1]; v1[3] += c1.y * xi[q + 1]; v1[4] += c2.x * xi[q + 1]; v1[5] += c2.y * xi[q + 1];
        }
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) xs[cc] -= (v0[cc] + v1[cc]);
      }
      if (lane == owner) {                             // row i leaves the window, row i-32 takes its lane
#pragma unroll
        for (int q = 0; q < 6; ++q) xs[q] = xn[q];
      }
      // x entry of the lane that becomes owner kBsXDist rows from now
      {
        const int it = i - kBsXDist;                   // its turn
        if (it >= 0 && lane == (it & 31)) {
          const int r = it - 32;
#pragma unroll
          for (int q = 0; q < 6; ++q) xn[q] = (r >= 0) ? x[6 * (long long)r + q] : 0.0;
        }
      }
      // the group of row i is consumed once its last row's block is in registers
      if ((itn % R) == R - 1 || i == 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + ((itn / R) % NS));
      }
      // labels: swap in the next chunk, fetch the one after
      if ((itn & 31) == 31) { lab_f = lab_f2; lab_rs = lab_rs2; }
      if ((itn & 31) == 0) {
        const int r = i - 32 - lane;
        lab_f2 = (r >= 0) ? e.first[r] : 0;
        lab_rs2 = (r >= 0) ? e.row_start[r] : 0;
      }
      if (i > 0) load_block(i - 1, itn + 1);
    }
  }
}

}  // namespace lvba
