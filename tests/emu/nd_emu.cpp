// nd_emu.cpp — TEST INFRASTRUCTURE (never part of liblvba_b200.so): the substructured block LDL^T of
// global-lvba_b200/csrc/nd_plan.h + nd_passes.h run on the CPU.  The plan builder, the job tables, the order of the
// steps (nd::run) and every data-layout pass are the very code the device runs; the four heavy steps the device does with
// kernels (factor_la.cuh, nd_kernels.cuh) are plain reference loops here that implement the documented semantics of
// FactorJob / SpikeJob / SyrkSeg / BacksolveJob.  tests/test_nd_solver_emu.py compares the result with dense numpy solves.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../global-lvba_b200/csrc/nd_passes.h"
#include "host_exec.h"

using namespace lvba;

namespace {

inline long long blk(const EnvView& e, int r, int c) { return e.row_start[r] + (c - e.first[r]); }

void inv6_sym_lower(const double* A, double* K) {        // inverse of the symmetric matrix given by the lower triangle of A
  double M[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { M[i][j] = (i >= j) ? A[i * 6 + j] : A[j * 6 + i]; M[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int p = 0; p < 6; ++p) {                            // Gauss-Jordan without pivoting (LDL^T without pivoting)
    const double ip = 1.0 / M[p][p];
    for (int j = 0; j < 12; ++j) M[p][j] *= ip;
    for (int i = 0; i < 6; ++i)
      if (i != p) { const double f = M[i][p]; for (int j = 0; j < 12; ++j) M[i][j] -= f * M[p][j]; }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) K[i * 6 + j] = M[i][6 + j];
}

struct NdHostExec {
  HostExec base;
  template <class F> void pass(long long n, const F& f) { base.for_each(n, f); }
  void zero(double* p, long long n) { std::memset(p, 0, (size_t)n * sizeof(double)); }
  void correct_apply(const nd::Tables& t, const int* ids, int n_ids, int stride) { base.for_each((long long)n_ids * stride, nd::CorrectApplyF{t, ids, stride}); }
  void copy(double* d, const double* s, long long n) { std::memcpy(d, s, (size_t)n * sizeof(double)); }

  void factor(const FactorJob* jobs, int nj, int /*max_col*/) {
    for (int q = 0; q < nj; ++q) {
      const FactorJob& J = jobs[q];
      const EnvView& e = J.e;
      const int n = e.n, ns = J.n_stop;
      // local copy of the view (rows x their envelope columns)
      std::vector<std::vector<double>> A((size_t)n);
      for (int i = 0; i < n; ++i) {
        A[i].resize((size_t)(i - e.first[i] + 1) * 36);
        std::memcpy(A[i].data(), J.L + blk(e, i, e.first[i]) * 36, A[i].size() * sizeof(double));
      }
      auto at = [&](int i, int j) -> double* { return A[i].data() + (size_t)(j - e.first[i]) * 36; };
      std::vector<double> z((size_t)n * 6);
      std::memcpy(z.data(), J.z, z.size() * sizeof(double));
      std::vector<double> T((size_t)n * 36), Lc((size_t)n * 36);
      for (int k = 0; k < ns; ++k) {
        double K[36];
        inv6_sym_lower(at(k, k), K);
        double chk = 0.0;
        for (int i = 0; i < 36; ++i) chk += K[i];
        if (!std::isfinite(chk)) J.status[0] = 1;
        std::memcpy(J.dinv + (long long)k * 36, K, sizeof K);
        int hi = k;
        for (int i = k + 1; i < n && e.first[i] <= k; ++i) {
          hi = i;
          std::memcpy(&T[(size_t)i * 36], at(i, k), 36 * sizeof(double));
          for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) {
              double s = 0.0;
              for (int c = 0; c < 6; ++c) s += T[(size_t)i * 36 + a * 6 + c] * K[c * 6 + b];
              Lc[(size_t)i * 36 + a * 6 + b] = s;
            }
          std::memcpy(at(i, k), &Lc[(size_t)i * 36], 36 * sizeof(double));
          for (int a = 0; a < 6; ++a) {
            double s = 0.0;
            for (int c = 0; c < 6; ++c) s += Lc[(size_t)i * 36 + a * 6 + c] * z[6 * k + c];
            z[6 * i + a] -= s;
          }
        }
        for (int i = k + 1; i <= hi; ++i)
          for (int j = k + 1; j <= i; ++j) {
            double* d = at(i, j);
            for (int a = 0; a < 6; ++a)
              for (int b = 0; b < 6; ++b) {
                double s = 0.0;
                for (int c = 0; c < 6; ++c) s += Lc[(size_t)i * 36 + a * 6 + c] * T[(size_t)j * 36 + b * 6 + c];
                d[a * 6 + b] -= s;
              }
          }
      }
      // what the kernel leaves behind: L_ik below the pivots, z of the pivots, the trailing window and its rhs in the dumps;
      // the diagonal blocks and the trailing part of L / z stay as they were
      for (int i = 0; i < n; ++i)
        for (int j = e.first[i]; j < i && j < ns; ++j) std::memcpy(J.L + blk(e, i, j) * 36, at(i, j), 36 * sizeof(double));
      for (int k = 0; k < ns; ++k) for (int a = 0; a < 6; ++a) J.z[6 * k + a] = z[6 * k + a];
      const int bs = n - ns;
      if (bs > 0 && J.wdump) {
        for (int hi = 0; hi < bs; ++hi)
          for (int lo = 0; lo <= hi; ++lo) {
            double* dst = J.wdump + ((long long)hi * bs + lo) * 36;
            const int i = ns + hi, j = ns + lo;
            if (j >= e.first[i]) std::memcpy(dst, at(i, j), 36 * sizeof(double));
            else std::memset(dst, 0, 36 * sizeof(double));
          }
        if (J.zdump) for (int o = 0; o < bs * 6; ++o) J.zdump[o] = z[6 * ns + o];
      }
    }
  }

  void factor_dense(const FactorJob* jobs, int nj, int mc) { factor(jobs, nj, mc); }

  void spike(const nd::SpikeJob* jobs, int nj, int, int) {
    for (int q = 0; q < nj; ++q) {
      const nd::SpikeJob& J = jobs[q];
      const EnvView& e = J.e;
      const int KS = J.KS;
      for (int k = 0; k < e.n; ++k) {
        const int jend = k < J.n_stop ? k : J.n_stop;
        for (int x = 0; x < 6; ++x)
          for (int c = 0; c < KS; ++c) {
            double acc = (k < J.nE && k < J.n_stop) ? J.E[((long long)k * 6 + x) * KS + c] : 0.0;
            for (int j = e.first[k]; j < jend; ++j) {
              const double* b = J.L + blk(e, k, j) * 36 + x * 6;
              for (int y = 0; y < 6; ++y) acc -= b[y] * J.Z[((long long)j * 6 + y) * KS + c];
            }
            J.Z[((long long)k * 6 + x) * KS + c] = acc;
          }
      }
    }
  }

  void syrk(const nd::SyrkSeg* segs, int ns, int, int) {
    for (int q = 0; q < ns; ++q) {
      const nd::SyrkSeg& G = segs[q];
      const int KS = G.KS;
      std::vector<double> KZ((size_t)6 * KS);
      for (int k = 0; k < G.rows; ++k) {
        const double* Kk = G.K + (long long)k * 36;
        const double* Zk = G.Z + (long long)k * 6 * KS;
        double Kw[6];
        for (int p = 0; p < 6; ++p) {
          double s = 0.0;
          for (int r = 0; r < 6; ++r) s += Kk[p * 6 + r] * G.w[(long long)k * 6 + r];
          Kw[p] = s;
          for (int c = 0; c < KS; ++c) {
            double t = 0.0;
            for (int r = 0; r < 6; ++r) t += Kk[p * 6 + r] * Zk[(long long)r * KS + c];
            KZ[(size_t)p * KS + c] = t;
          }
        }
        for (int ga = 0; ga < KS; ++ga) {
          const int bi = ga / 6;
          for (int gb = 0; gb < KS; ++gb) {
            const int bj = gb / 6;
            if (bj > bi) continue;
            double s = 0.0;
            for (int p = 0; p < 6; ++p) s += KZ[(size_t)p * KS + ga] * Zk[(long long)p * KS + gb];
            G.U[nd::tri_off(bi, bj) + (ga - 6 * bi) * 6 + (gb - 6 * bj)] -= s;
          }
          double s = 0.0;
          for (int p = 0; p < 6; ++p) s += Zk[(long long)p * KS + ga] * Kw[p];
          G.u[ga] -= s;
        }
      }
    }
  }

  void backsolve(const BacksolveJob* jobs, int nj) {
    for (int q = 0; q < nj; ++q) {
      const BacksolveJob& J = jobs[q];
      const EnvView& e = J.e;
      for (int i = e.n - 1; i >= 0; --i) {
        const int upto = i < J.n_given ? i : J.n_given;
        for (int j = e.first[i]; j < upto; ++j) {
          const double* b = J.L + blk(e, i, j) * 36;
          for (int c = 0; c < 6; ++c) {
            double s = 0.0;
            for (int a = 0; a < 6; ++a) s += b[a * 6 + c] * J.x[6 * (long long)i + a];
            J.x[6 * (long long)j + c] -= s;
          }
        }
      }
    }
  }
};

template <class T>
T* poisoned(std::vector<T>& v, size_t n) {
  v.resize(n + 1);
  std::memset((void*)v.data(), 0xA5, (n + 1) * sizeof(T));
  return v.data();
}

// one rank's buffers and tables
struct RankState {
  std::vector<long long> row_start;
  std::vector<int> last;
  std::vector<nd::NodeDev> nodes;
  std::vector<double> vH, vdadd, vL, vz, vzs, vd, vU, vZ, vE, vT, vW, vw, vx;
  std::vector<int> zeros, last_by_w, status, sep_row0, sep_rows;
  std::vector<long long> tri;
  std::vector<nd::LevelJobs> jobs;
  std::vector<nd::LevelDev> lv;
  nd::Tables t{};
  nd::RegionDev reg{};
};

static void setup_rank(RankState& S, const nd::Plan& P, int n, const int* first, const double* H, const double* dadd, const double* rhs,
                       long long nblocks, int rank) {
  S.nodes.clear();
  for (const nd::Node& v : P.nodes) S.nodes.push_back(nd::to_dev(v));
  S.vH.assign(H, H + (size_t)nblocks * 36);
  S.vdadd.assign(dadd, dadd + (size_t)n * 6);
  if (P.n_ranks > 1) {                                     // a rank only has the rows it owns: poison everything else
    const double nan = std::nan("");
    for (int r = 0; r < n; ++r)
      if (r < P.rank_row_begin[rank] || r >= P.rank_row_end[rank]) {
        for (long long b = S.row_start[r]; b < S.row_start[r + 1]; ++b) for (int q = 0; q < 36; ++q) S.vH[(size_t)b * 36 + q] = nan;
        for (int q = 0; q < 6; ++q) S.vdadd[(size_t)r * 6 + q] = nan;
      }
  }
  nd::Tables& t = S.t;
  t = nd::Tables{};
  t.n = n; t.first = first; t.row_start = S.row_start.data(); t.nodes = S.nodes.data(); t.H = S.vH.data(); t.dadd = S.vdadd.data();
  t.Hw = S.vH.data(); t.daddw = S.vdadd.data();
  t.L = poisoned(S.vL, (size_t)nblocks * 36);
  t.z = poisoned(S.vz, (size_t)n * 6);
  std::memcpy(t.z, rhs, (size_t)n * 6 * sizeof(double));
  t.zs = poisoned(S.vzs, (size_t)n * 6);
  t.dinv = poisoned(S.vd, (size_t)n * 36);
  t.x = poisoned(S.vx, (size_t)n * 6);
  t.U = poisoned(S.vU, (size_t)P.sizeU); t.u = t.U; t.Z = poisoned(S.vZ, (size_t)P.sizeZ);
  t.E = poisoned(S.vE, (size_t)P.sizeE); t.T = poisoned(S.vT, (size_t)P.sizeT); t.W = poisoned(S.vW, (size_t)P.sizeW);
  t.w = poisoned(S.vw, (size_t)P.sizew);
  S.zeros.assign(32, 0); S.last_by_w.assign(32 * 32, 0); S.tri.assign(33, 0);
  for (int i = 0; i < 33; ++i) S.tri[i] = (long long)i * (i + 1) / 2;
  for (int w = 0; w < 32; ++w) for (int i = 0; i < 32; ++i) S.last_by_w[w * 32 + i] = w - 1;
  nd::DenseViewArrays dv{S.zeros.data(), S.tri.data(), S.last_by_w.data()};
  S.status.assign(P.nodes.size(), 0);
  nd::build_level_jobs(P, t, P.first_rel.data(), P.rs_adj.data(), P.last_rel.data(), nblocks, dv, S.status.data(), S.jobs, rank);
  S.lv.clear();
  for (auto& J : S.jobs)
    S.lv.push_back(nd::LevelDev{J.ids.data(), (int)J.ids.size(), J.factor.data(), (int)J.factor.size(), J.spike.data(), (int)J.spike.size(),
                                J.syrk.data(), (int)J.syrk.size(), J.back.data(), (int)J.back.size(), J.asm_stride, J.corr_stride,
                                J.max_ks, J.max_rows, J.max_col});
  S.sep_row0.assign((size_t)P.n_ranks, -1); S.sep_rows.assign((size_t)P.n_ranks, 0);
  for (int r = 0; r + 1 < P.n_ranks; ++r) { S.sep_row0[r] = P.sep_start[(r + 1) * P.q]; S.sep_rows[r] = P.sep_width[(r + 1) * P.q]; }
  S.reg = nd::RegionDev{P.n_ranks, rank, P.slot_rows, P.max_col, P.region0, P.slot, P.slotU, P.slotu, P.slotH, S.sep_row0.data(), S.sep_rows.data()};
}

}  // namespace

// Solves (H + diag(dadd)) x = rhs with p_want chunks (fewer when the structure does not allow as many) as n_ranks ranks would:
// every rank sees only the rows it owns (the others are NaN), runs its own subtree, the exchange region is "all-gathered" by
// copying the slots, every rank runs the top tree and its downward sweep; x is assembled from the rows each rank owns.
// first[] must be monotone (Envelope::build); H is the envelope storage for it.  Returns the number of chunks used, 0 when
// no plan exists.  info: [0] tree depth (levels), [1] nodes, [2] longest interior, [3] widest separator
extern "C" int nd_emu_solve_ranks(int n, const int* first, const double* H, const double* dadd, const double* rhs, double* x,
                                  int p_want, int n_ranks, int* info) {
  std::vector<long long> row_start((size_t)n + 1, 0);
  for (int r = 0; r < n; ++r) row_start[r + 1] = row_start[r] + (r - first[r] + 1);
  const long long nblocks = row_start[n];
  std::vector<int> last((size_t)n, 0);
  int max_col = 0;
  {
    int i = 0;
    for (int k = 0; k < n; ++k) {
      if (i < k) i = k;
      while (i + 1 < n && first[i + 1] <= k) ++i;
      last[k] = i;
      max_col = std::max(max_col, i - k);
    }
  }
  nd::Plan P;
  const int p = nd::choose_chunks(n, first, last.data(), row_start.data(), max_col, p_want, P, n_ranks);
  if (p == 0) return 0;
  std::vector<RankState> R((size_t)n_ranks);
  NdHostExec ex;
  const long long leaf_e = nd::leaf_e_stride(P), leaf_fin = nd::leaf_final_stride(P);
  for (int r = 0; r < n_ranks; ++r) {
    R[r].row_start = row_start;
    setup_rank(R[r], P, n, first, H, dadd, rhs, nblocks, r);
    nd::run_up_local(ex, P, R[r].t, R[r].lv.data(), (int)R[r].lv.size(), nblocks, leaf_e, leaf_fin, &R[r].reg);
  }
  if (n_ranks > 1)
    for (int dst = 0; dst < n_ranks; ++dst)                 // the all-gather: slot r of rank r -> slot r of everybody
      for (int src = 0; src < n_ranks; ++src)
        if (src != dst) std::memcpy(R[dst].t.U + P.region0 + (long long)src * P.slot, R[src].t.U + P.region0 + (long long)src * P.slot, (size_t)P.slot * sizeof(double));
  int bad = 0;
  for (int r = 0; r < n_ranks; ++r) {
    nd::run_top_down(ex, P, R[r].t, R[r].lv.data(), (int)R[r].lv.size(), n_ranks > 1 ? &R[r].reg : nullptr);
    ex.pass((long long)6 * n, nd::ZeroForeignF{R[r].t.x, P.rank_row_begin[r], P.rank_row_end[r]});
    for (int st : R[r].status) bad |= st;
  }
  for (int i = 0; i < 6 * n; ++i) { double s = 0.0; for (int r = 0; r < n_ranks; ++r) s += R[r].t.x[i]; x[i] = s; }   // the all-reduce
  if (info) {
    info[0] = (int)P.levels.size(); info[1] = (int)P.nodes.size();
    int mi = 0, mw = 0;
    for (int c = 0; c < P.p; ++c) mi = std::max(mi, P.nodes[c].npiv);
    for (int j = 1; j < P.p; ++j) mw = std::max(mw, P.sep_width[j]);
    info[2] = mi; info[3] = mw;
  }
  return bad ? -p : p;
}

extern "C" int nd_emu_solve(int n, const int* first, const double* H, const double* dadd, const double* rhs, double* x,
                            int p_want, int* info) {
  return nd_emu_solve_ranks(n, first, H, dadd, rhs, x, p_want, 1, info);
}

// ------------------------------------------------------------------------------------------------
// One rank at a time, for the world-size-2 gloo test (tests/test_rank_flow_gloo.py): the collectives between the two halves
// are REAL torch.distributed calls between processes, not the in-process copies of nd_emu_solve_ranks.
//   nd_emu_plan_rows  -> the block rows each rank owns (what the library reports through lvba_lidar_owned_rows)
//   nd_emu_rank_up    -> this rank eliminates its own subtree from ITS rows of H (foreign rows are poisoned) and packs its slot;
//                        returns the exchange region (n_ranks slots; the caller all-gathers slot `rank` into everybody's region)
//   nd_emu_rank_down  -> top of the tree, downward sweep, foreign rows of x zeroed (the caller all-reduces x)
namespace {
struct RankRun {
  nd::Plan P;
  RankState S;
  std::vector<int> first, last;
  int n = 0, rank = 0, n_ranks = 1;
};
RankRun* g_run = nullptr;

int plan_for(int n, const int* first, int p_want, int n_ranks, std::vector<long long>& row_start, std::vector<int>& last, nd::Plan& P) {
  row_start.assign((size_t)n + 1, 0);
  for (int r = 0; r < n; ++r) row_start[r + 1] = row_start[r] + (r - first[r] + 1);
  last.assign((size_t)n, 0);
  int max_col = 0, i = 0;
  for (int k = 0; k < n; ++k) {
    if (i < k) i = k;
    while (i + 1 < n && first[i + 1] <= k) ++i;
    last[k] = i;
    max_col = std::max(max_col, i - k);
  }
  return nd::choose_chunks(n, first, last.data(), row_start.data(), max_col, p_want, P, n_ranks);
}
}  // namespace

extern "C" int nd_emu_plan_rows(int n, const int* first, int p_want, int n_ranks, int* row_begin, int* row_end) {
  nd::Plan P;
  std::vector<long long> rs;
  std::vector<int> last;
  const int p = plan_for(n, first, p_want, n_ranks, rs, last, P);
  if (p == 0) return 0;
  for (int r = 0; r < n_ranks; ++r) { row_begin[r] = P.rank_row_begin[r]; row_end[r] = P.rank_row_end[r]; }
  return p;
}

extern "C" int nd_emu_rank_up(int n, const int* first, const double* H, const double* dadd, const double* rhs, int p_want, int n_ranks,
                              int rank, double** region, long long* slot_doubles) {
  delete g_run;
  g_run = new RankRun();
  RankRun& G = *g_run;
  G.n = n; G.rank = rank; G.n_ranks = n_ranks;
  G.first.assign(first, first + n);
  const int p = plan_for(n, G.first.data(), p_want, n_ranks, G.S.row_start, G.last, G.P);
  if (p == 0) return 0;
  const long long nblocks = G.S.row_start[n];
  setup_rank(G.S, G.P, n, G.first.data(), H, dadd, rhs, nblocks, rank);
  NdHostExec ex;
  nd::run_up_local(ex, G.P, G.S.t, G.S.lv.data(), (int)G.S.lv.size(), nblocks, nd::leaf_e_stride(G.P), nd::leaf_final_stride(G.P), &G.S.reg);
  *region = G.S.t.U + G.P.region0;
  *slot_doubles = G.P.slot;
  return p;
}

extern "C" int nd_emu_rank_down(double* x) {
  if (!g_run) return -1;
  RankRun& G = *g_run;
  NdHostExec ex;
  nd::run_top_down(ex, G.P, G.S.t, G.S.lv.data(), (int)G.S.lv.size(), G.n_ranks > 1 ? &G.S.reg : nullptr);
  ex.pass((long long)6 * G.n, nd::ZeroForeignF{G.S.t.x, G.P.rank_row_begin[G.rank], G.P.rank_row_end[G.rank]});
  int bad = 0;
  for (int st : G.S.status) bad |= st;
  std::memcpy(x, G.S.t.x, (size_t)6 * G.n * sizeof(double));
  delete g_run;
  g_run = nullptr;
  return bad;
}
