// solver_lab — development harness for the block-envelope LDL^T kernels (not part of the product):
// builds random block-banded systems, runs every (factor kernel, backsolve kernel, twisted/single) combination of
// EnvSolver::solve, prints max differences against the first combination and CUDA-event timings.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o solver_lab solver_lab.cu -ldl
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#define LVBA_LAB 1
#include "../../global-lvba_b200/csrc/runtime.cuh"

using namespace lvba;

static double maxdiff(const std::vector<double>& a, const std::vector<double>& b, double* scale) {
  double d = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) { d = std::max(d, std::fabs(a[i] - b[i])); s = std::max(s, std::fabs(a[i])); }
  *scale = s;
  return d;
}

static bool g_prof = false;
static int run_case(int n, int b, int ragged, unsigned seed, bool timing = false) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  std::vector<int> first(n);
  for (int r = 0; r < n; ++r) first[r] = std::max(0, r - b + (ragged ? (int)(rng() % (ragged + 1)) : 0));
  Envelope env;
  int64_t bytes = 0;
  if (env.build(first, s, &bytes) != LVBA_OK) { printf("env build failed: %s\n", last_error_ref().c_str()); return 1; }
  std::vector<double> H((size_t)env.nblocks * 36), rhs((size_t)n * 6), dadd((size_t)n * 6, 0.0);
  for (int r = 0; r < n; ++r)
    for (int c = env.first[r]; c <= r; ++c) {
      double* blk = &H[(size_t)(env.row_start[r] + (c - env.first[r])) * 36];
      for (int q = 0; q < 36; ++q) blk[q] = 0.3 * U(rng);
      if (c == r) {
        for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) blk[j * 6 + i] = blk[i * 6 + j];
        for (int i = 0; i < 6; ++i) blk[i * 7] = 14.0 + 2.0 * U(rng);      // comfortably definite pivots
      }
    }
  for (auto& v : rhs) v = U(rng);
  for (auto& v : dadd) v = 0.1 + 0.05 * U(rng);
  DevBuf<double> dH, dD, dX;
  dH.upload(H, s); dD.upload(dadd, s); dX.alloc((size_t)n * 6);
  printf("=== case n=%d b=%d ragged=%d  nblocks=%lld max_col=%d\n", n, b, ragged, env.nblocks, env.max_col);
  std::vector<double> x0, L0, z0;
  int rc_all = 0;
  for (int tw = 1; tw >= 0; --tw) {
    setenv("LVBA_NO_TWIST", tw ? "0" : "1", 1);
    // cfg: factor kernel x backsolve kernel (+ lab modes that disable one side of the factor kernel: timing only)
    struct Cfg { int mode; int tile2; };
    const Cfg cfgs[] = {{0, 1}, {1, 1}, {2, 1}};
    for (int cfg = 0; cfg < (int)(sizeof cfgs / sizeof cfgs[0]); ++cfg) {
      if (cfgs[cfg].mode != 0 && !(timing && tw == 1)) continue;
      if (g_prof && !(cfg == 0 && tw == 1)) continue;     // profiling run: twisted, once
      EnvSolver sol;
      setenv("LVBA_FACTOR_TIMING", (timing && tw == 1 && !g_prof) ? "1" : "0", 1);
      sol.dbg_max_dumps = 1;
      cudaMemcpyToSymbol(g_la_mode, &cfgs[cfg].mode, sizeof(int));
      if (sol.prepare(env, s) != LVBA_OK) { printf("prepare failed: %s\n", last_error_ref().c_str()); return 1; }
      int64_t launches = 0;
      float best = 1e30f;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < (g_prof ? 1 : 4); ++rep) {
        cudaMemcpyAsync(sol.z.p, rhs.data(), rhs.size() * 8, cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(dX.p, 0, (size_t)n * 48, s);
        cudaEventRecord(e0, s);
        if (sol.solve(env, dH.p, dD.p, dX.p, s, &launches) != LVBA_OK) { printf("solve failed: %s\n", last_error_ref().c_str()); return 1; }
        cudaEventRecord(e1, s);
        cudaError_t err = cudaStreamSynchronize(s);
        if (err != cudaSuccess) { printf("  tw=%d CUDA ERROR %s\n", tw, cudaGetErrorString(err)); return 2; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
      }
      std::vector<double> x((size_t)n * 6), Lh((size_t)env.nblocks * 36), zh((size_t)n * 6);
      int st = 0;
      cudaMemcpy(x.data(), dX.p, x.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(Lh.data(), sol.L.p, Lh.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(zh.data(), sol.z.p, zh.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(&st, sol.status.p, 4, cudaMemcpyDeviceToHost);
      // residual of the damped system, lower envelope expanded
      std::vector<double> r(rhs);
      for (int i = 0; i < n; ++i)
        for (int c = env.first[i]; c <= i; ++c) {
          const double* blk = &H[(size_t)(env.row_start[i] + (c - env.first[i])) * 36];
          for (int aa = 0; aa < 6; ++aa)
            for (int bb = 0; bb < 6; ++bb) {
              double v = (c == i && bb > aa) ? blk[bb * 6 + aa] : blk[aa * 6 + bb];
              if (c == i && aa == bb) v += dadd[6 * i + aa];
              r[6 * i + aa] -= v * x[6 * c + bb];
              if (c != i) r[6 * c + bb] -= v * x[6 * i + aa];
            }
        }
      double rn = 0; for (double v : r) rn = std::max(rn, std::fabs(v));
      double sx = 0, dx = 0, sl = 0, dl = 0, sz = 0, dz = 0;
      if (x0.empty() || cfg == 0) { if (tw == 1 && cfg == 0) { x0 = x; } L0 = Lh; z0 = zh; }
      dx = maxdiff(x0, x, &sx); dl = maxdiff(L0, Lh, &sl); dz = maxdiff(z0, zh, &sz);
      printf("  twisted=%d tile2=%d mode=%d : %.3f ms  status=%d  |resid|=%.2e  dx=%.2e (of %.1e)  dL=%.2e (of %.1e)  dz=%.2e\n",
             tw, cfgs[cfg].tile2, cfgs[cfg].mode, best, st, rn, dx, sx, dl, sl, dz);
      fflush(stdout);
      if (cfgs[cfg].mode == 0 && (!(rn < 1e-9) || st != 0)) rc_all = 3;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
  }
  cudaStreamDestroy(s);
  return rc_all;
}

int main(int argc, char** argv) {
  int rc = 0;
  if (argc > 1 && std::string(argv[1]) == "prof") { g_prof = true; return run_case(2000, 30, 0, 1, true); }
  rc |= run_case(2000, 30, 0, 1, true);
  rc |= run_case(1999, 20, 0, 2, true);
  rc |= run_case(700, 30, 3, 3);
  rc |= run_case(300, 12, 2, 4);
  rc |= run_case(40, 5, 1, 5);
  rc |= run_case(24, 23, 0, 6);
  rc |= run_case(7, 3, 0, 7);
  printf(rc ? "LAB: FAILURES\n" : "LAB: all cases consistent\n");
  return rc;
}
