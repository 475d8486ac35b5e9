// chunk_lab — development harness of the EXPERIMENTAL four-chunk solve (tools/lab/four_chunk.cuh): random
// block-banded systems, solution compared with the production twisted solve and with the dense residual, timings.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o chunk_lab chunk_lab.cu -ldl
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include "four_chunk.cuh"

using namespace lvba;

static int run_case(int n, int b, unsigned seed) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  std::vector<int> first(n);
  for (int r = 0; r < n; ++r) first[r] = std::max(0, r - b);
  Envelope env;
  int64_t bytes = 0;
  if (env.build(first, s, &bytes) != LVBA_OK) { printf("env build failed\n"); return 1; }
  std::vector<double> H((size_t)env.nblocks * 36), rhs((size_t)n * 6), dadd((size_t)n * 6);
  for (int r = 0; r < n; ++r)
    for (int c = env.first[r]; c <= r; ++c) {
      double* blk = &H[(size_t)(env.row_start[r] + (c - env.first[r])) * 36];
      for (int q = 0; q < 36; ++q) blk[q] = 0.3 * U(rng);
      if (c == r) {
        for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) blk[j * 6 + i] = blk[i * 6 + j];
        for (int i = 0; i < 6; ++i) blk[i * 7] = 14.0 + 2.0 * U(rng);
      }
    }
  for (auto& v : rhs) v = U(rng);
  for (auto& v : dadd) v = 0.1 + 0.05 * U(rng);
  DevBuf<double> dH, dD, dR, dX;
  dH.upload(H, s); dD.upload(dadd, s); dR.upload(rhs, s); dX.alloc((size_t)n * 6);
  printf("=== n=%d b=%d\n", n, b);
  auto residual = [&](const std::vector<double>& x) {
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
    return rn;
  };
  std::vector<double> xref((size_t)n * 6), x4((size_t)n * 6);
  int64_t launches = 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  {
    EnvSolver sol;
    if (sol.prepare(env, s) != LVBA_OK) { printf("prepare failed: %s\n", last_error_ref().c_str()); return 1; }
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemcpyAsync(sol.z.p, dR.p, rhs.size() * 8, cudaMemcpyDeviceToDevice, s);
      cudaEventRecord(e0, s);
      sol.solve(env, dH.p, dD.p, dX.p, s, &launches);
      cudaEventRecord(e1, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) { printf("twisted: CUDA error\n"); return 2; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    cudaMemcpy(xref.data(), dX.p, xref.size() * 8, cudaMemcpyDeviceToHost);
    printf("  twisted (2 chunks) : %.3f ms  |resid| = %.2e\n", best, residual(xref));
  }
  int rc = 0;
  {
    ChunkedSolver cs;
    if (cs.prepare(env, s) != LVBA_OK) { printf("  four chunks: prepare refused: %s\n", last_error_ref().c_str()); return 0; }
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemsetAsync(dX.p, 0, (size_t)n * 48, s);
      cudaEventRecord(e0, s);
      if (cs.solve(env, dH.p, dD.p, dR.p, dX.p, s, &launches) != LVBA_OK) { printf("  four chunks: solve failed: %s\n", last_error_ref().c_str()); return 1; }
      cudaEventRecord(e1, s);
      cudaError_t err = cudaStreamSynchronize(s);
      if (err != cudaSuccess) { printf("  four chunks: CUDA error %s\n", cudaGetErrorString(err)); return 2; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    cudaMemcpy(x4.data(), dX.p, x4.size() * 8, cudaMemcpyDeviceToHost);
    const double rn = residual(x4);
    auto seg = [&](int r0, int r1) { double d = 0; for (int i = 6 * r0; i < 6 * r1; ++i) d = std::max(d, std::fabs(x4[i] - xref[i])); return d; };
    printf("  four chunks        : %.3f ms  |resid| = %.2e   |x - x_twisted|: T %.2e  S %.2e  B %.2e   (m=%d s=%d)\n", best, rn,
           seg(0, cs.m), seg(cs.m, cs.m + cs.s), seg(cs.m + cs.s, n), cs.m, cs.s);
    // where inside T / B the error sits: the four chunk boundaries of the inner twisted splits
    printf("     T: top %.2e  inner sep %.2e  bottom %.2e ;  B: top %.2e  inner sep %.2e  bottom %.2e\n",
           seg(0, cs.solT.tw_m), seg(cs.solT.tw_m, cs.solT.tw_send), seg(cs.solT.tw_send, cs.m),
           seg(cs.m + cs.s, cs.m + cs.s + cs.solB.tw_m), seg(cs.m + cs.s + cs.solB.tw_m, cs.m + cs.s + cs.solB.tw_send), seg(cs.m + cs.s + cs.solB.tw_send, n));
    if (!(rn < 1e-9)) rc = 3;
  }
  cudaStreamDestroy(s);
  return rc;
}

int main() {
  int rc = 0;
  rc |= run_case(2000, 30, 1);
  rc |= run_case(1999, 20, 2);
  rc |= run_case(5000, 30, 3);
  printf(rc ? "CHUNK LAB: FAILURES\n" : "CHUNK LAB: all cases consistent\n");
  return rc;
}
