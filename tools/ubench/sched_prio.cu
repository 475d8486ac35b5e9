// micro-benchmark: how fast does a DEPENDENT FP64 chain advance on a warp that shares its SMSP with four warps
// issuing long runs of independent DFMAs, as a function of the chain warp's id (first vs last warpgroup)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(640, 1) k(int chain_first, int heavy_iters, int chain_len, double* out, long long* cyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool chain = chain_first ? warp < 4 : warp >= 16;
  __syncthreads();
  if (chain) {
    double a = 1.0 + lane * 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < chain_len; ++i) a = fma(a, 1.0000001, 1e-9);
    const long long t1 = clock64();
    double r = a;
    const long long t2 = clock64();
    for (int i = 0; i < chain_len / 8; ++i) r = __drcp_rn(r) + 0.5;
    const long long t3 = clock64();
    out[threadIdx.x] = a + r;
    if (lane == 0) { cyc[warp * 2] = t1 - t0; cyc[warp * 2 + 1] = t3 - t2; }
  } else {
    double acc[36];
    for (int i = 0; i < 36; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    const long long t0 = clock64();
    for (int it = 0; it < heavy_iters; ++it)
#pragma unroll
      for (int i = 0; i < 36; ++i) acc[i] = fma(acc[i], 1.0000001, 1e-9);
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < 36; ++i) s += acc[i];
    out[threadIdx.x] = s;
    if (lane == 0) { cyc[warp * 2] = t1 - t0; cyc[warp * 2 + 1] = 0; }
  }
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 4096);
  for (int heavy = 0; heavy <= 1; ++heavy)
    for (int first = 0; first <= 1; ++first) {
      const int heavy_iters = heavy ? 4000 : 0, chain_len = 4000;
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 640>>>(first, heavy_iters, chain_len, out, cyc); cudaDeviceSynchronize(); }
      long long h[40]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      const int cw = first ? 0 : 16, hw = first ? 4 : 0;
      printf("heavy=%d chain warps %s: dependent DFMA %.1f clk/op, dependent (drcp+add) %.1f clk/op ; heavy warp: %.2f clk per warp-DFMA (x4 warps/SMSP)\n",
             heavy, first ? "FIRST (ids 0-3)" : "LAST (ids 16-19)", (double)h[cw * 2] / chain_len, (double)h[cw * 2 + 1] / (chain_len / 8),
             heavy ? (double)h[hw * 2] / (heavy_iters * 36.0) : 0.0);
    }
  return 0;
}
