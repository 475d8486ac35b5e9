// dmma.cu — latency / issue rate of the FP64 tensor-core instruction mma.sync.m8n8k4.f64 on B200, next to plain DFMA:
// decides whether the GEMM-shaped solver kernels (spike, SYRK: latency / instruction-count bound, profiles/r02_*) should use it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, long long* clk) {
  double c[CHAINS][2];
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x; c[i][1] = 1.0; }
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int CHAINS>
__global__ void k_dfma(double* out, int iters, long long* clk) {
  double c[CHAINS];
  for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x + i;
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <class K>
void run(const char* name, K kern, int chains, int threads, int blocks, int flop_per_instr_warp) {
  double* out; long long* clk;
  cudaMalloc(&out, sizeof(double) * threads * blocks); cudaMalloc(&clk, 8);
  const int iters = 4096;
  kern<<<blocks, threads>>>(out, iters, clk);
  cudaDeviceSynchronize();
  kern<<<blocks, threads>>>(out, iters, clk);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)iters * chains);
  printf("%-28s chains %d warps/CTA %2d CTAs %4d : %7.2f clk per instruction (per warp)   %s\n", name, chains, threads / 32, blocks, per,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run("DMMA m8n8k4 dependent", k_dmma<1>, 1, 32, 1, 512);
  run("DMMA m8n8k4", k_dmma<2>, 2, 32, 1, 512);
  run("DMMA m8n8k4", k_dmma<4>, 4, 32, 1, 512);
  run("DMMA m8n8k4", k_dmma<8>, 8, 32, 1, 512);
  run("DMMA m8n8k4", k_dmma<4>, 4, 128, 1, 512);
  run("DMMA m8n8k4", k_dmma<4>, 4, 512, 1, 512);
  run("DMMA m8n8k4 (all SMs)", k_dmma<4>, 4, 512, 148, 512);
  run("DFMA dependent", k_dfma<1>, 1, 32, 1, 64);
  run("DFMA", k_dfma<2>, 2, 32, 1, 64);
  run("DFMA", k_dfma<4>, 4, 32, 1, 64);
  run("DFMA", k_dfma<8>, 8, 32, 1, 64);
  run("DFMA", k_dfma<8>, 8, 128, 1, 64);
  run("DFMA", k_dfma<8>, 8, 512, 1, 64);
  return 0;
}
