// micro-benchmark: FP64 FMA issue rate and dependent latency on one SM (B200)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP> void run(int threads, int iters) {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
  k<ILP><<<1, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<ILP><<<1, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double warp_fma = (double)iters * ILP * (threads / 32);
  printf("ILP=%2d threads=%4d: %8lld cycles, %.2f cycles per warp-DFMA per SMSP, %.1f FMA/clk/SM\n", ILP, threads, h,
         (double)h / (warp_fma / 4.0 > 0 ? warp_fma / (threads >= 128 ? 4.0 : threads / 32.0) : 1), warp_fma * 32 / h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(32, 4096);    // dependent latency
  run<2>(32, 4096);
  run<4>(32, 4096);
  run<8>(32, 4096);
  run<16>(32, 2048);
  run<8>(128, 2048);   // 1 warp per SMSP
  run<8>(256, 2048);
  run<8>(512, 2048);
  run<8>(640, 2048);
  run<16>(512, 1024);
  return 0;
}
