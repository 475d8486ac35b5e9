// micro-benchmark of the factorisation's P3 phase: 16 warps, each thread C(6x6) -= L_i T_j^T with operands in
// shared memory (36 LDS.128 + 216 DFMA per thread and step), one __syncthreads per step.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int S = 38, P = 31;
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) k(double* out, int steps, long long* cyc, int stagger) {
  __shared__ __align__(16) double sL[P * S], sT[P * S];
  const int tid = threadIdx.x;
  for (int i = tid; i < P * S; i += 512) { sL[i] = 1e-3 * (i % 7); sT[i] = 1e-3 * (i % 5); }
  double C[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) C[q] = tid + q;
  // slot pattern: warp shares few rows (like the 8x8 super-block map)
  const int w = tid >> 5, l = tid & 31;
  const int islot = (w * 2 + (l >> 3)) % P, jslot = (w + (l & 7) * 3) % P;
  __syncthreads();
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    if (stagger > 0 && (tid & 128)) { const unsigned c0 = (unsigned)clock(); while ((unsigned)clock() - c0 < (unsigned)stagger) {} }
    const double2* lp = reinterpret_cast<const double2*>(sL + ((islot + s) % P) * S);
    const double2* tp = reinterpret_cast<const double2*>(sT + ((jslot + s) % P) * S);
    if (VARIANT == 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const double2 t0 = tp[3 * q], t1 = tp[3 * q + 1], t2 = tp[3 * q + 2];
#pragma unroll
        for (int xx = 0; xx < 3; ++xx) {
          const double2 l = lp[3 * q + xx];
          double* c0 = C + (2 * xx) * 6; double* c1 = C + (2 * xx + 1) * 6;
          c0[0] -= l.x * t0.x; c0[1] -= l.x * t0.y; c0[2] -= l.x * t1.x; c0[3] -= l.x * t1.y; c0[4] -= l.x * t2.x; c0[5] -= l.x * t2.y;
          c1[0] -= l.y * t0.x; c1[1] -= l.y * t0.y; c1[2] -= l.y * t1.x; c1[3] -= l.y * t1.y; c1[4] -= l.y * t2.x; c1[5] -= l.y * t2.y;
        }
      }
    } else if (VARIANT == 1) {
      // software pipelined: operands of q+1 are requested before the FMAs of q
      double2 t0 = tp[0], t1 = tp[1], t2 = tp[2], l0 = lp[0], l1 = lp[1], l2 = lp[2];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double2 nt0, nt1, nt2, nl0, nl1, nl2;
        if (q < 5) { nt0 = tp[3 * q + 3]; nt1 = tp[3 * q + 4]; nt2 = tp[3 * q + 5]; nl0 = lp[3 * q + 3]; nl1 = lp[3 * q + 4]; nl2 = lp[3 * q + 5]; }
        const double2 ls[3] = {l0, l1, l2};
#pragma unroll
        for (int xx = 0; xx < 3; ++xx) {
          const double2 l = ls[xx];
          double* c0 = C + (2 * xx) * 6; double* c1 = C + (2 * xx + 1) * 6;
          c0[0] -= l.x * t0.x; c0[1] -= l.x * t0.y; c0[2] -= l.x * t1.x; c0[3] -= l.x * t1.y; c0[4] -= l.x * t2.x; c0[5] -= l.x * t2.y;
          c1[0] -= l.y * t0.x; c1[1] -= l.y * t0.y; c1[2] -= l.y * t1.x; c1[3] -= l.y * t1.y; c1[4] -= l.y * t2.x; c1[5] -= l.y * t2.y;
        }
        if (q < 5) { t0 = nt0; t1 = nt1; t2 = nt2; l0 = nl0; l1 = nl1; l2 = nl2; }
      }
    } else if (VARIANT == 2) {
      // LDS only (no FMAs): shared-memory floor
      double acc = 0;
#pragma unroll
      for (int q = 0; q < 18; ++q) { const double2 a = tp[q], b = lp[q]; acc += a.x + a.y + b.x + b.y; }
      C[s & 1] += acc;
    } else {
      // DFMA only: FP64 floor
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int q = 0; q < 36; ++q) C[q] -= C[(q + 7) % 36] * 1e-9;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  double sum = 0;
#pragma unroll
  for (int q = 0; q < 36; ++q) sum += C[q];
  out[tid] = sum;
  if (tid == 0) cyc[0] = t1 - t0;
}
template <int V> void run(const char* name, int stagger) {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  const int steps = 2000;
  k<V><<<1, 512>>>(out, steps, cyc, stagger); cudaDeviceSynchronize();
  k<V><<<1, 512>>>(out, steps, cyc, stagger); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<V>);
  printf("%-28s stagger=%4d : %7.0f cycles/step  (regs %d, local %zu B)  err=%s\n", name, stagger, (double)h / steps, fa.numRegs, fa.localSizeBytes, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<2>("LDS only", 0);
  run<3>("DFMA only", 0);
  run<0>("baseline", 0);
  run<0>("baseline", 150);
  run<0>("baseline", 300);
  run<0>("baseline", 600);
  run<1>("sw-pipelined", 0);
  run<1>("sw-pipelined", 300);
  return 0;
}
