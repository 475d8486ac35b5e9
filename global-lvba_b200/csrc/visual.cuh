// visual.cuh — hot path B kernels: per-track reprojection / point-to-plane residuals and analytic
// Jacobians, Schur elimination of the 3x3 landmark blocks straight into the block-envelope reduced
// camera system, back-substitution and the model-cost evaluation.
//
// Restates (analytic Jacobians instead of ceres::Jet auto-diff; block-sparse instead of DENSE_SCHUR):
//   ReprojErrorWhitenedDistorted::operator()   reference include/utils.hpp:61-111
//   PointPlaneErrorWhitened::operator()        reference include/utils.hpp:133-139
//   problem structure of optimizeCameraPoses   reference src/lvba_system.cpp:1578-1640
//   ceres-solver 2.1.0 (not vendored): Jacobi column scaling, LM diagonal, Schur complement,
//   EigenQuaternionManifold on {w,x,y,z} memory.
//
// Schedule: the observations of the landmarks that have a valid plane are stored CSR and cut into
// batches of consecutive landmarks with <= kSlots observations; one CTA of kSlots threads per batch.
//   phase 1  thread = observation: r (2), J_cam (2x6, tangent), J_X (2x3); J_X^T J_X, J_X^T r staged
//   phase 2  thread = landmark   : C = sum J_X^T J_X + plane term + D_p^2, C^-1, g_p, w = C^-1 g_p
//   phase 3  thread = observation: E = J_c^T J_X, Y = E C^-1, diagonal block J_c^T J_c - Y E^T,
//                                  reduced rhs -(g_c - E w), column norms; factors [Y|E] to shared memory
//   phase 4  warp = 8 camera pairs: S(hi,lo)[a][b] -= sum_m Y_hi[a][m] E_lo[b][m]  (3 FMA + RED.ADD.F64)
// The Jacobian is never materialised in HBM; back-substitution recomputes it.
#pragma once
#include "common.cuh"
#include "envelope.cuh"
#include "lidar.cuh"   // kSlots, kStageStride, reduce kernels

namespace lvba {

constexpr int kMaxTrkPerBatch = 128;
constexpr int kTrkParams = 16;     // Cinv(6) g_p(3) w(3) y_p(3) pad
constexpr int kVFStride = 37;      // Y(18) E(18) + pad

struct VisualView {
  int n_batches;
  const int* trk_ptr;        // [Tv+1] obs offsets (valid landmarks only, compact)
  const int* trk_id;         // [Tv]   original landmark index
  const int* batch_trk;      // [n_batches+1]
  const long long* batch_pair;   // [n_batches+1]
  const unsigned* pairs;     // (hi | lo<<8 | local trk<<16)
  const int* obs_cam;        // [nnz] original camera index
  const int* obs_row;        // [nnz] reduced-system row of that camera, -1 when constant / unused
  const float2* obs_uv;      // [nnz]
  const double* plane;       // [Tv*4]
  double intr[8];
  double inv_sigma_px, inv_sigma_pl;
};

struct VisualState {
  const double* q;   // [M*4]
  const double* t;   // [M*3]
  const double* X;   // [T*3]
};

struct ObsEval {
  double r[2];
  double Jc[12];     // 2x6  [rho][0..2 rotation tangent | 3..5 translation]
  double JX[6];      // 2x3
};

// utils.hpp:61-111 forward model + analytic Jacobians (SURVEY.md A.2, Q9 tangent basis)
template <bool kJac>
LVBA_DEV void obs_eval(const VisualView& vv, const double* __restrict__ qp, const double* __restrict__ tp,
                       const double* X, float2 uv, ObsEval& o) {
  double q0 = qp[0], q1 = qp[1], q2 = qp[2], q3 = qp[3];
  const double inv = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);   // QuaternionRotatePoint normalises
  q0 *= inv; q1 *= inv; q2 *= inv; q3 *= inv;
  const double v[3] = {q1, q2, q3};
  double vxX[3], t2[3];
  cross3(v, X, vxX);
  const double uvv[3] = {2.0 * vxX[0], 2.0 * vxX[1], 2.0 * vxX[2]};
  cross3(v, uvv, t2);
  const double Xc[3] = {X[0] + q0 * uvv[0] + t2[0] + tp[0], X[1] + q0 * uvv[1] + t2[1] + tp[1],
                        X[2] + q0 * uvv[2] + t2[2] + tp[2]};
  const double z = Xc[2];
  if (!(z > 1e-8)) {                                   // utils.hpp:78
    o.r[0] = 0.0; o.r[1] = 0.0;
    if (kJac) {
#pragma unroll
      for (int i = 0; i < 12; ++i) o.Jc[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) o.JX[i] = 0.0;
    }
    return;
  }
  const double fx = vv.intr[0], fy = vv.intr[1], cx = vv.intr[2], cy = vv.intr[3];
  const double k1 = vv.intr[4], k2 = vv.intr[5], p1 = vv.intr[6], p2 = vv.intr[7];
  const double iz = 1.0 / z;
  const double xn = Xc[0] * iz, yn = Xc[1] * iz;
  const double r2 = xn * xn + yn * yn;
  const double rad = 1.0 + k1 * r2 + k2 * r2 * r2;
  const double xd = xn * rad + 2.0 * p1 * xn * yn + p2 * (r2 + 2.0 * xn * xn);
  const double yd = yn * rad + p1 * (r2 + 2.0 * yn * yn) + 2.0 * p2 * xn * yn;
  const double is = vv.inv_sigma_px;
  o.r[0] = (fx * xd + cx - (double)uv.x) * is;
  o.r[1] = (fy * yd + cy - (double)uv.y) * is;
  if (!kJac) return;
  const double g = 2.0 * (k1 + 2.0 * k2 * r2);
  const double d00 = rad + xn * xn * g + 2.0 * p1 * yn + 6.0 * p2 * xn;
  const double d01 = xn * yn * g + 2.0 * p1 * xn + 2.0 * p2 * yn;
  const double d11 = rad + yn * yn * g + 6.0 * p1 * yn + 2.0 * p2 * xn;
  // Jpix = diag(fx,fy)/sigma * Dd * [1/z 0 -xn/z ; 0 1/z -yn/z]
  double Jp[6];
  const double sx = fx * is, sy = fy * is;
  Jp[0] = sx * d00 * iz; Jp[1] = sx * d01 * iz; Jp[2] = -sx * (d00 * xn + d01 * yn) * iz;
  Jp[3] = sy * d01 * iz; Jp[4] = sy * d11 * iz; Jp[5] = -sy * (d01 * xn + d11 * yn) * iz;
  // rotation matrix of the unit quaternion
  double R[9];
  R[0] = 1 - 2 * (q2 * q2 + q3 * q3); R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = 1 - 2 * (q1 * q1 + q3 * q3); R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = 1 - 2 * (q1 * q1 + q2 * q2);
#pragma unroll
  for (int rho = 0; rho < 2; ++rho)
#pragma unroll
    for (int m = 0; m < 3; ++m)
      o.JX[3 * rho + m] = Jp[3 * rho] * R[m] + Jp[3 * rho + 1] * R[3 + m] + Jp[3 * rho + 2] * R[6 + m];
  // ambient d(RX)/dq (3x4): col0 = 2 v x X ; cols1..3 = -2w hat(X) - 2 hat(v x X) - 2 X v^T + 2 (v.X) I
  double Ja[12];
  const double vX = dot3(v, X);
  double hX[9], hC[9];
  hat3(X, hX);
  hat3(vxX, hC);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Ja[4 * i] = 2.0 * vxX[i];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Ja[4 * i + 1 + j] = -2.0 * q0 * hX[3 * i + j] - 2.0 * hC[3 * i + j] - 2.0 * X[i] * v[j] + ((i == j) ? 2.0 * vX : 0.0);
  }
  // EigenQuaternionManifold::PlusJacobian on memory (m0..m3) = (w,x,y,z)  (Q9)
  const double m0 = q0, m1 = q1, m2 = q2, m3 = q3;
  const double PJ[12] = {m3, m2, -m1,   -m2, m3, m0,   m1, -m0, m3,   -m0, -m1, -m2};
  double Jt3[9];   // 3x3 = Ja (3x4) * PJ (4x3)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Jt3[3 * i + j] = Ja[4 * i] * PJ[j] + Ja[4 * i + 1] * PJ[3 + j] + Ja[4 * i + 2] * PJ[6 + j] + Ja[4 * i + 3] * PJ[9 + j];
#pragma unroll
  for (int rho = 0; rho < 2; ++rho)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      o.Jc[6 * rho + j] = Jp[3 * rho] * Jt3[j] + Jp[3 * rho + 1] * Jt3[3 + j] + Jp[3 * rho + 2] * Jt3[6 + j];
      o.Jc[6 * rho + 3 + j] = Jp[3 * rho + j];
    }
}

// utils.hpp:133-139
LVBA_DEV void plane_eval(const VisualView& vv, const double* pl, const double* X, double& r, double J[3]) {
  const double e = -(pl[0] * X[0] + pl[1] * X[1] + pl[2] * X[2] + pl[3]);
  const double root = sqrt(e * e + 1e-12);
  r = root * vv.inv_sigma_pl;
  const double k = (e / root) * vv.inv_sigma_pl;
  J[0] = -k * pl[0]; J[1] = -k * pl[1]; J[2] = -k * pl[2];
}

LVBA_DEV void sym3_inverse(const double* c /*xx xy xz yy yz zz*/, double* o) {
  const double a = c[0], b = c[1], cc = c[2], d = c[3], e = c[4], f = c[5];
  const double A = d * f - e * e, B = cc * e - b * f, Cc = b * e - cc * d;
  const double det = a * A + b * B + cc * Cc;
  const double id = 1.0 / det;
  o[0] = A * id; o[1] = B * id; o[2] = Cc * id;
  o[3] = (a * f - cc * cc) * id; o[4] = (b * cc - a * e) * id; o[5] = (a * d - b * b) * id;
}

struct VisualLM {
  double radius, min_diag, max_diag;
  const double* s_cam;   // [n_rows*6] Jacobi scale of the camera tangent columns
  const double* s_pt;    // [Tv*3]
};

// ------------------------------------------------------------------------------------------------
// phase 1+2 shared by build and back-substitution.  Leaves per-landmark params in sT:
//   [0..5] Cinv  [6..8] g_p  [9..11] w = Cinv g_p ; returns this thread's observation in `o` (scaled)
template <bool kNeedJac>
LVBA_DEV void visual_front(const VisualView& vv, const VisualState& st, const VisualLM& lm, int b,
                           double* stage, double* sT, ObsEval& o, int& row, int& ltrk, double* Xown,
                           double& cost_part) {
  const int tid = threadIdx.x;
  const int t0 = vv.batch_trk[b], t1 = vv.batch_trk[b + 1], nt = t1 - t0;
  const int s0 = vv.trk_ptr[t0], ns = vv.trk_ptr[t1] - s0;
  __shared__ unsigned char sTrkOf[kSlots];
  if (tid < nt) {
    const int lo = vv.trk_ptr[t0 + tid] - s0, hi = vv.trk_ptr[t0 + tid + 1] - s0;
    for (int q = lo; q < hi; ++q) sTrkOf[q] = (unsigned char)tid;
  }
  __syncthreads();
  cost_part = 0.0;
  row = -1; ltrk = 0;
  if (tid < ns) {
    ltrk = sTrkOf[tid];
    const int tr = vv.trk_id[t0 + ltrk];
    const int cam = vv.obs_cam[s0 + tid];
    row = vv.obs_row[s0 + tid];
    Xown[0] = st.X[3 * (long long)tr]; Xown[1] = st.X[3 * (long long)tr + 1]; Xown[2] = st.X[3 * (long long)tr + 2];
    obs_eval<kNeedJac>(vv, st.q + 4 * (long long)cam, st.t + 3 * (long long)cam, Xown, vv.obs_uv[s0 + tid], o);
    cost_part = o.r[0] * o.r[0] + o.r[1] * o.r[1];
    if (kNeedJac) {
      const double* sp = lm.s_pt + 3 * (long long)(t0 + ltrk);
#pragma unroll
      for (int rho = 0; rho < 2; ++rho)
#pragma unroll
        for (int m = 0; m < 3; ++m) o.JX[3 * rho + m] *= sp[m];
      if (row >= 0) {
        const double* sc = lm.s_cam + 6 * (long long)row;
#pragma unroll
        for (int rho = 0; rho < 2; ++rho)
#pragma unroll
          for (int a = 0; a < 6; ++a) o.Jc[6 * rho + a] *= sc[a];
      } else {
#pragma unroll
        for (int a = 0; a < 12; ++a) o.Jc[a] = 0.0;     // constant / unused camera: no columns
      }
      double* sg = stage + tid * kStageStride;
      sg[0] = o.JX[0] * o.JX[0] + o.JX[3] * o.JX[3];
      sg[1] = o.JX[0] * o.JX[1] + o.JX[3] * o.JX[4];
      sg[2] = o.JX[0] * o.JX[2] + o.JX[3] * o.JX[5];
      sg[3] = o.JX[1] * o.JX[1] + o.JX[4] * o.JX[4];
      sg[4] = o.JX[1] * o.JX[2] + o.JX[4] * o.JX[5];
      sg[5] = o.JX[2] * o.JX[2] + o.JX[5] * o.JX[5];
      sg[6] = o.JX[0] * o.r[0] + o.JX[3] * o.r[1];
      sg[7] = o.JX[1] * o.r[0] + o.JX[4] * o.r[1];
      sg[8] = o.JX[2] * o.r[0] + o.JX[5] * o.r[1];
    }
  }
  __syncthreads();
  if (tid < nt) {
    const int lo = vv.trk_ptr[t0 + tid] - s0, hi = vv.trk_ptr[t0 + tid + 1] - s0;
    const int tr = vv.trk_id[t0 + tid];
    const double Xp[3] = {st.X[3 * (long long)tr], st.X[3 * (long long)tr + 1], st.X[3 * (long long)tr + 2]};
    double rp, Jpl[3];
    plane_eval(vv, vv.plane + 4 * (long long)(t0 + tid), Xp, rp, Jpl);
    cost_part += rp * rp;
    if (kNeedJac) {
      const double* sp = lm.s_pt + 3 * (long long)(t0 + tid);
      Jpl[0] *= sp[0]; Jpl[1] *= sp[1]; Jpl[2] *= sp[2];
      double acc[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) acc[q] = 0.0;
      for (int s = lo; s < hi; ++s)
#pragma unroll
        for (int q = 0; q < 9; ++q) acc[q] += stage[s * kStageStride + q];
      acc[0] += Jpl[0] * Jpl[0]; acc[1] += Jpl[0] * Jpl[1]; acc[2] += Jpl[0] * Jpl[2];
      acc[3] += Jpl[1] * Jpl[1]; acc[4] += Jpl[1] * Jpl[2]; acc[5] += Jpl[2] * Jpl[2];
      acc[6] += Jpl[0] * rp; acc[7] += Jpl[1] * rp; acc[8] += Jpl[2] * rp;
      // LM diagonal of the point columns: clamp(||J~[:,j]||^2)/radius   (Ceres LevenbergMarquardtStrategy)
      const double ir = 1.0 / lm.radius;
      acc[0] += fmin(fmax(acc[0], lm.min_diag), lm.max_diag) * ir;
      acc[3] += fmin(fmax(acc[3], lm.min_diag), lm.max_diag) * ir;
      acc[5] += fmin(fmax(acc[5], lm.min_diag), lm.max_diag) * ir;
      double* p = sT + tid * kTrkParams;
      sym3_inverse(acc, p);
      p[6] = acc[6]; p[7] = acc[7]; p[8] = acc[8];
      p[9] = p[0] * acc[6] + p[1] * acc[7] + p[2] * acc[8];
      p[10] = p[1] * acc[6] + p[3] * acc[7] + p[4] * acc[8];
      p[11] = p[2] * acc[6] + p[4] * acc[7] + p[5] * acc[8];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// 1/2 sum r^2 partial per batch (candidate evaluation)
__global__ void __launch_bounds__(kSlots)
visual_cost_kernel(VisualView vv, VisualState st, double* __restrict__ batch_cost) {
  __shared__ double red[32];
  __shared__ double dummyT[1];
  ObsEval o;
  int row, ltrk;
  double Xown[3], cp;
  VisualLM lm{1.0, 0.0, 0.0, nullptr, nullptr};
  visual_front<false>(vv, st, lm, blockIdx.x, nullptr, dummyT, o, row, ltrk, Xown, cp);
  const double tot = block_sum<kSlots>(cp, red);
  if (threadIdx.x == 0) batch_cost[blockIdx.x] = 0.5 * tot;
}

// ------------------------------------------------------------------------------------------------
// unscaled squared column norms (Jacobi scaling, computed once at iteration 0)
__global__ void __launch_bounds__(kSlots)
visual_colnorm_kernel(VisualView vv, VisualState st, double* __restrict__ cam_colsq, double* __restrict__ pt_colsq) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int t0 = vv.batch_trk[b], t1 = vv.batch_trk[b + 1], nt = t1 - t0;
  const int s0 = vv.trk_ptr[t0], ns = vv.trk_ptr[t1] - s0;
  __shared__ double stage[kSlots * 4];
  if (tid < ns) {
    const int lt_lo = 0; (void)lt_lo;
    // locate landmark by scanning is avoided: recompute from trk_ptr with a small search
    int lo = 0, hi = nt;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (vv.trk_ptr[t0 + mid] - s0 <= tid) lo = mid; else hi = mid; }
    const int tr = vv.trk_id[t0 + lo];
    const int cam = vv.obs_cam[s0 + tid];
    const int row = vv.obs_row[s0 + tid];
    const double X[3] = {st.X[3 * (long long)tr], st.X[3 * (long long)tr + 1], st.X[3 * (long long)tr + 2]};
    ObsEval o;
    obs_eval<true>(vv, st.q + 4 * (long long)cam, st.t + 3 * (long long)cam, X, vv.obs_uv[s0 + tid], o);
    if (row >= 0)
#pragma unroll
      for (int a = 0; a < 6; ++a) atomicAdd(cam_colsq + 6 * (long long)row + a, o.Jc[a] * o.Jc[a] + o.Jc[6 + a] * o.Jc[6 + a]);
    stage[tid * 4 + 0] = o.JX[0] * o.JX[0] + o.JX[3] * o.JX[3];
    stage[tid * 4 + 1] = o.JX[1] * o.JX[1] + o.JX[4] * o.JX[4];
    stage[tid * 4 + 2] = o.JX[2] * o.JX[2] + o.JX[5] * o.JX[5];
  }
  __syncthreads();
  if (tid < nt) {
    const int lo = vv.trk_ptr[t0 + tid] - s0, hi = vv.trk_ptr[t0 + tid + 1] - s0;
    const int tr = vv.trk_id[t0 + tid];
    const double X[3] = {st.X[3 * (long long)tr], st.X[3 * (long long)tr + 1], st.X[3 * (long long)tr + 2]};
    double rp, J[3];
    plane_eval(vv, vv.plane + 4 * (long long)(t0 + tid), X, rp, J);
    double a0 = J[0] * J[0], a1 = J[1] * J[1], a2 = J[2] * J[2];
    for (int s = lo; s < hi; ++s) { a0 += stage[s * 4]; a1 += stage[s * 4 + 1]; a2 += stage[s * 4 + 2]; }
    double* o = pt_colsq + 3 * (long long)(t0 + tid);
    o[0] = a0; o[1] = a1; o[2] = a2;
  }
}

// scale = 1 / (1 + sqrt(colsq))  (Ceres jacobi_scaling) ; or 1 when disabled
__global__ void visual_scale_kernel(long long n, const double* __restrict__ colsq, int enabled, double* __restrict__ scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) scale[i] = enabled ? 1.0 / (1.0 + sqrt(colsq[i])) : 1.0;
}

// ------------------------------------------------------------------------------------------------
// linearise + Schur-eliminate: S (envelope, lower), rhs, camera column norms, camera gradient
__global__ void __launch_bounds__(kSlots)
visual_build_kernel(VisualView vv, EnvView env, VisualState st, VisualLM lm, double* __restrict__ S,
                    double* __restrict__ rhs, double* __restrict__ cam_colsq, double* __restrict__ cam_grad,
                    double* __restrict__ batch_cost, double* __restrict__ batch_gmax) {
  extern __shared__ double sm[];
  double* stage = sm;                                    // [kSlots][kStageStride]
  double* sF = stage + kSlots * kStageStride;            // [kSlots][kVFStride]   Y | E
  double* sG = sF + kSlots * kVFStride;                  // [kSlots][18]  rhs | colsq | grad
  double* sT = sG + kSlots * 18;                         // [kMaxTrkPerBatch][kTrkParams]
  double* red = sT + kMaxTrkPerBatch * kTrkParams;       // [32]
  long long* sDiag = reinterpret_cast<long long*>(red + 32);     // [kSlots]
  long long* sRowRS = sDiag + kSlots;                            // [kSlots] envelope row_start of the camera row
  int* sRow = reinterpret_cast<int*>(sRowRS + kSlots);           // [kSlots]
  int* sRowFirst = sRow + kSlots;                                // [kSlots]

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = vv.batch_trk[b], t1 = vv.batch_trk[b + 1], nt = t1 - t0;
  const int s0 = vv.trk_ptr[t0], ns = vv.trk_ptr[t1] - s0;
  ObsEval o;
  int row, ltrk;
  double Xown[3], cp;
  visual_front<true>(vv, st, lm, b, stage, sT, o, row, ltrk, Xown, cp);
  // gradient max-norm (unscaled) of the point blocks
  double gm = 0.0;
  if (tid < nt) {
    const double* p = sT + tid * kTrkParams;
    const double* sp = lm.s_pt + 3 * (long long)(t0 + tid);
    gm = fmax(fabs(p[6] / sp[0]), fmax(fabs(p[7] / sp[1]), fabs(p[8] / sp[2])));
  }
  {
    const double tot = block_sum<kSlots>(cp, red);
    if (tid == 0) batch_cost[b] = 0.5 * tot;
    // max via the same scratch
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, off));
    if (lane == 0) red[warp] = gm;
    __syncthreads();
    if (tid == 0) {
      double m = 0.0;
      for (int w = 0; w < kSlots / 32; ++w) m = fmax(m, red[w]);
      batch_gmax[b] = m;
    }
    __syncthreads();
  }

  // ---- phase 3
  if (tid < ns) {
    sRow[tid] = row;
    double* sf = sF + tid * kVFStride;
    double* sg = sG + tid * 18;
    double* sd = stage + tid * kStageStride;
    if (row >= 0) {
      const double* p = sT + ltrk * kTrkParams;
      const double Ci[9] = {p[0], p[1], p[2], p[1], p[3], p[4], p[2], p[4], p[5]};
      double E[18], Y[18];
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) E[3 * a + m] = o.Jc[a] * o.JX[m] + o.Jc[6 + a] * o.JX[3 + m];
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) Y[3 * a + m] = E[3 * a] * Ci[m] + E[3 * a + 1] * Ci[3 + m] + E[3 * a + 2] * Ci[6 + m];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double gc = o.Jc[a] * o.r[0] + o.Jc[6 + a] * o.r[1];
        sg[a] = -(gc - (E[3 * a] * p[9] + E[3 * a + 1] * p[10] + E[3 * a + 2] * p[11]));
        sg[6 + a] = o.Jc[a] * o.Jc[a] + o.Jc[6 + a] * o.Jc[6 + a];
        sg[12 + a] = gc;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          sd[6 * a + c] = o.Jc[a] * o.Jc[c] + o.Jc[6 + a] * o.Jc[6 + c]
                        - (Y[3 * a] * E[3 * c] + Y[3 * a + 1] * E[3 * c + 1] + Y[3 * a + 2] * E[3 * c + 2]);
      }
#pragma unroll
      for (int q = 0; q < 18; ++q) { sf[q] = Y[q]; sf[18 + q] = E[q]; }
      const long long rs = env.row_start[row];
      const int fr = env.first[row];
      sRowRS[tid] = rs; sRowFirst[tid] = fr;
      sDiag[tid] = (rs + (row - fr)) * 36;
    } else {
      sDiag[tid] = -1;
    }
  }
  __syncthreads();
  // ---- phase 4a: flush diagonal blocks, rhs, column norms, gradient
  for (int e = tid; e < ns * 36; e += kSlots) {
    const int sl = e / 36, el = e - sl * 36;
    if (sDiag[sl] >= 0) atomicAdd(S + sDiag[sl] + el, stage[sl * kStageStride + el]);
  }
  for (int e = tid; e < ns * 18; e += kSlots) {
    const int sl = e / 18, el = e - sl * 18;
    const int r = sRow[sl];
    if (r >= 0) {
      const int which = el / 6, a = el - which * 6;
      double* dst = (which == 0) ? rhs : (which == 1) ? cam_colsq : cam_grad;
      atomicAdd(dst + 6 * (long long)r + a, sG[e]);
    }
  }
  // ---- phase 4b: camera pairs.  Lane (pq, k) = (lane >> 2, lane & 3) owns the elements 4 m + k, m = 0..8, of pair pq (the
  //      mapping of lidar_build_kernel's phase 4b): one aligned 32-byte sector per pair and instruction, factors read once.
  const long long p0 = vv.batch_pair[b], np = vv.batch_pair[b + 1] - p0;
  const int pq = lane >> 2, k = lane & 3;
  const bool upper = k >= 2;
  const int col0 = k, col1 = upper ? k - 2 : 4 + k, col2 = 2 + k;
  unsigned code_next = 0;
  {
    const long long c0 = (long long)warp * 8;
    if (c0 + pq < np) code_next = vv.pairs[p0 + c0 + pq];
  }
  for (long long c = (long long)warp * 8; c < np; c += (kSlots / 32) * 8) {
    const bool live = c + pq < np;
    const unsigned code = code_next;
    {
      const long long cn = c + (kSlots / 32) * 8;
      code_next = (cn + pq < np) ? vv.pairs[p0 + cn + pq] : 0u;
    }
    if (live) {
      const int hi = code & 0xff, lo = (code >> 8) & 0xff;
      double* dst = S + (sRowRS[hi] + (sRow[lo] - sRowFirst[hi])) * 36 + k;
      const double* y = sF + hi * kVFStride;               // Y_hi, row a at 3 a
      const double* ee = sF + lo * kVFStride + 18;         // E_lo, row bq at 3 bq
      double ce[3][3], ry[6][3];
#pragma unroll
      for (int t = 0; t < 3; ++t) { ce[0][t] = ee[3 * col0 + t]; ce[1][t] = ee[3 * col1 + t]; ce[2][t] = ee[3 * col2 + t]; }
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int t = 0; t < 3; ++t) ry[a][t] = y[3 * a + t];
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        const int j2 = 2 * (m / 3), ph = m % 3;
        double r0, r1, r2;
        if (ph == 0) { r0 = ry[j2][0]; r1 = ry[j2][1]; r2 = ry[j2][2]; }
        else if (ph == 2) { r0 = ry[j2 + 1][0]; r1 = ry[j2 + 1][1]; r2 = ry[j2 + 1][2]; }
        else { r0 = upper ? ry[j2 + 1][0] : ry[j2][0]; r1 = upper ? ry[j2 + 1][1] : ry[j2][1]; r2 = upper ? ry[j2 + 1][2] : ry[j2][2]; }
        const double val = -(r0 * ce[ph][0] + r1 * ce[ph][1] + r2 * ce[ph][2]);
        atomicAdd(dst + 4 * m, val);
      }
    }
  }
}

constexpr size_t visual_build_smem_bytes() {
  return sizeof(double) * (kSlots * kStageStride + kSlots * kVFStride + kSlots * 18 + kMaxTrkPerBatch * kTrkParams + 32)
       + sizeof(long long) * 2 * kSlots + sizeof(int) * 2 * kSlots + 64;
}

// ------------------------------------------------------------------------------------------------
// back-substitution y_p = -C^-1 (g_p + sum E^T y_c), model cost change, candidate landmarks
__global__ void __launch_bounds__(kSlots)
visual_backsub_kernel(VisualView vv, VisualState st, VisualLM lm, const double* __restrict__ y_cam,
                      double* __restrict__ X_cand, double* __restrict__ pt_step,
                      double* __restrict__ batch_out /* [n_batches][4]: model, step^2, x^2, - */) {
  extern __shared__ double sm[];
  double* stage = sm;                                    // [kSlots][kStageStride]
  double* sT = stage + kSlots * kStageStride;            // [kMaxTrkPerBatch][kTrkParams]
  double* red = sT + kMaxTrkPerBatch * kTrkParams;       // [32]
  const int b = blockIdx.x, tid = threadIdx.x;
  const int t0 = vv.batch_trk[b], t1 = vv.batch_trk[b + 1], nt = t1 - t0;
  const int s0 = vv.trk_ptr[t0], ns = vv.trk_ptr[t1] - s0;
  ObsEval o;
  int row, ltrk;
  double Xown[3], cp;
  visual_front<true>(vv, st, lm, b, stage, sT, o, row, ltrk, Xown, cp);
  double yc[6] = {0, 0, 0, 0, 0, 0};
  if (tid < ns) {
    if (row >= 0)
#pragma unroll
      for (int a = 0; a < 6; ++a) yc[a] = y_cam[6 * (long long)row + a];
    // E^T y_c = J_X^T (J_c y_c)
    const double j0 = o.Jc[0] * yc[0] + o.Jc[1] * yc[1] + o.Jc[2] * yc[2] + o.Jc[3] * yc[3] + o.Jc[4] * yc[4] + o.Jc[5] * yc[5];
    const double j1 = o.Jc[6] * yc[0] + o.Jc[7] * yc[1] + o.Jc[8] * yc[2] + o.Jc[9] * yc[3] + o.Jc[10] * yc[4] + o.Jc[11] * yc[5];
    double* sg = stage + tid * kStageStride;
    sg[0] = o.JX[0] * j0 + o.JX[3] * j1;
    sg[1] = o.JX[1] * j0 + o.JX[4] * j1;
    sg[2] = o.JX[2] * j0 + o.JX[5] * j1;
    sg[3] = j0; sg[4] = j1;
  }
  __syncthreads();
  double model = 0.0, step2 = 0.0, x2 = 0.0;
  if (tid < nt) {
    const int lo = vv.trk_ptr[t0 + tid] - s0, hi = vv.trk_ptr[t0 + tid + 1] - s0;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int s = lo; s < hi; ++s) { a0 += stage[s * kStageStride]; a1 += stage[s * kStageStride + 1]; a2 += stage[s * kStageStride + 2]; }
    double* p = sT + tid * kTrkParams;
    const double y0 = -(p[9] + p[0] * a0 + p[1] * a1 + p[2] * a2);
    const double y1 = -(p[10] + p[1] * a0 + p[3] * a1 + p[4] * a2);
    const double y2 = -(p[11] + p[2] * a0 + p[4] * a1 + p[5] * a2);
    p[12] = y0; p[13] = y1; p[14] = y2;
    const int tr = vv.trk_id[t0 + tid];
    const double* sp = lm.s_pt + 3 * (long long)(t0 + tid);
    const double X[3] = {st.X[3 * (long long)tr], st.X[3 * (long long)tr + 1], st.X[3 * (long long)tr + 2]};
    const double d[3] = {sp[0] * y0, sp[1] * y1, sp[2] * y2};
    X_cand[3 * (long long)tr] = X[0] + d[0]; X_cand[3 * (long long)tr + 1] = X[1] + d[1]; X_cand[3 * (long long)tr + 2] = X[2] + d[2];
    if (pt_step) { pt_step[3 * (long long)tr] = d[0]; pt_step[3 * (long long)tr + 1] = d[1]; pt_step[3 * (long long)tr + 2] = d[2]; }
    step2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    x2 = X[0] * X[0] + X[1] * X[1] + X[2] * X[2];
    // plane residual contribution to the model cost change
    double rp, Jpl[3];
    plane_eval(vv, vv.plane + 4 * (long long)(t0 + tid), X, rp, Jpl);
    const double jy = Jpl[0] * d[0] + Jpl[1] * d[1] + Jpl[2] * d[2];   // J~ y = J (s o y)
    model -= jy * (rp + 0.5 * jy);
  }
  __syncthreads();
  if (tid < ns) {
    const double* p = sT + ltrk * kTrkParams;
    const double* sg = stage + tid * kStageStride;
    const double jy0 = sg[3] + o.JX[0] * p[12] + o.JX[1] * p[13] + o.JX[2] * p[14];
    const double jy1 = sg[4] + o.JX[3] * p[12] + o.JX[4] * p[13] + o.JX[5] * p[14];
    model -= jy0 * (o.r[0] + 0.5 * jy0) + jy1 * (o.r[1] + 0.5 * jy1);
  }
  const double tm = block_sum<kSlots>(model, red);
  const double ts = block_sum<kSlots>(step2, red);
  const double tx = block_sum<kSlots>(x2, red);
  if (tid == 0) { batch_out[4 * b] = tm; batch_out[4 * b + 1] = ts; batch_out[4 * b + 2] = tx; batch_out[4 * b + 3] = 0.0; }
}

constexpr size_t visual_backsub_smem_bytes() {
  return sizeof(double) * (kSlots * kStageStride + kMaxTrkPerBatch * kTrkParams + 32) + 64;
}

// camera candidates: q <- Plus(q, s o y[0:3]) with the Q9 manifold, t <- t + s o y[3:6].
// out[2 rows] partial sums: step^2 and x^2 in the ambient space (Ceres ParameterToleranceReached).
__global__ void visual_cam_update_kernel(int n_rows, const int* __restrict__ cam_of_row, const double* __restrict__ q,
                                         const double* __restrict__ t, const double* __restrict__ y,
                                         const double* __restrict__ s_cam, double* __restrict__ q_cand,
                                         double* __restrict__ t_cand, double* __restrict__ cam_step,
                                         double* __restrict__ out /* [gridDim][2] */) {
  __shared__ double red[32];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double step2 = 0.0, x2 = 0.0;
  if (r < n_rows) {
    const int c = cam_of_row[r];
    const double* yy = y + 6 * (long long)r;
    const double* sc = s_cam + 6 * (long long)r;
    const double d[6] = {sc[0] * yy[0], sc[1] * yy[1], sc[2] * yy[2], sc[3] * yy[3], sc[4] * yy[4], sc[5] * yy[5]};
    const double m0 = q[4 * c], m1 = q[4 * c + 1], m2 = q[4 * c + 2], m3 = q[4 * c + 3];
    const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double n0 = m0, n1 = m1, n2 = m2, n3 = m3;
    if (nd > 0.0) {
      double sn, cs;
      sincos(nd, &sn, &cs);
      const double k = sn / nd;
      const double s0 = k * d[0], s1 = k * d[1], s2 = k * d[2];
      // (cs; s) (x) (m3; m0,m1,m2)  — Hamilton product, Eigen memory order (Q9)
      n3 = cs * m3 - (s0 * m0 + s1 * m1 + s2 * m2);
      n0 = cs * m0 + m3 * s0 + (s1 * m2 - s2 * m1);
      n1 = cs * m1 + m3 * s1 + (s2 * m0 - s0 * m2);
      n2 = cs * m2 + m3 * s2 + (s0 * m1 - s1 * m0);
    }
    q_cand[4 * c] = n0; q_cand[4 * c + 1] = n1; q_cand[4 * c + 2] = n2; q_cand[4 * c + 3] = n3;
    const double t0 = t[3 * c], t1 = t[3 * c + 1], t2 = t[3 * c + 2];
    t_cand[3 * c] = t0 + d[3]; t_cand[3 * c + 1] = t1 + d[4]; t_cand[3 * c + 2] = t2 + d[5];
    if (cam_step)
#pragma unroll
      for (int a = 0; a < 6; ++a) cam_step[6 * (long long)c + a] = d[a];
    step2 = (n0 - m0) * (n0 - m0) + (n1 - m1) * (n1 - m1) + (n2 - m2) * (n2 - m2) + (n3 - m3) * (n3 - m3)
          + d[3] * d[3] + d[4] * d[4] + d[5] * d[5];
    x2 = m0 * m0 + m1 * m1 + m2 * m2 + m3 * m3 + t0 * t0 + t1 * t1 + t2 * t2;
  }
  const double ts = block_sum<128>(step2, red);
  const double tx = block_sum<128>(x2, red);
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = ts; out[2 * blockIdx.x + 1] = tx; }
}

// dadd = clamp(colsq, min, max) / radius  (camera LM diagonal) ; also camera gradient max-norm (unscaled)
__global__ void visual_cam_diag_kernel(int n6, const double* __restrict__ colsq, const double* __restrict__ grad,
                                       const double* __restrict__ s_cam, double min_d, double max_d, double radius,
                                       double* __restrict__ dadd, double* __restrict__ gmax_out) {
  __shared__ double red[32];
  double gm = 0.0;
  for (int i = threadIdx.x; i < n6; i += blockDim.x) {
    dadd[i] = fmin(fmax(colsq[i], min_d), max_d) / radius;
    gm = fmax(gm, fabs(grad[i] / s_cam[i]));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    gmax_out[0] = m;
  }
}

// generic strided partial reductions: out[j] = sum_i part[i*stride + j] (j < ncol) ; max variant
__global__ void reduce_cols_kernel(const double* __restrict__ part, int n, int stride, int ncol, double* __restrict__ out) {
  __shared__ double red[32];
  for (int j = 0; j < ncol; ++j) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += part[(long long)i * stride + j];
    const double tot = block_sum<256>(s, red);
    if (threadIdx.x == 0) out[j] = tot;
  }
}
__global__ void reduce_max_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) m = fmax(m, part[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int w = 0; w < 8; ++w) r = fmax(r, red[w]);
    out[0] = fmax(out[0], r);
  }
}

// multi-GPU gather of the landmark updates: every rank owns a disjoint set of landmarks
__global__ void visual_delta_kernel(long long Tv, const int* __restrict__ trk_id, const double* __restrict__ X,
                                    const double* __restrict__ X0, double* __restrict__ delta) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= Tv * 3) return;
  const long long k = i / 3, q = i - 3 * k;
  const long long tr = trk_id[k];
  delta[3 * tr + q] = X[3 * tr + q] - X0[3 * tr + q];
}
__global__ void visual_add_kernel(long long n, const double* __restrict__ X0, const double* __restrict__ delta, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = X0[i] + delta[i];
}

}  // namespace lvba
