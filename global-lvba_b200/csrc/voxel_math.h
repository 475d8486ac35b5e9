// voxel_math.h — per-point / per-cluster arithmetic of the set-up stage (boundary B3):
// root-voxel key, octant descent, PointCluster push / transform / covariance, plane test.
//
// Everything here is plain C++ that compiles both as device code (nvcc) and as host code (g++): the same functions
// run inside the CUDA kernels of voxel_api.cuh and inside the host emulation the CPU tests use to check the
// pipeline without a GPU (tests/emu/voxel_emu.cpp).  To make the emulation an exact predictor of the device,
// every product and sum goes through mul_/add_/sub_ (round-to-nearest, never contracted into an FMA); division and
// sqrt are IEEE on both sides.
//
// Reference (xuankuzcr/Global-LVBA):
//   cut_voxel            include/BALM/bavoxel.hpp:799-836
//   cut_func             include/BALM/bavoxel.hpp:391-418
//   judge_eigen          include/BALM/bavoxel.hpp:335-352
//   PointCluster         include/BALM/tools.hpp:407-456
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LVBA_HD __host__ __device__ __forceinline__
#else
#define LVBA_HD inline
#endif

namespace lvba {
namespace vox {

LVBA_HD double mul_(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
LVBA_HD double add_(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
LVBA_HD double sub_(double a, double b) { return add_(a, -b); }
LVBA_HD float fadd_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
LVBA_HD float fmul_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}

struct VoxParams {
  double voxel_size;
  float eigen_ratio[4];   // eigen_ratio_array, bavoxel.hpp:17-22 (float in the reference; compared in double)
  int layer_limit;        // bavoxel.hpp:13
  int min_points;         // min_ps, bavoxel.hpp:24
};

constexpr int64_t kKeyLimit = (int64_t)1 << 30;   // root keys beyond +-2^30 voxels per axis are refused

// pw = R * p + t, each coefficient (r0 p0 + r1 p1) + r2 p2, then + t   (bavoxel.hpp:806-807, :399)
LVBA_HD void world_point(const double* pose, const float* p, double w[3]) {
  const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
  for (int k = 0; k < 3; ++k)
    w[k] = add_(add_(add_(mul_(pose[3 * k], x), mul_(pose[3 * k + 1], y)), mul_(pose[3 * k + 2], z)), pose[9 + k]);
}

// loc = (float)(pw / voxel_size); if (loc < 0) loc -= 1.0; key = (int64)loc      (bavoxel.hpp:810-816)
// returns false for a non-finite coordinate or a key outside +-kKeyLimit (the reference's cast would be undefined).
LVBA_HD bool root_key_axis(double w, double voxel_size, int64_t* key) {
  float loc = (float)(w / voxel_size);
  if (loc < 0) loc = (float)((double)loc - 1.0);
  if (!(loc > -(float)kKeyLimit && loc < (float)kKeyLimit)) return false;
  *key = (int64_t)loc;
  return true;
}

// voxel_center = (0.5 + key) * voxel_size stored to float; quater_length = voxel_size / 4 (float)   (:829-832)
LVBA_HD float root_centre_axis(int64_t key, double voxel_size) { return (float)mul_(add_(0.5, (double)key), voxel_size); }
LVBA_HD float root_quater(double voxel_size) { return (float)(voxel_size / 4.0); }

// cut_func :399-403 — strict '>' of the double coordinate against the float centre; leafnum = 4x + 2y + z
LVBA_HD int octant(const double w[3], const float c[3], int bits[3]) {
  for (int k = 0; k < 3; ++k) bits[k] = (w[k] > (double)c[k]) ? 1 : 0;
  return 4 * bits[0] + 2 * bits[1] + bits[2];
}
// :407-410 — child centre = centre + (2 bit - 1) * quater_length in float; child quater = quater / 2
LVBA_HD void child_centre(const float c[3], const int bits[3], float quater, float out[3]) {
  for (int k = 0; k < 3; ++k) out[k] = fadd_(c[k], fmul_((float)(2 * bits[k] - 1), quater));
}
LVBA_HD void octant_bits(int o, int bits[3]) { bits[0] = (o >> 2) & 1; bits[1] = (o >> 1) & 1; bits[2] = o & 1; }

// ---------------------------------------------------------------- PointCluster: 10 doubles Pxx Pxy Pxz Pyy Pyz Pzz vx vy vz N
LVBA_HD void cluster_zero(double c[10]) { for (int k = 0; k < 10; ++k) c[k] = 0.0; }
// push(vec): N++, P += vec vec^T, v += vec      (tools.hpp:427-432)
LVBA_HD void cluster_push(double c[10], const float* p) {
  const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
  c[0] = add_(c[0], mul_(x, x)); c[1] = add_(c[1], mul_(x, y)); c[2] = add_(c[2], mul_(x, z));
  c[3] = add_(c[3], mul_(y, y)); c[4] = add_(c[4], mul_(y, z)); c[5] = add_(c[5], mul_(z, z));
  c[6] = add_(c[6], x); c[7] = add_(c[7], y); c[8] = add_(c[8], z);
  c[9] = add_(c[9], 1.0);
}

LVBA_HD double dot3_(const double* a, const double* b) { return add_(add_(mul_(a[0], b[0]), mul_(a[1], b[1])), mul_(a[2], b[2])); }

// transform(sigv, stat) (tools.hpp:443-449) of a body-frame cluster into the world frame, ADDED to (Pm, vm, Nm):
//   v' = R v + N p ;  rp = (R v) p^T ;  P' = R P R^T + rp + rp^T + N p p^T
LVBA_HD void cluster_transform_add(const double c[10], const double* pose, double Pm[9], double vm[3], double* Nm) {
  const double* R = pose;
  const double* t = pose + 9;
  const double P[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
  const double N = c[9];
  double Rv[3], RP[9];
  for (int i = 0; i < 3; ++i) Rv[i] = dot3_(R + 3 * i, c + 6);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      RP[3 * i + j] = add_(add_(mul_(R[3 * i], P[j]), mul_(R[3 * i + 1], P[3 + j])), mul_(R[3 * i + 2], P[6 + j]));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double rpr = dot3_(RP + 3 * i, R + 3 * j);                                   // (R P) R^T
      const double v = add_(add_(add_(rpr, mul_(Rv[i], t[j])), mul_(Rv[j], t[i])), mul_(mul_(N, t[i]), t[j]));
      Pm[3 * i + j] = add_(Pm[3 * i + j], v);
    }
  for (int i = 0; i < 3; ++i) vm[i] = add_(vm[i], add_(Rv[i], mul_(N, t[i])));
  *Nm = add_(*Nm, N);
}

// ---------------------------------------------------------------- symmetric 3x3 eigen decomposition (cyclic Jacobi)
// Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d> (bavoxel.hpp:346): eigenvalues ascending, u0 = eigenvector of
// the smallest one (its sign is not defined by either solver; every use downstream is even in it).
LVBA_HD void jacobi_rotate(double& app, double& aqq, double& apq, double& apr, double& aqr, double v[3][3], int P, int Q, int& rot) {
  if (apq == 0.0 || !(fabs(apq) > 1e-22 * (fabs(app) + fabs(aqq)))) return;
  const double theta = sub_(aqq, app) / mul_(2.0, apq);
  const double tt = copysign(1.0, theta) / add_(fabs(theta), sqrt(add_(mul_(theta, theta), 1.0)));
  const double c = 1.0 / sqrt(add_(mul_(tt, tt), 1.0)), s = mul_(tt, c);
  app = sub_(app, mul_(tt, apq)); aqq = add_(aqq, mul_(tt, apq)); apq = 0.0;
  const double t1 = sub_(mul_(c, apr), mul_(s, aqr)), t2 = add_(mul_(s, apr), mul_(c, aqr));
  apr = t1; aqr = t2;
  for (int k = 0; k < 3; ++k) {
    const double vp = v[k][P], vq = v[k][Q];
    v[k][P] = sub_(mul_(c, vp), mul_(s, vq)); v[k][Q] = add_(mul_(s, vp), mul_(c, vq));
  }
  ++rot;
}
LVBA_HD void eig3(const double A[9], double lam[3], double u0[3]) {
  double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[4], a12 = A[5], a22 = A[8];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep) {
    int rot = 0;
    jacobi_rotate(a00, a11, a01, a02, a12, v, 0, 1, rot);
    jacobi_rotate(a00, a22, a02, a01, a12, v, 0, 2, rot);
    jacobi_rotate(a11, a22, a12, a01, a02, v, 1, 2, rot);
    if (rot == 0) break;
  }
  double l[3] = {a00, a11, a22};
  int o[3] = {0, 1, 2};
  if (l[o[0]] > l[o[1]]) { int k = o[0]; o[0] = o[1]; o[1] = k; }
  if (l[o[1]] > l[o[2]]) { int k = o[1]; o[1] = o[2]; o[2] = k; }
  if (l[o[0]] > l[o[1]]) { int k = o[0]; o[0] = o[1]; o[1] = k; }
  for (int k = 0; k < 3; ++k) lam[k] = l[o[k]];
  for (int k = 0; k < 3; ++k) u0[k] = v[k][o[0]];
}

// judge_eigen :335-352 on the merged world-frame cluster: centre = v/N, cov = P/N - centre centre^T,
// plane  <=>  !(lambda0 / lambda2 > eigen_ratio_array[layer])   (a NaN ratio therefore counts as a plane, as there).
LVBA_HD bool plane_test(const double Pm[9], const double vm[3], double Nm, float ratio_limit, double centre[3], double direct[3],
                        double lam[3]) {
  double cov[9];
  for (int k = 0; k < 3; ++k) centre[k] = vm[k] / Nm;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[3 * i + j] = sub_(Pm[3 * i + j] / Nm, mul_(centre[i], centre[j]));
  eig3(cov, lam, direct);
  const double ratio = lam[0] / lam[2];
  return !(ratio > (double)ratio_limit);
}

}  // namespace vox
}  // namespace lvba
