// track_pipeline.h — the per-track numerics of the track fusion (second half):
//   TriangulateTrackDLT   src/lvba_system.cpp:52-111     undistort every selected observation (utils.hpp:207-233), two DLT rows
//                                                        per view accumulated into A^T A (4x4), eigenvector of the smallest
//                                                        eigenvalue, dehomogenise, then the mean reprojection error
//   ComputeMeanReproj     src/lvba_system.cpp:8-50       mean pixel distance of a 3-D point over a set of observations
// One item per track; the selection of the observations (one per image, the greedy view-angle filter whose outcome depends on
// the iteration order of std::unordered_map, :995-1000, :1056-1096, :1120-1150) stays with the caller, who passes the selected
// observations of every track as a CSR list.  Same Exec-policy scheme as voxel_pipeline.h / depth_pipeline.h.
#pragma once
#include "depth_pipeline.h"

namespace lvba {
namespace track {

using depth::finite_;

// projectWorldToPixel (utils.hpp:199-205) without the integer truncation of the renderer
LVBA_HD bool project_world(const double* cam, const double* intr, const double* Xw, double* u, double* v) {
  double pc[3];
  depth::affine3(cam, Xw, pc);
  const double Z = pc[2];
  if (!(finite_(pc[0]) && finite_(pc[1]) && finite_(pc[2])) || !(Z > 1e-12)) return false;
  const double x = pc[0] / Z, y = pc[1] / Z;
  if (!(finite_(x) && finite_(y))) return false;
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3], k1 = intr[4], k2 = intr[5], p1 = intr[6], p2 = intr[7];
  const double r2 = x * x + y * y, r4 = r2 * r2;
  const double radial = 1.0 + k1 * r2 + k2 * r4;
  const double x_tan = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
  const double y_tan = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
  const double xd = x * radial + x_tan, yd = y * radial + y_tan;
  if (!(finite_(xd) && finite_(yd))) return false;
  *u = fx * xd + cx; *v = fy * yd + cy;
  return finite_(*u) && finite_(*v);
}

// ComputeMeanReproj (:8-50) over observations [a, b)
LVBA_HD bool mean_reproj(const int32_t* obs_cam, const float* obs_uv, int64_t a, int64_t b, int n_cams, const double* cams,
                         const double* intr, const double* Xw, int min_count, double* mean, int32_t* count) {
  double sum = 0.0;
  int cnt = 0;
  for (int64_t q = a; q < b; ++q) {
    const int c = obs_cam[q];
    if (c < 0 || c >= n_cams) continue;
    double uh, vh;
    if (!project_world(cams + 12 * (int64_t)c, intr, Xw, &uh, &vh)) continue;
    const double du = uh - (double)obs_uv[2 * q], dv = vh - (double)obs_uv[2 * q + 1];
    sum += sqrt(du * du + dv * dv);
    ++cnt;
  }
  *count = cnt;
  if (cnt < min_count) return false;
  *mean = sum / (double)cnt;
  return finite_(*mean);
}

// eigenvector of the smallest eigenvalue of a symmetric 4x4 (cyclic Jacobi) — SelfAdjointEigenSolver<Matrix4d>::eigenvectors().col(0)
LVBA_HD void smallest_eigvec4(const double A[16], double x[4]) {
  double a[4][4], v[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = A[4 * i + j]; v[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rot = 0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0 || !(fabs(apq) > 1e-22 * (fabs(a[p][p]) + fabs(a[q][q])))) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 4; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 4; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
        ++rot;
      }
    if (!rot) break;
  }
  int m = 0;
  for (int i = 1; i < 4; ++i) if (a[i][i] < a[m][m]) m = i;
  for (int k = 0; k < 4; ++k) x[k] = v[k][m];
}

struct TriangulateF {         // one item per track: TriangulateTrackDLT (:52-111)
  const int64_t* obs_ptr; const int32_t* obs_cam; const float* obs_uv; int n_cams; const double* cams; double intr[8];
  double* Xw; double* mean; int32_t* count; uint8_t* ok;
  LVBA_HD void operator()(int64_t t) const {
    ok[t] = 0; count[t] = 0; mean[t] = 0.0; Xw[3 * t] = Xw[3 * t + 1] = Xw[3 * t + 2] = 0.0;
    const int64_t a = obs_ptr[t], b = obs_ptr[t + 1];
    if (b - a < 4) return;                                                   // selected_ids.size() < 4  (:63)
    double AtA[16];
    for (int i = 0; i < 16; ++i) AtA[i] = 0.0;
    int rows = 0;
    for (int64_t q = a; q < b; ++q) {
      const int c = obs_cam[q];
      if (c < 0 || c >= n_cams) continue;
      double x, y;
      if (!depth::undistort_pixel(intr, (double)obs_uv[2 * q], (double)obs_uv[2 * q + 1], &x, &y)) continue;
      const double* P = cams + 12 * (int64_t)c;                             // rows of [Rcw | tcw]: P.row(r) = (R[3r..3r+2], t[r])
      double ru[4], rv[4];
      for (int k = 0; k < 3; ++k) { ru[k] = x * P[6 + k] - P[k]; rv[k] = y * P[6 + k] - P[3 + k]; }
      ru[3] = x * P[11] - P[9]; rv[3] = y * P[11] - P[10];
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) AtA[4 * i + j] += ru[i] * ru[j];     // :92-93
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) AtA[4 * i + j] += rv[i] * rv[j];
      rows += 2;
    }
    if (rows < 8) return;                                                    // :97
    double Xh[4];
    smallest_eigvec4(AtA, Xh);
    if (fabs(Xh[3]) < 1e-12) return;                                         // :103
    double X[3] = {Xh[0] / Xh[3], Xh[1] / Xh[3], Xh[2] / Xh[3]};
    if (!(finite_(X[0]) && finite_(X[1]) && finite_(X[2]))) return;
    Xw[3 * t] = X[0]; Xw[3 * t + 1] = X[1]; Xw[3 * t + 2] = X[2];
    double m; int32_t cnt;
    const bool good = mean_reproj(obs_cam, obs_uv, a, b, n_cams, cams, intr, X, 4, &m, &cnt);       // :108-110
    count[t] = cnt;
    if (good) { mean[t] = m; ok[t] = 1; }
  }
};

struct MeanReprojF {          // one item per track: ComputeMeanReproj of a given 3-D point
  const int64_t* obs_ptr; const int32_t* obs_cam; const float* obs_uv; int n_cams; const double* cams; double intr[8];
  const double* Xw; int min_count; double* mean; int32_t* count; uint8_t* ok;
  LVBA_HD void operator()(int64_t t) const {
    double m = 0.0; int32_t cnt = 0;
    const bool good = mean_reproj(obs_cam, obs_uv, obs_ptr[t], obs_ptr[t + 1], n_cams, cams, intr, Xw + 3 * t, min_count, &m, &cnt);
    mean[t] = good ? m : 0.0; count[t] = cnt; ok[t] = good ? 1 : 0;
  }
};

}  // namespace track
}  // namespace lvba
