// lvba_shim.hpp — host-side C++ mirror of the two reference call sites, written against the C ABI of
// include/lvba_b200.h.  Header only, C++17, no Eigen / Ceres / PCL / ROS includes of its own: the functions
// are templates over the reference's own types (IMUST, PointCluster, VOX_HESS, Eigen matrices are only
// touched through operator()(i,j) / operator()(i) / public members), so the same header compiles inside
// src/lvba_system.cpp of the reference and inside tests/shim/test_shim.cpp with plain mock structs.
//
//   lvba_b200::damping_iter(x_stats, voxhess)        replaces BALM2::damping_iter   include/BALM/bavoxel.hpp:662-767
//   lvba_b200::solve_visual(...)                     replaces the Ceres block       src/lvba_system.cpp:1571-1656
//   lvba_b200::WindowBatch                           collects the windows of runWindowBA (src/lvba_system.cpp:232-302) and
//                                                    solves them in one lvba_lidar_lm_batch call
//
// Same names, argument meaning and error behaviour as the reference: void-like use (the reference ignores
// solver failure), state written back only on success, size mismatches throw std::runtime_error like
// src/lvba_system.cpp:1427-1432.  See INTEGRATION.md for the three-line patches.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "lvba_b200.h"

namespace lvba_b200 {

// ---- B1: x_stats[i].R (3x3), x_stats[i].p (3) ; voxhess.plvec_voxels[a] -> const vector<PointCluster>* with
//      (*ptr)[i].P (3x3), .v (3), .N ; voxhess.win_size
template <class PoseVec, class VoxHess>
inline int damping_iter(PoseVec& x_stats, VoxHess& voxhess, const lvba_lidar_opts* opts = nullptr,
                        lvba_summary* summary = nullptr) {
  const int W = voxhess.win_size;
  if ((int)x_stats.size() < W) throw std::runtime_error("lvba_b200::damping_iter: x_stats smaller than win_size");
  const int64_t V = (int64_t)voxhess.plvec_voxels.size();
  std::vector<int64_t> vox_ptr(V + 1, 0);
  std::vector<int32_t> pose_idx;
  std::vector<double> clusters;
  for (int64_t a = 0; a < V; ++a) {                    // pack slots with N != 0 (bavoxel.hpp:91, :113)
    const auto& sig = *voxhess.plvec_voxels[a];
    for (int i = 0; i < W; ++i) {
      if (sig[i].N == 0) continue;
      pose_idx.push_back(i);
      const auto& P = sig[i].P; const auto& v = sig[i].v;
      const double rec[10] = {P(0, 0), P(0, 1), P(0, 2), P(1, 1), P(1, 2), P(2, 2), v(0), v(1), v(2), (double)sig[i].N};
      clusters.insert(clusters.end(), rec, rec + 10);
    }
    vox_ptr[a + 1] = (int64_t)pose_idx.size();
  }
  std::vector<double> poses((size_t)W * 12);
  for (int i = 0; i < W; ++i) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) poses[12 * i + 3 * r + c] = x_stats[i].R(r, c);
    for (int r = 0; r < 3; ++r) poses[12 * i + 9 + r] = x_stats[i].p(r);
  }
  const int rc = lvba_lidar_lm(W, V, vox_ptr.data(), pose_idx.data(), clusters.data(), poses.data(), opts, summary);
  if (rc != LVBA_OK) return rc;                        // reference: silent on failure, state untouched
  for (int i = 0; i < W; ++i) {                        // x_stats = x_stats_temp (bavoxel.hpp:746)
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) x_stats[i].R(r, c) = poses[12 * i + 3 * r + c];
    for (int r = 0; r < 3; ++r) x_stats[i].p(r) = poses[12 * i + 9 + r];
  }
  return LVBA_OK;
}

// ---- B1 batched: runWindowBA builds one (x_win, voxhess) pair per window and calls damping_iter on each
//      (src/lvba_system.cpp:239-264).  WindowBatch::add() packs a window instead of solving it; solve() runs every
//      packed window in one call and writes the poses back into the x_win vectors it was given (which must outlive
//      solve()).  Windows below the reference's `plvec_voxels.size() < 3 * x_win.size()` rule (:262-266) are packed too
//      and come back untouched with LVBA_TERM_SKIPPED, so the caller's bookkeeping (win_skipped) can read the summary.
template <class PoseVec>
class WindowBatch {
 public:
  template <class VoxHess>
  void add(PoseVec& x_win, const VoxHess& voxhess) {
    const int W = voxhess.win_size;
    if ((int)x_win.size() < W) throw std::runtime_error("lvba_b200::WindowBatch::add: x_win smaller than win_size");
    const int32_t base = win_ptr_.back();
    for (const auto* sigp : voxhess.plvec_voxels) {
      const auto& sig = *sigp;
      for (int i = 0; i < W; ++i) {
        if (sig[i].N == 0) continue;
        pose_idx_.push_back(base + i);
        const auto& P = sig[i].P; const auto& v = sig[i].v;
        const double rec[10] = {P(0, 0), P(0, 1), P(0, 2), P(1, 1), P(1, 2), P(2, 2), v(0), v(1), v(2), (double)sig[i].N};
        clusters_.insert(clusters_.end(), rec, rec + 10);
      }
      vox_ptr_.push_back((int64_t)pose_idx_.size());
    }
    for (int i = 0; i < W; ++i) {
      double rec[12];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) rec[3 * r + c] = x_win[i].R(r, c);
      for (int r = 0; r < 3; ++r) rec[9 + r] = x_win[i].p(r);
      poses_.insert(poses_.end(), rec, rec + 12);
    }
    win_ptr_.push_back(base + W);
    targets_.push_back(&x_win);
  }
  int size() const { return (int)targets_.size(); }
  // returns LVBA_OK or a negative lvba_status; on failure no window is modified (reference: silent, state untouched)
  int solve(int min_voxels_per_pose = 3, const lvba_lidar_opts* opts = nullptr, std::vector<lvba_summary>* summaries = nullptr,
            lvba_summary* total = nullptr) {
    if (targets_.empty()) return LVBA_OK;
    std::vector<lvba_summary> local((size_t)size());
    const int rc = lvba_lidar_lm_batch(size(), win_ptr_.data(), (int64_t)vox_ptr_.size() - 1, vox_ptr_.data(), pose_idx_.data(),
                                       clusters_.data(), poses_.data(), min_voxels_per_pose, opts, local.data(), total);
    if (rc != LVBA_OK) return rc;
    for (int w = 0; w < size(); ++w) {
      PoseVec& x = *targets_[w];
      for (int i = 0; i < win_ptr_[w + 1] - win_ptr_[w]; ++i) {
        const double* rec = poses_.data() + 12 * (size_t)(win_ptr_[w] + i);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) x[i].R(r, c) = rec[3 * r + c];
        for (int r = 0; r < 3; ++r) x[i].p(r) = rec[9 + r];
      }
    }
    if (summaries) *summaries = local;
    return LVBA_OK;
  }

 private:
  std::vector<int32_t> win_ptr_{0};
  std::vector<int64_t> vox_ptr_{0};
  std::vector<int32_t> pose_idx_;
  std::vector<double> clusters_, poses_;
  std::vector<PoseVec*> targets_;
};

// ---- B2: the flat arrays optimizeCameraPoses already builds (qs, ts, Xs, plane_n, plane_d) plus the
//      observation list it walks at :1614-1631.  obs_of_point[pi] = list of (cam_id, u, v).
struct Observation { int cam; float u, v; };

inline int solve_visual(std::vector<std::array<double, 4>>& qs, std::vector<std::array<double, 3>>& ts,
                        std::vector<std::array<double, 3>>& Xs, const std::vector<std::array<double, 3>>& plane_n,
                        const std::vector<double>& plane_d, const std::vector<std::vector<Observation>>& obs_of_point,
                        double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2,
                        double sigma_px = 0.5, double sigma_plane = 0.01, const lvba_visual_opts* opts = nullptr,
                        lvba_summary* summary = nullptr) {
  const int M = (int)qs.size();
  const int64_t T = (int64_t)Xs.size();
  if ((int)ts.size() != M) throw std::runtime_error("lvba_b200::solve_visual: qs/ts size mismatch");      // cf. :1427
  if ((int64_t)plane_n.size() != T || (int64_t)plane_d.size() != T || (int64_t)obs_of_point.size() != T)
    throw std::runtime_error("lvba_b200::solve_visual: per-point array size mismatch");
  std::vector<int64_t> obs_ptr(T + 1, 0);
  std::vector<int32_t> obs_cam;
  std::vector<float> obs_uv;
  std::vector<double> plane_nd((size_t)T * 4);
  for (int64_t i = 0; i < T; ++i) {
    for (const auto& o : obs_of_point[i]) { obs_cam.push_back(o.cam); obs_uv.push_back(o.u); obs_uv.push_back(o.v); }
    obs_ptr[i + 1] = (int64_t)obs_cam.size();
    plane_nd[4 * i] = plane_n[i][0]; plane_nd[4 * i + 1] = plane_n[i][1]; plane_nd[4 * i + 2] = plane_n[i][2]; plane_nd[4 * i + 3] = plane_d[i];
  }
  const double intr[8] = {fx, fy, cx, cy, k1, k2, p1, p2};
  // qs/ts/Xs are contiguous arrays of std::array<double,N>: exactly the M*4 / M*3 / T*3 layout of the ABI
  return lvba_visual_lm(M, T, qs.empty() ? nullptr : qs[0].data(), ts.empty() ? nullptr : ts[0].data(),
                        Xs.empty() ? nullptr : Xs[0].data(), plane_nd.data(), obs_ptr.data(), obs_cam.data(), obs_uv.data(), intr,
                        sigma_px, sigma_plane, /*fixed_cam=*/0, opts, summary);
}

}  // namespace lvba_b200
