// oracle/_ref/libbalm_ref.so — the REFERENCE'S OWN SOURCE FILES for the hot path, compiled where they lie.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Built only where /root/reference exists (this container), by oracle/Makefile
// (`make ref`), output into oracle/_ref/ (git-ignored).  Nothing under global-lvba_b200/ links, loads or calls it; only
// tests/ (to pin the oracles, tests/test_ref_pin.py) and tests/golden/make_golden_ref.py (to write the fixtures that
// travel to the GPU box) load it.
//
// What is compiled, unmodified, straight from /root/reference/include (no reference source is copied into this repo):
//   BALM/tools.hpp     Exp / hat / IMUST / PointCluster (+ transform) / pl_transform / down_sampling_voxel2
//   BALM/bavoxel.hpp   VOX_HESS (push_voxel, acc_evaluate2, evaluate_only_residual), OCTO_TREE_NODE (findCorrespondPoint,
//                      judge_eigen, cut_func, recut, tras_opt), BALM2 (divide_thread, only_residual, damping_iter), cut_voxel
//   utils.hpp          ReprojErrorWhitenedDistorted, PointPlaneErrorWhitened, distortNormalized, projectCameraToPixel,
//                      undistortPixelToNormalized, backProjectPixelDepthDistorted, fetchDepthBilinear, camToWorld
// What is NOT the reference: the libraries underneath.  Eigen, PCL, OpenCV, Ceres and Sophus are not installed and there is
// no network, so the headers those files include resolve to the stand-ins under oracle/ref_shim/ (own code, each file says what it
// is).  The consequences for what a fixture made here pins are listed in ref_shim/mini_eigen.h and DESIGN.md §2: every line of the
// reference's own arithmetic and control flow runs as written; the 3x3 eigen-solver, the sparse LDL^T and the dual numbers under it
// are ours.  src/lvba_system.cpp (ROS node, 2176 lines) is not compiled: its call sequences are mirrored below with the line
// cited at each step, and those few glue lines are the only restated code in this file.
//
// This file contains no algorithm: it moves flat arrays in and out of the reference's types and calls the reference's functions.
#include "BALM/bavoxel.hpp"
#include "utils.hpp"

#include <cstdint>
#include <map>

namespace {

std::vector<IMUST> poses_in(int W, const double* poses) {            // (W, 12): R row-major, then p
  std::vector<IMUST> xs(W);
  for (int i = 0; i < W; ++i) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) xs[i].R(r, c) = poses[12 * i + 3 * r + c];
    for (int r = 0; r < 3; ++r) xs[i].p(r) = poses[12 * i + 9 + r];
  }
  return xs;
}
void poses_out(const std::vector<IMUST>& xs, double* poses) {
  for (size_t i = 0; i < xs.size(); ++i) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) poses[12 * i + 3 * r + c] = xs[i].R(r, c);
    for (int r = 0; r < 3; ++r) poses[12 * i + 9 + r] = xs[i].p(r);
  }
}

// CSR voxels (oracle/lidar_oracle.py storage) -> the reference's dense vector<PointCluster>(win_size) per voxel
struct CsrVoxels {
  std::vector<std::vector<PointCluster>> sig;
  std::vector<PLV(3)> no_points;                  // push_voxel's second argument is never read (bavoxel.hpp:45-54)
  CsrVoxels(int W, int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx, const double* c)
      : sig((size_t)V, std::vector<PointCluster>(W)) {
    for (int64_t a = 0; a < V; ++a)
      for (int64_t s = vox_ptr[a]; s < vox_ptr[a + 1]; ++s) {
        PointCluster& pc = sig[a][pose_idx[s]];
        const double* q = c + 10 * s;            // Pxx Pxy Pxz Pyy Pyz Pzz vx vy vz N
        pc.P << q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5];
        pc.v << q[6], q[7], q[8];
        pc.N = (int)q[9];
      }
  }
  void push_all(VOX_HESS& vh) { for (auto& s : sig) vh.push_voxel(&s, &no_points); }
};

struct RefMap {
  int W = 0;
  std::unordered_map<VOXEL_LOC, OCTO_TREE_ROOT*> surf_map;
  std::unique_ptr<VOX_HESS> voxhess;
  std::vector<IMUST> poses;
  struct Meta { VOXEL_LOC key; int path[4]; OCTO_TREE_NODE* node; };
  std::map<const std::vector<PointCluster>*, Meta> by_sig;
  ~RefMap() { for (auto& kv : surf_map) delete kv.second; }
  void index(OCTO_TREE_NODE* n, const VOXEL_LOC& key, int depth, int* path) {      // plain traversal, no decisions
    Meta m; m.key = key; m.node = n;
    for (int i = 0; i < 4; ++i) m.path[i] = i < depth ? path[i] : -1;
    by_sig[&n->sig_orig] = m;
    for (int i = 0; i < 8; ++i)
      if (n->leaves[i] != nullptr && depth < 4) { path[depth] = i; index(n->leaves[i], key, depth + 1, path); }
  }
};

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ path A (B1)
// mode 0: ONE call of VOX_HESS::acc_evaluate2 over all kept voxels; residual = sum of lambda_0.
// mode 1: BALM2::divide_thread (thd_num = 16 threads, private Hessians summed); residual = sum / kept (AVG_THR).
// H: (6W)^2 row-major (symmetric), g: 6W.  Returns the number of voxels push_voxel kept.
int64_t ref_lidar_hessian(int W, int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx, const double* clusters,
                          const double* poses, int mode, double* H, double* g, double* residual) {
  CsrVoxels vox(W, V, vox_ptr, pose_idx, clusters);
  VOX_HESS vh(W);
  vox.push_all(vh);
  std::vector<IMUST> xs = poses_in(W, poses);
  const int n = 6 * W;
  Eigen::MatrixXd Hess(n, n);
  Eigen::VectorXd JacT(n);
  if (mode == 0) {
    vh.acc_evaluate2(xs, 0, (int)vh.plvec_voxels.size(), Hess, JacT, *residual);
  } else {
    BALM2 opt(W);
    std::vector<IMUST> x_ab(W);
    *residual = opt.divide_thread(xs, vh, x_ab, Hess, JacT);
  }
  for (int i = 0; i < n; ++i) { g[i] = JacT(i); for (int j = 0; j < n; ++j) H[(size_t)i * n + j] = Hess(i, j); }
  return (int64_t)vh.plvec_voxels.size();
}

// VOX_HESS::evaluate_only_residual over the kept voxels: sum of lambda_0 (not averaged).
double ref_lidar_residual(int W, int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx, const double* clusters,
                          const double* poses) {
  CsrVoxels vox(W, V, vox_ptr, pose_idx, clusters);
  VOX_HESS vh(W);
  vox.push_all(vh);
  std::vector<IMUST> xs = poses_in(W, poses);
  double r = 0;
  vh.evaluate_only_residual(xs, r);
  return r;
}

// BALM2::damping_iter (u0 = 0.01, v0 = 2, <= 10 passes, AVG_THR stop test): poses in, optimised poses out.
int64_t ref_lidar_damping_iter(int W, int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx, const double* clusters,
                               double* poses) {
  CsrVoxels vox(W, V, vox_ptr, pose_idx, clusters);
  VOX_HESS vh(W);
  vox.push_all(vh);
  std::vector<IMUST> xs = poses_in(W, poses);
  BALM2 opt(W);
  opt.damping_iter(xs, vh);
  poses_out(xs, poses);
  return (int64_t)vh.plvec_voxels.size();
}

// Exp (tools.hpp:62-77) on rows of w.
void ref_so3_exp(int64_t n, const double* w, double* R) {
  for (int64_t i = 0; i < n; ++i) {
    const Eigen::Matrix3d E = Exp(Eigen::Vector3d(w[3 * i], w[3 * i + 1], w[3 * i + 2]));
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[9 * i + 3 * r + c] = E(r, c);
  }
}

// ------------------------------------------------------------------------------------------------ set-up (B3)
// The sequence of runWindowBA src/lvba_system.cpp:247-258 (= runLidarBA :361-378): cut_voxel per scan, then recut + tras_opt per
// root.  eigen_ratio4 goes through set_eigen_ratio_array (:358); its first entry is cut_voxel's (unused) last argument.
void* ref_map_create(int W, const int64_t* scan_ptr, const float* xyz, const double* poses, double voxel_size,
                     const float* eigen_ratio4) {
  RefMap* m = new RefMap();
  m->W = W;
  m->poses = poses_in(W, poses);
  set_eigen_ratio_array({eigen_ratio4[0], eigen_ratio4[1], eigen_ratio4[2], eigen_ratio4[3]});
  for (int j = 0; j < W; ++j) {
    pcl::PointCloud<PointType> pl;
    pl.reserve((size_t)(scan_ptr[j + 1] - scan_ptr[j]));
    for (int64_t k = scan_ptr[j]; k < scan_ptr[j + 1]; ++k) {
      PointType p;
      p.x = xyz[3 * k]; p.y = xyz[3 * k + 1]; p.z = xyz[3 * k + 2];
      pl.push_back(p);
    }
    cut_voxel(m->surf_map, pl, m->poses[j], j, W, voxel_size, eigen_ratio4[0]);
  }
  m->voxhess.reset(new VOX_HESS(W));
  for (auto iter = m->surf_map.begin(); iter != m->surf_map.end(); ++iter) {
    iter->second->recut(m->poses);
    iter->second->tras_opt(*m->voxhess);
  }
  int path[4];
  for (auto& kv : m->surf_map) m->index(kv.second, kv.first, 0, path);
  return m;
}
void ref_map_destroy(void* h) { delete static_cast<RefMap*>(h); }
int64_t ref_map_num_voxels(void* h) { return (int64_t)static_cast<RefMap*>(h)->voxhess->plvec_voxels.size(); }
int64_t ref_map_num_slots(void* h) {
  RefMap* m = static_cast<RefMap*>(h);
  int64_t s = 0;
  for (auto* sig : m->voxhess->plvec_voxels) for (auto& pc : *sig) s += pc.N != 0;
  return s;
}
// The voxels VOX_HESS holds, in its own (unordered_map) order: key (V,3), path (V,4; -1 beyond the node's layer), layer, CSR
// clusters, and judge_eigen's centre / direct / eigenvalues of each plane node.
void ref_map_export(void* h, int64_t* key, int32_t* path, int32_t* layer, int64_t* vox_ptr, int32_t* pose_idx,
                    double* clusters, double* centre, double* direct, double* eigenvalues) {
  RefMap* m = static_cast<RefMap*>(h);
  int64_t s = 0, a = 0;
  vox_ptr[0] = 0;
  for (auto* sig : m->voxhess->plvec_voxels) {
    const RefMap::Meta& mt = m->by_sig.at(sig);
    key[3 * a] = mt.key.x; key[3 * a + 1] = mt.key.y; key[3 * a + 2] = mt.key.z;
    for (int i = 0; i < 4; ++i) path[4 * a + i] = mt.path[i];
    layer[a] = mt.node->layer;
    for (int i = 0; i < 3; ++i) { centre[3 * a + i] = mt.node->center(i); direct[3 * a + i] = mt.node->direct(i); eigenvalues[3 * a + i] = mt.node->value_vector(i); }
    for (int i = 0; i < m->W; ++i) {
      const PointCluster& pc = (*sig)[i];
      if (pc.N == 0) continue;
      double* q = clusters + 10 * s;
      q[0] = pc.P(0, 0); q[1] = pc.P(0, 1); q[2] = pc.P(0, 2); q[3] = pc.P(1, 1); q[4] = pc.P(1, 2); q[5] = pc.P(2, 2);
      q[6] = pc.v(0); q[7] = pc.v(1); q[8] = pc.v(2); q[9] = (double)pc.N;
      pose_idx[s++] = i;
    }
    vox_ptr[++a] = s;
  }
}
// damping_iter on the map's own voxels and poses (runWindowBA :264 without the skip rule of :259-263, which the caller applies).
void ref_map_damping_iter(void* h, double* poses_out_) {
  RefMap* m = static_cast<RefMap*>(h);
  std::vector<IMUST> xs = m->poses;
  BALM2 opt(m->W);
  opt.damping_iter(xs, *m->voxhess);
  poses_out(xs, poses_out_);
}
// The node a world point falls in: root key as recompute_local_planes computes it (src/lvba_system.cpp:1537-1543, restated: three
// lines of float arithmetic inside a lambda of the ROS node), then OCTO_TREE_NODE::findCorrespondPoint (reference code).
// state: -1 no root voxel, else the node's octo_state (0 UNKNOWN, 1 MID_NODE, 2 PLANE); direct / centre of the node.
void ref_map_lookup(void* h, int64_t n, const double* X, double voxel_size, int32_t* state, double* direct, double* centre) {
  RefMap* m = static_cast<RefMap*>(h);
  for (int64_t pi = 0; pi < n; ++pi) {
    Eigen::Vector3d x(X[3 * pi], X[3 * pi + 1], X[3 * pi + 2]);
    state[pi] = -1;
    for (int j = 0; j < 3; ++j) direct[3 * pi + j] = centre[3 * pi + j] = 0.0;
    if (!x.allFinite()) continue;
    float loc_xyz[3];
    for (int j = 0; j < 3; ++j) {
      loc_xyz[j] = x[j] / voxel_size;
      if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0f;
    }
    VOXEL_LOC key((int64_t)loc_xyz[0], (int64_t)loc_xyz[1], (int64_t)loc_xyz[2]);
    auto it = m->surf_map.find(key);
    if (it == m->surf_map.end()) continue;
    OCTO_TREE_NODE* node = it->second->findCorrespondPoint(x);
    state[pi] = (int32_t)node->octo_state;
    if (node->octo_state != PLANE) continue;
    for (int j = 0; j < 3; ++j) { direct[3 * pi + j] = node->direct(j); centre[3 * pi + j] = node->center(j); }
  }
}

// ------------------------------------------------------------------------------------------------ anchor clouds (B6)
// Tail of the window loop, src/lvba_system.cpp:284-297: every scan of the window through pl_transform(tmp, rel), merged in scan
// order, then down_sampling_voxel2(merged, leaf).  rel: (W, 12).  out: capacity >= number of input points; returns the count.
// Output order is the unordered_map's.
int64_t ref_anchor_cloud(int W, const int64_t* scan_ptr, const float* xyz, const double* rel, double leaf, float* out) {
  std::vector<IMUST> r = poses_in(W, rel);
  pcl::PointCloud<PointType> merged;
  for (int j = 0; j < W; ++j) {
    pcl::PointCloud<PointType> tmp;
    for (int64_t k = scan_ptr[j]; k < scan_ptr[j + 1]; ++k) {
      PointType p;
      p.x = xyz[3 * k]; p.y = xyz[3 * k + 1]; p.z = xyz[3 * k + 2];
      tmp.push_back(p);
    }
    pl_transform(tmp, r[j]);
    for (auto& p : tmp.points) merged.push_back(p);                 // `*merged += tmp` (:292)
  }
  down_sampling_voxel2(merged, leaf);
  for (size_t i = 0; i < merged.size(); ++i) { out[3 * i] = merged[i].x; out[3 * i + 1] = merged[i].y; out[3 * i + 2] = merged[i].z; }
  return (int64_t)merged.size();
}

// ------------------------------------------------------------------------------------------------ path B functors (B2)
// ReprojErrorWhitenedDistorted (utils.hpp:51-125) as ceres::AutoDiffCostFunction<..., 2, 4, 3, 3> evaluates it: residuals with
// T = double, Jacobians with one Jet partial per parameter coordinate.  J: (n, 2, 10) = d r / d [q(4, ambient) | t(3) | X(3)].
void ref_reproj(int64_t n, const double* q, const double* t, const double* X, const double* uv, const double* intr,
                double su, double sv, double* r, double* J) {
  typedef ceres::Jet<double, 10> JetT;
  for (int64_t i = 0; i < n; ++i) {
    lvba::ReprojErrorWhitenedDistorted f(uv[2 * i], uv[2 * i + 1], intr[0], intr[1], intr[2], intr[3], intr[4], intr[5],
                                         intr[6], intr[7], su, sv);
    f(q + 4 * i, t + 3 * i, X + 3 * i, r + 2 * i);
    if (J == nullptr) continue;
    JetT jq[4], jt[3], jx[3], jr[2];
    for (int k = 0; k < 4; ++k) jq[k] = JetT(q[4 * i + k], k);
    for (int k = 0; k < 3; ++k) jt[k] = JetT(t[3 * i + k], 4 + k);
    for (int k = 0; k < 3; ++k) jx[k] = JetT(X[3 * i + k], 7 + k);
    f(jq, jt, jx, jr);
    for (int a = 0; a < 2; ++a) for (int k = 0; k < 10; ++k) J[20 * i + 10 * a + k] = jr[a].v[k];
  }
}
// PointPlaneErrorWhitened (utils.hpp:129-147) as AutoDiffCostFunction<..., 1, 3>.  nd: (n, 4) = normal, d.  J: (n, 3).
void ref_point_plane(int64_t n, const double* X, const double* nd, double sigma, double* r, double* J) {
  typedef ceres::Jet<double, 3> JetT;
  for (int64_t i = 0; i < n; ++i) {
    lvba::PointPlaneErrorWhitened f(Eigen::Vector3d(nd[4 * i], nd[4 * i + 1], nd[4 * i + 2]), nd[4 * i + 3], sigma);
    f(X + 3 * i, r + i);
    if (J == nullptr) continue;
    JetT jx[3], jr[1];
    for (int k = 0; k < 3; ++k) jx[k] = JetT(X[3 * i + k], k);
    f(jx, jr);
    for (int k = 0; k < 3; ++k) J[3 * i + k] = jr[0].v[k];
  }
}

// ------------------------------------------------------------------------------------------------ camera helpers (B4 / B7)
static lvba::CameraIntrinsics intr_in(const double* a) {
  lvba::CameraIntrinsics c;
  c.fx = a[0]; c.fy = a[1]; c.cx = a[2]; c.cy = a[3]; c.k1 = a[4]; c.k2 = a[5]; c.p1 = a[6]; c.p2 = a[7];
  return c;
}
void ref_project_camera_to_pixel(int64_t n, const double* intr, const double* Xc, double* uv, double* z, uint8_t* ok) {
  const lvba::CameraIntrinsics cam = intr_in(intr);
  for (int64_t i = 0; i < n; ++i) {
    double u = 0, v = 0, zc = 0;
    ok[i] = lvba::projectCameraToPixel(cam, Eigen::Vector3d(Xc[3 * i], Xc[3 * i + 1], Xc[3 * i + 2]), &u, &v, &zc);
    uv[2 * i] = u; uv[2 * i + 1] = v; z[i] = zc;
  }
}
void ref_undistort_pixel(int64_t n, const double* intr, const double* uv, double* xy, uint8_t* ok) {
  const lvba::CameraIntrinsics cam = intr_in(intr);
  for (int64_t i = 0; i < n; ++i) {
    double x = 0, y = 0;
    ok[i] = lvba::undistortPixelToNormalized(cam, uv[2 * i], uv[2 * i + 1], &x, &y);
    xy[2 * i] = x; xy[2 * i + 1] = y;
  }
}
// fetchDepthBilinear (CV_32FC1) -> backProjectPixelDepthDistorted -> camToWorld, the chain of BuildTracksAndFuse3D's depth candidate
// (src/lvba_system.cpp:1023-1040); cam: R_cw row-major (9), t_cw (3).  ok bit 0: depth fetched, bit 1: back-projected.
void ref_depth_candidate(int h, int w, const float* depth, const double* intr, const double* cam, int64_t n, const float* uv,
                         float* d_out, double* Xw, uint8_t* ok) {
  cv::Mat img(h, w, CV_32FC1);
  for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) img.at<float>(y, x) = depth[(size_t)y * w + x];
  const lvba::CameraIntrinsics ci = intr_in(intr);
  Eigen::Matrix3d Rcw; Eigen::Vector3d tcw;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Rcw(r, c) = cam[3 * r + c]; tcw(r) = cam[9 + r]; }
  for (int64_t i = 0; i < n; ++i) {
    float d = 0.f;
    ok[i] = 0; d_out[i] = 0.f;
    for (int j = 0; j < 3; ++j) Xw[3 * i + j] = 0.0;
    if (!lvba::fetchDepthBilinear(img, uv[2 * i], uv[2 * i + 1], d)) continue;
    ok[i] = 1; d_out[i] = d;
    Eigen::Vector3d Xc;
    if (!lvba::backProjectPixelDepthDistorted(ci, uv[2 * i], uv[2 * i + 1], d, &Xc)) continue;
    ok[i] = 3;
    const Eigen::Vector3d X = lvba::camToWorld(Xc, Rcw, tcw);
    for (int j = 0; j < 3; ++j) Xw[3 * i + j] = X(j);
  }
}

}  // extern "C"
