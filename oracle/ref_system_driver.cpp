// oracle/_ref/liblvba_system_ref.so — THE REFERENCE'S PIPELINE SOURCE (src/lvba_system.cpp, all 2176 lines, unmodified, compiled
// where it lies) driven through its own public members.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ref_driver.cpp, which does the same for the BALM headers alone).  Built only where
// /root/reference exists, by oracle/Makefile (`make ref`).  Nothing under global-lvba_b200/ links, loads or calls it.
//
// The file below is included, not copied: `#include "<reference>/src/lvba_system.cpp"` puts LvbaSystem's member functions —
// runWindowBA, runLidarBA, buildGridMapFromOptimized, updateCameraPosesFromLidar, generateDepthWithVoxel, BuildTracksAndFuse3D
// (with ComputeMeanReproj / TriangulateTrackDLT), optimizeCameraPoses, loadFromColmapDB — into this translation unit together with
// the BALM / utils headers they call.  ROS, PCL, OpenCV, Ceres, Sophus, SiftGPU and Eigen resolve to the stand-ins under
// oracle/ref_shim/ (own code; NOT those libraries): publishers swallow their messages, image files are not written, and
// ceres::Problem RECORDS the problem optimizeCameraPoses builds and hands it to the hook below instead of solving it.  SQLite is the
// system's libsqlite3.so.0 (declarations in ref_shim/sqlite3.h).  src/dataset_io.cpp (the dataset loader) is included the same way; its PCD
// files go through the stand-in reader of ref_shim/pcl/io/pcd_io.h (NOT PCL's parser).  LvbaSystem's constructor runs the loader on the
// parameter table's data path; the driver's setters then overwrite the public fields for the tests that bring their own arrays.
//
// What a fixture made with this pins: the reference's own sequencing and arithmetic of the stages named above (skip rule, anchors,
// relative poses, the two global stages, depth splatting, the match graph / component walk / candidate choice of the track fusion,
// which tracks and planes enter the Ceres problem and with which constancy, manifold and loss).  What it does not: see
// ref_shim/mini_eigen.h (eigen-solver, LDL^T), and the Ceres solve itself — the hook lets the CALLER solve the recorded problem
// (the tests use the restated Ceres loop of oracle/visual_oracle.py) and writes the answer back into the reference's parameter blocks,
// after which the reference's own code carries on.  No algorithm lives in this file.
#include REF_SYSTEM_CPP
#ifndef ROOT_DIR
#define ROOT_DIR ""          // CMakeLists.txt:13 prefixes data_config/data_path with the package directory; the tests pass absolute paths
#endif
#include REF_DATASET_CPP

#include <cstdint>
#include <sstream>


namespace {

struct Sys {
  ros::NodeHandle nh;
  std::unique_ptr<lvba::LvbaSystem> s;
  // optimizeCameraPoses: the recorded problem, flattened
  std::vector<double*> cam_q, cam_t, pts;
  std::vector<uint8_t> cam_const;
  std::vector<int32_t> cam_manifold;                 // 0 none, 1 EigenQuaternionManifold, 2 QuaternionManifold
  struct Obs { int cam, pt; double u, v, intr[8], su, sv; int has_loss; };
  struct Pl { int pt; double nd[4], sigma; int has_loss; };
  std::vector<Obs> obs;
  std::vector<Pl> planes;
  double opt[8] = {0};
  void (*solve_cb)(void*) = nullptr;
  void* solve_user = nullptr;
  Sys() { s.reset(new lvba::LvbaSystem(nh)); }
};

IMUST pose_in(const double* p, double t) {
  IMUST x;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) x.R(r, c) = p[3 * r + c];
  for (int r = 0; r < 3; ++r) x.p(r) = p[9 + r];
  x.t = t;
  return x;
}
void pose_out(const Eigen::Matrix3d& R, const Eigen::Vector3d& p, double* o) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = R(r, c);
  for (int r = 0; r < 3; ++r) o[9 + r] = p(r);
}

void on_solve(Sys* y, const ceres::Solver::Options& o, ceres::Problem* pr, ceres::Solver::Summary* sm) {
  typedef ceres::AutoDiffCostFunction<lvba::ReprojErrorWhitenedDistorted, 2, 4, 3, 3> ReprojCost;
  typedef ceres::AutoDiffCostFunction<lvba::PointPlaneErrorWhitened, 1, 3> PlaneCost;
  y->cam_q.clear(); y->cam_t.clear(); y->pts.clear(); y->cam_const.clear(); y->cam_manifold.clear(); y->obs.clear(); y->planes.clear();
  std::map<double*, int> cam_of, pt_of;
  // optimizeCameraPoses adds (q_k, t_k) for every camera first (src/lvba_system.cpp:1578-1581), then one block per kept point
  size_t b = 0;
  while (b + 1 < pr->params.size() && pr->params[b].size == 4) {
    cam_of[pr->params[b].p] = cam_of[pr->params[b + 1].p] = (int)y->cam_q.size();
    y->cam_q.push_back(pr->params[b].p); y->cam_t.push_back(pr->params[b + 1].p);
    y->cam_const.push_back((pr->params[b].constant ? 1 : 0) | (pr->params[b + 1].constant ? 2 : 0));
    y->cam_manifold.push_back(dynamic_cast<ceres::EigenQuaternionManifold*>(pr->params[b].manifold) ? 1
                              : dynamic_cast<ceres::QuaternionManifold*>(pr->params[b].manifold) ? 2 : 0);
    b += 2;
  }
  for (; b < pr->params.size(); ++b) { pt_of[pr->params[b].p] = (int)y->pts.size(); y->pts.push_back(pr->params[b].p); }
  for (const auto& r : pr->residuals) {
    if (auto* c = dynamic_cast<ReprojCost*>(r.cost)) {
      const auto& f = c->functor();
      Sys::Obs ob{cam_of.at(r.params[0]), pt_of.at(r.params[2]), f.u_, f.v_, {f.fx_, f.fy_, f.cx_, f.cy_, f.k1_, f.k2_, f.p1_, f.p2_}, f.su_, f.sv_, r.loss != nullptr};
      if (cam_of.at(r.params[1]) != ob.cam) std::abort();
      y->obs.push_back(ob);
    } else if (auto* c2 = dynamic_cast<PlaneCost*>(r.cost)) {
      const auto& f = c2->functor();
      y->planes.push_back(Sys::Pl{pt_of.at(r.params[0]), {f.nx_, f.ny_, f.nz_, f.d_}, f.s_, r.loss != nullptr});
    } else {
      std::abort();
    }
  }
  y->opt[0] = o.max_num_iterations; y->opt[1] = (double)o.linear_solver_type; y->opt[2] = o.function_tolerance; y->opt[3] = o.gradient_tolerance;
  y->opt[4] = o.parameter_tolerance; y->opt[5] = o.num_threads;
  sm->termination_type = ceres::CONVERGENCE;
  if (y->solve_cb) y->solve_cb(y->solve_user);       // the caller solves the recorded problem and writes the answer back
}

}  // namespace

extern "C" {

void* sys_create() { return new Sys(); }
void sys_destroy(void* h) { delete static_cast<Sys*>(h); }
// The plane test of OCTO_TREE_NODE::judge_eigen reads a process-wide array (bavoxel.hpp:17-22); runLidarBA sets it per stage (:358) and
// optimizeCameraPoses inherits whatever the last stage left there.  Tests that run the camera half alone state it explicitly.
void sys_set_eigen_ratio_array(const float* r) { set_eigen_ratio_array({r[0], r[1], r[2], r[3]}); }
void sys_set_param(const char* name, double v) { ros::ParamValue p; p.num = v; ros::param_table()[name] = p; }
void sys_set_param_str(const char* name, const char* v) { ros::ParamValue p; p.str = v; ros::param_table()[name] = p; }
void sys_set_param_vec(const char* name, int n, const double* v) { ros::ParamValue p; p.vec.assign(v, v + n); ros::param_table()[name] = p; }
void sys_clear_params() { ros::param_table().clear(); }
// what DatasetIO's constructor loaded (src/dataset_io.cpp): sizes, then the arrays
void sys_dataset_sizes(void* h, int64_t* n_frames, int64_t* n_points, int64_t* n_images) {
  auto& d = *static_cast<Sys*>(h)->s->dataset_io_;
  *n_frames = (int64_t)d.x_buf_.size(); *n_images = (int64_t)d.images_ids_.size(); *n_points = 0;
  for (auto& pl : d.pl_fulls_) *n_points += (int64_t)pl->size();
}
void sys_dataset_get(void* h, double* frame_ts, double* frame_poses, int64_t* scan_ptr, float* xyz, float* intensity, double* image_ts, double* image_poses,
                     double* cam /* width height fx fy cx cy k1 k2 p1 p2 scale stride */, char* paths /* data path \n db path, 4096 */) {
  auto& d = *static_cast<Sys*>(h)->s->dataset_io_;
  for (size_t i = 0; i < d.x_buf_.size(); ++i) { frame_ts[i] = d.x_buf_[i].t; pose_out(d.x_buf_[i].R, d.x_buf_[i].p, frame_poses + 12 * i); }
  int64_t o = 0;
  scan_ptr[0] = 0;
  for (size_t i = 0; i < d.pl_fulls_.size(); ++i) {
    for (auto& p : d.pl_fulls_[i]->points) { xyz[3 * o] = p.x; xyz[3 * o + 1] = p.y; xyz[3 * o + 2] = p.z; intensity[o] = p.intensity; ++o; }
    scan_ptr[i + 1] = o;
  }
  for (size_t k = 0; k < d.images_ids_.size(); ++k) image_ts[k] = d.images_ids_[k];
  for (size_t k = 0; k < d.image_poses_.size(); ++k) pose_out(d.image_poses_[k].rotation_matrix(), d.image_poses_[k].translation(), image_poses + 12 * k);
  const double c[12] = {(double)d.width_, (double)d.height_, d.fx_, d.fy_, d.cx_, d.cy_, d.k1_, d.k2_, d.p1_, d.p2_, d.resize_scale_, (double)d.image_stride_};
  for (int j = 0; j < 12; ++j) cam[j] = c[j];
  std::snprintf(paths, 4096, "%s\n%s", d.dataset_path_.c_str(), d.colmap_db_path_.c_str());
}

// ---------------------------------------------------------------------------------------------- inputs (public fields of DatasetIO)
void sys_set_lidar(void* h, int W, const int64_t* scan_ptr, const float* xyz, const double* poses, const double* ts) {
  auto& d = *static_cast<Sys*>(h)->s->dataset_io_;
  d.x_buf_.clear(); d.pl_fulls_.clear();
  for (int j = 0; j < W; ++j) {
    d.x_buf_.push_back(pose_in(poses + 12 * j, ts ? ts[j] : (double)j));
    pcl::PointCloud<PointType>::Ptr pl(new pcl::PointCloud<PointType>());
    for (int64_t k = scan_ptr[j]; k < scan_ptr[j + 1]; ++k) { PointType p; p.x = xyz[3 * k]; p.y = xyz[3 * k + 1]; p.z = xyz[3 * k + 2]; pl->push_back(p); }
    d.pl_fulls_.push_back(pl);
  }
  d.x_buf_before_ = d.x_buf_;
}
// x_buf_ after the LiDAR half (runLidarBA leaves its result there, :405); x_buf_before_ keeps the odometry
void sys_set_lidar_optimised(void* h, const double* poses) {
  auto& d = *static_cast<Sys*>(h)->s->dataset_io_;
  for (size_t j = 0; j < d.x_buf_.size(); ++j) d.x_buf_[j] = pose_in(poses + 12 * j, d.x_buf_[j].t);
}
void sys_set_stages(void* h, int window_enable, int window_size, double anchor_leaf, int use_rel, int stage1_enable, double s1_voxel,
                    const float* s1_ratio, double s2_voxel, const float* s2_ratio) {
  auto& d = *static_cast<Sys*>(h)->s->dataset_io_;
  d.window_ba_enable_ = window_enable; d.window_ba_size_ = window_size; d.anchor_leaf_size_ = anchor_leaf; d.use_window_ba_rel_ = use_rel;
  d.stage1_enable_ = stage1_enable; d.stage1_root_voxel_size_ = s1_voxel; d.stage2_root_voxel_size_ = s2_voxel;
  d.stage1_eigen_ratio_array_.assign(s1_ratio, s1_ratio + 4); d.stage2_eigen_ratio_array_.assign(s2_ratio, s2_ratio + 4);
}
// camera side: intrinsics, LiDAR->camera and LiDAR->IMU extrinsics as the YAML holds them, image timestamps and odometry image poses
void sys_set_camera(void* h, int width, int height, const double* intr, const double* Rcl, const double* tcl, const double* Ril, const double* til,
                    int M, const double* image_ts, const double* image_poses) {
  auto& d = *static_cast<Sys*>(h)->s->dataset_io_;
  d.width_ = width; d.height_ = height; d.resize_scale_ = 1.0;
  d.fx_ = intr[0]; d.fy_ = intr[1]; d.cx_ = intr[2]; d.cy_ = intr[3]; d.k1_ = intr[4]; d.k2_ = intr[5]; d.p1_ = intr[6]; d.p2_ = intr[7];
  d.cameraextrinR_.assign(Rcl, Rcl + 9); d.cameraextrinT_.assign(tcl, tcl + 3); d.extrinR_.assign(Ril, Ril + 9); d.extrinT_.assign(til, til + 3);
  d.images_ids_.assign(image_ts, image_ts + M);
  d.image_poses_.clear();
  for (int k = 0; k < M; ++k) { IMUST x = pose_in(image_poses + 12 * k, 0); d.image_poses_.push_back(Sophus::SE3(x.R, x.p)); }
  d.dataset_path_ = "/tmp/lvba_ref_unused/";
}

// ---------------------------------------------------------------------------------------------- the LiDAR half
// runWindowBA (src/lvba_system.cpp:204-302) alone.  Returns the number of anchors; cloud_ptr (anchors + 1), clouds: capacity = all points.
int sys_run_window_ba(void* h, double* anchor_poses, int64_t* cloud_ptr, float* clouds, double* rel_poses, int32_t* anchor_index) {
  auto& s = *static_cast<Sys*>(h)->s;
  std::vector<IMUST> x = s.dataset_io_->x_buf_;
  const int total = (int)x.size();
  s.rel_poses_to_anchor_.assign(total, IMUST());          // as runLidarBA does before the call (:330-331)
  s.anchor_index_per_frame_.assign(total, -1);
  std::vector<IMUST> ap;
  std::vector<pcl::PointCloud<PointType>::Ptr> ac;
  s.runWindowBA(x, s.dataset_io_->pl_fulls_, ap, ac);
  cloud_ptr[0] = 0;
  for (size_t a = 0; a < ap.size(); ++a) {
    pose_out(ap[a].R, ap[a].p, anchor_poses + 12 * a);
    int64_t o = cloud_ptr[a];
    for (auto& p : ac[a]->points) { clouds[3 * o] = p.x; clouds[3 * o + 1] = p.y; clouds[3 * o + 2] = p.z; ++o; }
    cloud_ptr[a + 1] = o;
  }
  for (int i = 0; i < total; ++i) { pose_out(s.rel_poses_to_anchor_[i].R, s.rel_poses_to_anchor_[i].p, rel_poses + 12 * i); anchor_index[i] = s.anchor_index_per_frame_[i]; }
  return (int)ap.size();
}
// runLidarBA (:304-409): window stage, the two global stages on the anchors, poses of every frame from its anchor.  The function asks
// the terminal for a '1' before it starts (:323-328); the driver answers.  out: (W, 12) = dataset_io_->x_buf_ afterwards.
void sys_run_lidar_ba(void* h, double* out) {
  auto& s = *static_cast<Sys*>(h)->s;
  std::istringstream yes("1\n");
  std::streambuf* old = std::cin.rdbuf(yes.rdbuf());
  s.runLidarBA();
  std::cin.rdbuf(old);
  const auto& x = s.dataset_io_->x_buf_;
  for (size_t i = 0; i < x.size(); ++i) pose_out(x[i].R, x[i].p, out + 12 * i);
}

// ---------------------------------------------------------------------------------------------- the camera half, stage by stage
void sys_init_from_dataset(void* h) { static_cast<Sys*>(h)->s->initFromDatasetIO(); }
void sys_build_grid(void* h) { static_cast<Sys*>(h)->s->buildGridMapFromOptimized(); }
void sys_update_camera_poses(void* h, double* poses /* (M, 12) body poses of the images */) {
  auto& s = *static_cast<Sys*>(h)->s;
  s.updateCameraPosesFromLidar();
  for (size_t k = 0; k < s.poses_.size(); ++k) pose_out(s.poses_[k].rotation_matrix(), s.poses_[k].translation(), poses + 12 * k);
}
// generateDepthWithVoxel (:835-919): depth (M, height, width) float32; cams / cams_opt (M, 12) = (Rcw, tcw) from the odometry / updated poses
void sys_generate_depth(void* h, float* depth, double* cams, double* cams_opt) {
  auto& s = *static_cast<Sys*>(h)->s;
  s.generateDepthWithVoxel();
  const size_t hw = (size_t)s.image_height_ * s.image_width_;
  for (size_t k = 0; k < s.all_depths_.size(); ++k) {
    for (int y = 0; y < s.image_height_; ++y) for (int x = 0; x < s.image_width_; ++x) depth[k * hw + (size_t)y * s.image_width_ + x] = s.all_depths_[k].at<float>(y, x);
    pose_out(s.Rcw_all_[k], s.tcw_all_[k], cams + 12 * k);
    pose_out(s.Rcw_all_optimized_[k], s.tcw_all_optimized_[k], cams_opt + 12 * k);
  }
}
// direct set-up of what the fusion reads, for scenes that do not come from a LiDAR run: cameras (Rcw, tcw), depth images, keypoints, matches
void sys_set_fusion_inputs(void* h, int M, int width, int height, const double* intr, const double* cams, const float* depth,
                           const int64_t* kp_ptr, const float* kp_uv, int64_t n_matches, const int32_t* matches /* (n, 4): img_a kp_a img_b kp_b */,
                           int obser_thr) {
  auto& s = *static_cast<Sys*>(h)->s;
  s.image_width_ = width; s.image_height_ = height;
  s.fx_ = intr[0]; s.fy_ = intr[1]; s.cx_ = intr[2]; s.cy_ = intr[3]; s.d0_ = intr[4]; s.d1_ = intr[5]; s.d2_ = intr[6]; s.d3_ = intr[7];
  s.obser_thr_ = obser_thr;
  s.Rcw_all_.assign(M, Eigen::Matrix3d()); s.tcw_all_.assign(M, Eigen::Vector3d());
  s.all_depths_.clear(); s.all_keypoints_.assign(M, {}); s.images_ids_.assign(M, 0.0);
  const size_t hw = (size_t)height * width;
  for (int k = 0; k < M; ++k) {
    IMUST x = pose_in(cams + 12 * k, 0);
    s.Rcw_all_[k] = x.R; s.tcw_all_[k] = x.p; s.images_ids_[k] = k;
    cv::Mat d(height, width, CV_32FC1);
    for (int y = 0; y < height; ++y) for (int xx = 0; xx < width; ++xx) d.at<float>(y, xx) = depth[k * hw + (size_t)y * width + xx];
    s.all_depths_.push_back(d);
    for (int64_t q = kp_ptr[k]; q < kp_ptr[k + 1]; ++q) { sift::Keypoint kp{}; kp.x = kp_uv[2 * q]; kp.y = kp_uv[2 * q + 1]; s.all_keypoints_[k].push_back(kp); }
  }
  s.Rcw_all_optimized_ = s.Rcw_all_; s.tcw_all_optimized_ = s.tcw_all_;
  s.all_matches_.assign((size_t)M * (M - 1) / 2, {});
  for (int64_t m = 0; m < n_matches; ++m) {
    const int a = matches[4 * m], ka = matches[4 * m + 1], b = matches[4 * m + 2], kb = matches[4 * m + 3];
    s.all_matches_[lvba::pairIndex(a, b, M)].push_back({ka, kb});          // a < b, in the caller's order
  }
}
void sys_set_keypoints_and_matches(void* h, const int64_t* kp_ptr, const float* kp_uv, int64_t n_matches, const int32_t* matches) {
  auto& s = *static_cast<Sys*>(h)->s;
  const int M = (int)s.images_ids_.size();
  s.all_keypoints_.assign(M, {});
  for (int k = 0; k < M; ++k)
    for (int64_t q = kp_ptr[k]; q < kp_ptr[k + 1]; ++q) { sift::Keypoint kp{}; kp.x = kp_uv[2 * q]; kp.y = kp_uv[2 * q + 1]; s.all_keypoints_[k].push_back(kp); }
  s.all_matches_.assign((size_t)M * (M - 1) / 2, {});
  for (int64_t m = 0; m < n_matches; ++m)
    s.all_matches_[lvba::pairIndex(matches[4 * m], matches[4 * m + 2], M)].push_back({matches[4 * m + 1], matches[4 * m + 3]});
}
// loadFromColmapDB (:510-685) against a real database (SQLite = the system's library): fills all_keypoints_ / all_matches_ of the images
// sys_set_camera listed.  dataset_path: the directory getImagePath (:2146) names the images under; returns what the reference returns.
int sys_load_colmap_db(void* h, const char* dataset_path, const char* db_path) {
  auto& s = *static_cast<Sys*>(h)->s;
  s.dataset_path_ = dataset_path; s.dataset_io_->dataset_path_ = dataset_path; s.dataset_io_->colmap_db_path_ = db_path;
  return s.loadFromColmapDB() ? 1 : 0;
}
void sys_frontend_sizes(void* h, int64_t* n_kp, int64_t* n_matches) {
  auto& s = *static_cast<Sys*>(h)->s;
  *n_kp = *n_matches = 0;
  for (auto& k : s.all_keypoints_) *n_kp += (int64_t)k.size();
  for (auto& m : s.all_matches_) *n_matches += (int64_t)m.size();
}
// keypoints per image (kp_ptr: M + 1) and matches as (img_a, kp_a, img_b, kp_b) in all_matches_ order (pairIndex order, then stored order)
void sys_get_frontend(void* h, int64_t* kp_ptr, float* kp_uv, int32_t* matches) {
  auto& s = *static_cast<Sys*>(h)->s;
  const int M = (int)s.all_keypoints_.size();
  int64_t q = 0;
  kp_ptr[0] = 0;
  for (int k = 0; k < M; ++k) { for (auto& kp : s.all_keypoints_[k]) { kp_uv[2 * q] = kp.x; kp_uv[2 * q + 1] = kp.y; ++q; } kp_ptr[k + 1] = q; }
  int64_t m = 0;
  for (int a = 0; a < M - 1; ++a)
    for (int b = a + 1; b < M; ++b)
      for (auto& pr : s.all_matches_[lvba::pairIndex(a, b, M)]) { matches[4 * m] = a; matches[4 * m + 1] = pr.first; matches[4 * m + 2] = b; matches[4 * m + 3] = pr.second; ++m; }
}
// The COLMAP text model as the reference writes it — VisualizeOptComparison (:1932-2143, reached through pubRGBCloud -> showTracksComparePCL at the end of
// runVisualBAWithLidarAssist): images.txt (one pose line + "0.0 0.0 -1" per image that loads and has LiDAR within +-0.5 s) and points3D.txt (the LiDAR points
// nearest per pixel in every image, merged, down_sampling_voxel2 at colmap_output/filter_size_points3D, coloured from the image).  There is no image decoder
// here: every image file reads as width x height pixels of one grey value (ref_shim/opencv2/opencv.hpp), so the colours written are that grey.
// Uses Rcw_all_ / Rcw_all_optimized_ as generateDepthWithVoxel left them; frees dataset_io_->pl_fulls_ at the end, as the reference does.
void sys_colmap_export(void* h, const char* dataset_path, int width, int height, int grey) {
  auto& s = *static_cast<Sys*>(h)->s;
  s.dataset_path_ = dataset_path;
  cv::stub_image().rows = height; cv::stub_image().cols = width; cv::stub_image().grey = grey;
  s.VisualizeOptComparison(s.images_ids_, true);
  s.fout_poses_after.close(); s.fout_points_after.close();
  cv::stub_image().rows = 0;
}
// BuildTracksAndFuse3D (:921-1263).  Returns the number of tracks; sizes through the two counters.
int64_t sys_build_tracks(void* h, int64_t* n_obs, int64_t* n_inl) {
  auto& s = *static_cast<Sys*>(h)->s;
  s.BuildTracksAndFuse3D();
  *n_obs = *n_inl = 0;
  for (const auto& t : s.tracks_) { *n_obs += (int64_t)t.observations.size(); *n_inl += (int64_t)t.inlier_indices.size(); }
  return (int64_t)s.tracks_.size();
}
void sys_get_tracks(void* h, int64_t* obs_ptr, int32_t* obs, int64_t* inl_ptr, int32_t* inl, double* Xw) {
  auto& s = *static_cast<Sys*>(h)->s;
  int64_t o = 0, q = 0;
  obs_ptr[0] = inl_ptr[0] = 0;
  for (size_t k = 0; k < s.tracks_.size(); ++k) {
    const auto& t = s.tracks_[k];
    for (auto& ob : t.observations) { obs[2 * o] = ob.first; obs[2 * o + 1] = ob.second; ++o; }
    for (int i : t.inlier_indices) inl[q++] = i;
    obs_ptr[k + 1] = o; inl_ptr[k + 1] = q;
    for (int j = 0; j < 3; ++j) Xw[3 * k + j] = t.Xw_fused(j);
  }
}

// TriangulateTrackDLT (:52-111) and ComputeMeanReproj (:8-50) — file-scope helpers of src/lvba_system.cpp, reachable because that file is part of this
// translation unit — on CSR tracks (one observation per camera): obs_ptr [T+1], obs_cam, obs_uv [n][2]; cams [M][12] = (Rcw, tcw).
// Xw_in == nullptr: triangulate every track (Xw, mean, count, ok out); otherwise the mean reprojection error of the given point per track.
void sys_track_helpers(int64_t T, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int M, const double* cams, const double* intr,
                       const double* Xw_in, int min_count, double* Xw, double* mean, int32_t* count, uint8_t* ok) {
  lvba::CameraIntrinsics cam;
  cam.fx = intr[0]; cam.fy = intr[1]; cam.cx = intr[2]; cam.cy = intr[3]; cam.k1 = intr[4]; cam.k2 = intr[5]; cam.p1 = intr[6]; cam.p2 = intr[7];
  std::vector<Eigen::Matrix3d> R(M); std::vector<Eigen::Vector3d> t(M);
  for (int k = 0; k < M; ++k) { IMUST x = pose_in(cams + 12 * k, 0); R[k] = x.R; t[k] = x.p; }
  for (int64_t tr = 0; tr < T; ++tr) {
    const int n = (int)(obs_ptr[tr + 1] - obs_ptr[tr]);
    std::vector<std::vector<sift::Keypoint>> kps(M, std::vector<sift::Keypoint>(n));
    std::vector<std::pair<int, int>> component;
    std::unordered_map<int, int> selected;
    for (int j = 0; j < n; ++j) {
      const int c = obs_cam[obs_ptr[tr] + j];
      kps[c][j].x = obs_uv[2 * (obs_ptr[tr] + j)]; kps[c][j].y = obs_uv[2 * (obs_ptr[tr] + j) + 1];
      component.push_back({c, j});
      selected[c] = j;
    }
    double m = 0.0; int cnt = 0; bool good;
    if (Xw_in == nullptr) {
      Eigen::Vector3d X = Eigen::Vector3d::Zero();
      good = lvba::TriangulateTrackDLT(selected, component, kps, R, t, cam, X, m, cnt);
      for (int j = 0; j < 3; ++j) Xw[3 * tr + j] = X(j);
    } else {
      good = lvba::ComputeMeanReproj(Eigen::Vector3d(Xw_in[3 * tr], Xw_in[3 * tr + 1], Xw_in[3 * tr + 2]), selected, component, kps, R, t, cam, min_count, m, cnt);
    }
    mean[tr] = m; count[tr] = cnt; ok[tr] = good ? 1 : 0;
  }
}

// Iteration order of a std::unordered_map<int,int> after reserve(reserve_n) and the insertion of `keys` in the given order — the order the
// reference's three `for (auto& kv : map)` loops of the track fusion run in (:1057, :1069, :1124).  The language leaves it unspecified; this is the
// answer of the C++ library the reference is compiled with here.
void sys_unordered_map_order(int64_t reserve_n, int64_t n, const int32_t* keys, int32_t* out) {
  std::unordered_map<int, int> m;
  m.reserve((size_t)reserve_n);
  for (int64_t i = 0; i < n; ++i) if (!m.count(keys[i])) m[keys[i]] = (int)i;
  int64_t k = 0;
  for (const auto& kv : m) out[k++] = kv.first;
}

// bucket_count() of a std::unordered_map<int,int> after reserve(n): the library's own answer
uint64_t sys_bucket_count_after_reserve(uint64_t n) {
  std::unordered_map<int, int> m;
  m.reserve((size_t)n);
  return (uint64_t)m.bucket_count();
}

// optimizeCameraPoses (:1409-1660) with the problem it builds recorded and handed to `cb` in place of ceres::Solve; after `cb` returns the
// reference's own code reads the parameter blocks back (:1651-1667).  Returns the number of residual blocks recorded.
int64_t sys_optimize_camera_poses(void* h, void (*cb)(void*), void* user) {
  Sys* y = static_cast<Sys*>(h);
  y->solve_cb = cb; y->solve_user = user;
  y->obs.clear(); y->planes.clear();
  ceres::solve_hook() = [y](const ceres::Solver::Options& o, ceres::Problem* p, ceres::Solver::Summary* sm) { on_solve(y, o, p, sm); };
  y->s->optimizeCameraPoses();
  ceres::solve_hook() = nullptr;
  return (int64_t)(y->obs.size() + y->planes.size());
}
void sys_problem_sizes(void* h, int64_t* n_cam, int64_t* n_pt, int64_t* n_obs, int64_t* n_pl) {
  Sys* y = static_cast<Sys*>(h);
  *n_cam = (int64_t)y->cam_q.size(); *n_pt = (int64_t)y->pts.size(); *n_obs = (int64_t)y->obs.size(); *n_pl = (int64_t)y->planes.size();
}
// valid only inside the callback (the parameter blocks live on optimizeCameraPoses' stack)
void sys_problem_get(void* h, double* q, double* t, uint8_t* cam_const, int32_t* cam_manifold, double* X, int32_t* obs_cam, int32_t* obs_pt,
                     double* obs_uv, double* obs_intr, double* obs_sigma, int32_t* obs_loss, int32_t* pl_pt, double* pl_nd, double* pl_sigma,
                     int32_t* pl_loss, double* options) {
  Sys* y = static_cast<Sys*>(h);
  for (size_t k = 0; k < y->cam_q.size(); ++k) {
    for (int j = 0; j < 4; ++j) q[4 * k + j] = y->cam_q[k][j];
    for (int j = 0; j < 3; ++j) t[3 * k + j] = y->cam_t[k][j];
    cam_const[k] = y->cam_const[k]; cam_manifold[k] = y->cam_manifold[k];
  }
  for (size_t k = 0; k < y->pts.size(); ++k) for (int j = 0; j < 3; ++j) X[3 * k + j] = y->pts[k][j];
  for (size_t k = 0; k < y->obs.size(); ++k) {
    const auto& o = y->obs[k];
    obs_cam[k] = o.cam; obs_pt[k] = o.pt; obs_uv[2 * k] = o.u; obs_uv[2 * k + 1] = o.v; obs_sigma[2 * k] = o.su; obs_sigma[2 * k + 1] = o.sv; obs_loss[k] = o.has_loss;
    for (int j = 0; j < 8; ++j) obs_intr[8 * k + j] = o.intr[j];
  }
  for (size_t k = 0; k < y->planes.size(); ++k) {
    const auto& p = y->planes[k];
    pl_pt[k] = p.pt; pl_sigma[k] = p.sigma; pl_loss[k] = p.has_loss;
    for (int j = 0; j < 4; ++j) pl_nd[4 * k + j] = p.nd[j];
  }
  for (int j = 0; j < 8; ++j) options[j] = y->opt[j];
}
void sys_problem_set(void* h, const double* q, const double* t, const double* X) {
  Sys* y = static_cast<Sys*>(h);
  for (size_t k = 0; k < y->cam_q.size(); ++k) {
    for (int j = 0; j < 4; ++j) y->cam_q[k][j] = q[4 * k + j];
    for (int j = 0; j < 3; ++j) y->cam_t[k][j] = t[3 * k + j];
  }
  for (size_t k = 0; k < y->pts.size(); ++k) for (int j = 0; j < 3; ++j) y->pts[k][j] = X[3 * k + j];
}
void sys_get_cameras_optimized(void* h, double* cams /* (M, 12) */) {
  auto& s = *static_cast<Sys*>(h)->s;
  for (size_t k = 0; k < s.Rcw_all_optimized_.size(); ++k) pose_out(s.Rcw_all_optimized_[k], s.tcw_all_optimized_[k], cams + 12 * k);
}

}  // extern "C"
