// lvba_offline — the LiDAR half of the reference's pipeline without ROS: loads a dataset directory in the
// reference's layout (global-lvba_b200/host/lvba_dataset.hpp), runs LvbaSystem::runLidarBA as configured with
// window_ba/enable = false (every scan is its own anchor, src/lvba_system.cpp:218-226): for stage 1 and stage 2 one adaptive
// voxel map over all scans (B3) and one BALM2::damping_iter on its plane voxels (B1), then writes the optimised trajectory
// as TUM lines.  Everything numeric happens in liblvba_b200.so; without a GPU the run stops with the library's error.
//
//   lvba_offline --data DIR [--out FILE] [--stage1-voxel 0.5] [--stage2-voxel 0.5] [--eigen1 a,b,c,d] [--eigen2 a,b,c,d]
//                [--no-stage1] [--window N] [--check]
//   --window N   window_ba/enable = true, window_ba/size = N: all of runWindowBA first (window stage = windowed voxel map + batched
//                LM; then the anchors: aligned poses, merged and down-sampled anchor clouds, boundary B6), the global stages run on
//                the ANCHORS, and every frame is placed through its anchor (src/lvba_system.cpp:391-403)
//   --anchor-leaf L (0.1)   --window-rel   window_ba/anchor_leaf_size, window_ba/use_window_ba_rel
//   --check      load and summarise the dataset only (no GPU needed)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../global-lvba_b200/host/lvba_dataset.hpp"
#include "../global-lvba_b200/host/lvba_shim.hpp"

using lvba_b200::dataset::LidarDataset;
using lvba_b200::dataset::Pose;

static bool parse4(const char* s, float out[4]) { return std::sscanf(s, "%f,%f,%f,%f", &out[0], &out[1], &out[2], &out[3]) == 4; }

int main(int argc, char** argv) {
  std::string data, out;
  double voxel[2] = {0.5, 0.5};                                              // BALM_stage1/2 root_voxel_size defaults (dataset_io.cpp:55-57)
  float eigen[2][4] = {{0.3f, 0.1f, 0.06f, 0.03f}, {0.3f, 0.1f, 0.06f, 0.03f}};   // bavoxel.hpp:17
  bool stage1 = true, check = false, window_rel = false;
  int window = 0;
  double anchor_leaf = 0.1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value after %s\n", a.c_str()); std::exit(64); } return argv[++i]; };
    if (a == "--data") data = next();
    else if (a == "--out") out = next();
    else if (a == "--stage1-voxel") voxel[0] = std::atof(next());
    else if (a == "--stage2-voxel") voxel[1] = std::atof(next());
    else if (a == "--eigen1") { if (!parse4(next(), eigen[0])) return 64; }
    else if (a == "--eigen2") { if (!parse4(next(), eigen[1])) return 64; }
    else if (a == "--no-stage1") stage1 = false;
    else if (a == "--window") window = std::atoi(next());
    else if (a == "--anchor-leaf") anchor_leaf = std::atof(next());
    else if (a == "--window-rel") window_rel = true;
    else if (a == "--check") check = true;
    else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 64; }
  }
  if (data.empty()) { std::fprintf(stderr, "usage: lvba_offline --data DIR [--out FILE] [--check] ...\n"); return 64; }
  if (data.back() != '/') data += '/';
  LidarDataset ds;
  std::string err;
  if (!lvba_b200::dataset::load_lidar_dataset(data, ds, &err)) { std::fprintf(stderr, "load failed: %s\n", err.c_str()); return 1; }
  for (const auto& w : ds.warnings) std::fprintf(stderr, "warning: %s\n", w.c_str());
  const size_t n = ds.size();
  size_t points = 0;
  double sum = 0.0, pose_sum = 0.0;
  for (size_t i = 0; i < n; ++i) { points += ds.clouds[i].points.size(); for (const auto& p : ds.clouds[i].points) sum += (double)p.x + (double)p.y + (double)p.z; }
  for (size_t i = 0; i < n; ++i) { for (int k = 0; k < 9; ++k) pose_sum += (k + 1) * ds.x_buf[i].R.m[k]; for (int k = 0; k < 3; ++k) pose_sum += (k + 10) * ds.x_buf[i].p.v[k]; }
  std::printf("{\"scans\": %zu, \"poses\": %zu, \"points\": %zu, \"coordinate_sum\": %.6f, \"pose_sum\": %.9f, \"first_ts\": %.6f, \"last_ts\": %.6f}\n",
              ds.clouds.size(), ds.x_buf.size(), points, sum, pose_sum, n ? ds.x_buf[0].t : 0.0, n ? ds.x_buf[n - 1].t : 0.0);
  if (check) return 0;
  std::vector<Pose> frames(ds.x_buf.begin(), ds.x_buf.begin() + (long)n);              // x_buf_full
  std::vector<lvba_b200::dataset::Cloud*> frame_clouds(ds.pl_fulls.begin(), ds.pl_fulls.begin() + (long)n);
  lvba_b200::WindowBAResult<std::vector<Pose>> wba;
  std::vector<lvba_b200::AnchorCloud*> anchor_clouds;
  const bool windows = window > 0;
  if (windows) {                                                             // runWindowBA (:205-316)
    const int rc = lvba_b200::run_window_ba(frame_clouds, frames, window, voxel[0], eigen[0], anchor_leaf, window_rel, wba);
    if (rc != LVBA_OK) { std::fprintf(stderr, "window BA failed (%d): %s\n", rc, lvba_last_error()); return rc == LVBA_ERR_NO_DEVICE ? 2 : 1; }
    size_t pts = 0;
    for (auto& c : wba.anchor_clouds) { anchor_clouds.push_back(&c); pts += c.points.size(); }
    std::printf("{\"stage\": \"windows\", \"windows\": %d, \"skipped\": %d, \"anchors\": %zu, \"anchor_points\": %zu}\n",
                wba.win_total, wba.win_skipped, wba.anchor_poses.size(), pts);
  }
  std::vector<Pose>& poses = windows ? wba.anchor_poses : frames;            // anchor_poses / anchor_clouds of runLidarBA (:329-334)
  for (int idx = stage1 ? 0 : 1; idx < 2 && !poses.empty(); ++idx) {         // the two passes (:358-389)
    lvba_voxel_summary vs{};
    lvba_summary s{};
    int rc;
    lvba_b200::SurfMap<std::vector<Pose>> surf;
    rc = windows ? surf.build(anchor_clouds, poses, voxel[idx], eigen[idx], &vs) : surf.build(frame_clouds, poses, voxel[idx], eigen[idx], &vs);
    if (rc != LVBA_OK) { std::fprintf(stderr, "stage %d voxel map failed (%d): %s\n", idx + 1, rc, lvba_last_error()); return rc == LVBA_ERR_NO_DEVICE ? 2 : 1; }
    rc = surf.damping_iter(poses, 0, nullptr, &s);
    if (rc != LVBA_OK) { std::fprintf(stderr, "stage %d LM failed (%d): %s\n", idx + 1, rc, lvba_last_error()); return 1; }
    std::printf("{\"stage\": %d, \"voxels\": %lld, \"clusters\": %lld, \"map_ms\": %.3f, \"iterations\": %d, \"accepted\": %d, \"cost_first\": %.9e, \"cost_last\": %.9e, \"lm_ms\": %.3f}\n",
                idx + 1, (long long)vs.n_voxels, (long long)vs.nnz, vs.ms_total, s.iterations, s.accepted, s.cost_first, s.cost_last, s.ms_total);
  }
  if (windows) {                                                             // optimized_x_buf_ (:391-403): frame = anchor * rel
    for (size_t i = 0; i < n; ++i) {
      const int a = wba.anchor_index_per_frame[i];
      if (a < 0) continue;                                                   // frames of skipped windows keep their odometry pose
      const Pose& A = wba.anchor_poses[(size_t)a];
      const Pose& r = wba.rel_poses_to_anchor[i];
      Pose o = frames[i];
      for (int x = 0; x < 3; ++x) {
        for (int y = 0; y < 3; ++y) o.R(x, y) = A.R(x, 0) * r.R(0, y) + A.R(x, 1) * r.R(1, y) + A.R(x, 2) * r.R(2, y);
        o.p(x) = (A.R(x, 0) * r.p(0) + A.R(x, 1) * r.p(1) + A.R(x, 2) * r.p(2)) + A.p(x);
      }
      frames[i] = o;
    }
  }
  if (out.empty()) out = data + "all_pcd_body/lidar_poses_optimized.txt";
  if (!lvba_b200::dataset::save_poses_tum(out, frames)) { std::fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
  std::printf("{\"written\": \"%s\", \"poses\": %zu}\n", out.c_str(), frames.size());
  return 0;
}
