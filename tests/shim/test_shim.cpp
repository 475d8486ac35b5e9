// Compiles global-lvba_b200/host/lvba_shim.hpp against mock types shaped like the reference's
// IMUST / PointCluster / VOX_HESS (tools.hpp:147-207, 407-466, bavoxel.hpp:32-54) and calls through the C ABI.
// Exit codes: 0 = solved on a GPU, 2 = library refused for lack of a CUDA device (expected on CPU boxes).
#include <cmath>
#include <cstdio>
#include "../../global-lvba_b200/host/lvba_shim.hpp"

struct M3 { double m[9]; double& operator()(int r, int c) { return m[3 * r + c]; } double operator()(int r, int c) const { return m[3 * r + c]; } };
struct V3 { double v[3]; double& operator()(int r) { return v[r]; } double operator()(int r) const { return v[r]; } };
struct IMUST { M3 R; V3 p; };
struct PointCluster { M3 P{}; V3 v{}; int N = 0; void push(const double* x) { ++N; for (int i = 0; i < 3; ++i) { v.v[i] += x[i]; for (int j = 0; j < 3; ++j) P.m[3 * i + j] += x[i] * x[j]; } } };
struct VOX_HESS { std::vector<const std::vector<PointCluster>*> plvec_voxels; int win_size; };
struct PointXYZINormal { float x, y, z, pad0, nx, ny, nz, pad1, intensity, curvature, pad2, pad3; };   // 48 B, like PCL's
struct Cloud { std::vector<PointXYZINormal> points; };

// B3: raw scans of two perpendicular walls and a floor -> SurfMap -> damping_iter / plane lookup
static int surf_map_case() {
  const int W = 3;
  std::vector<IMUST> xs(W);
  std::vector<Cloud> store(W);
  std::vector<Cloud*> clouds;
  unsigned s = 777;
  auto rnd = [&] { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  for (int i = 0; i < W; ++i) {
    xs[i].R = M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; xs[i].p = V3{{0.3 * i, 0.0, 0.0}};
    for (int k = 0; k < 6000; ++k) {
      double w[3] = {4.0 * rnd(), 4.0 * rnd(), 4.0 * rnd()};
      w[k % 3] = (k % 3 == 0 ? 1.9 : k % 3 == 1 ? -1.7 : -1.2) + 0.01 * rnd();
      PointXYZINormal p{};
      p.x = (float)(w[0] - 0.3 * i); p.y = (float)w[1]; p.z = (float)w[2];
      store[i].points.push_back(p);
    }
    clouds.push_back(&store[i]);
  }
  const float ratios[4] = {0.3f, 0.1f, 0.06f, 0.03f};
  lvba_b200::SurfMap<std::vector<IMUST>> map;
  lvba_voxel_summary vs{};
  int rc = map.build(clouds, xs, 1.0, ratios, &vs);
  if (rc == LVBA_ERR_NO_DEVICE) return 2;
  if (rc != LVBA_OK) { std::printf("surf map error %d: %s\n", rc, lvba_last_error()); return 1; }
  if (vs.n_points != 3 * 6000 || vs.n_voxels < 10) { std::printf("surf map: %lld voxels\n", (long long)vs.n_voxels); return 1; }
  for (int i = 1; i < W; ++i) xs[i].p(1) += 0.02;                            // perturb, then let the LM pull it back
  lvba_summary sum{};
  rc = map.damping_iter(xs, 3, nullptr, &sum);
  if (rc != LVBA_OK || !(sum.cost_last < sum.cost_first)) { std::printf("surf map LM error %d\n", rc); return 1; }
  std::vector<std::array<double, 3>> Xs = {{0.5, 0.5, -1.2}, {1.9, 0.3, 0.4}, {50.0, 50.0, 50.0}}, pn;
  std::vector<double> pd;
  rc = map.recompute_local_planes(Xs, pn, pd);
  if (rc != LVBA_OK) return 1;
  const bool floor_ok = std::fabs(std::fabs(pn[0][2]) - 1.0) < 1e-2 && std::fabs(std::fabs(pd[0]) - 1.2) < 5e-2;
  const bool wall_ok = std::fabs(std::fabs(pn[1][0]) - 1.0) < 1e-2 && std::fabs(std::fabs(pd[1]) - 1.9) < 5e-2;
  const bool none_ok = pn[2][0] == 0 && pn[2][1] == 0 && pn[2][2] == 0 && pd[2] == 0;
  std::printf("surf map ok: %lld voxels from %lld points, LM %.3e -> %.3e, planes %d%d%d\n", (long long)vs.n_voxels,
              (long long)vs.n_points, sum.cost_first, sum.cost_last, (int)floor_ok, (int)wall_ok, (int)none_ok);
  return (floor_ok && wall_ok && none_ok) ? 0 : 1;
}

// B3 + B1 batched: two windows of the surf-map scene through run_window_stage; with_anchors: all of runWindowBA (B6 as well)
static int window_stage_case(bool with_anchors) {
  const int W = 6;
  std::vector<IMUST> xs(W);
  std::vector<Cloud> store(W);
  std::vector<Cloud*> clouds;
  unsigned s = 991;
  auto rnd = [&] { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  for (int i = 0; i < W; ++i) {
    xs[i].R = M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; xs[i].p = V3{{0.3 * i, 0.01 * (i % 3), 0.0}};
    for (int k = 0; k < 6000; ++k) {
      double w[3] = {4.0 * rnd() + 0.9, 4.0 * rnd(), 4.0 * rnd()};
      w[k % 3] = (k % 3 == 0 ? 2.9 : k % 3 == 1 ? -1.7 : -1.2) + 0.01 * rnd();
      PointXYZINormal p{};
      p.x = (float)(w[0] - 0.3 * i); p.y = (float)w[1]; p.z = (float)w[2];
      store[i].points.push_back(p);
    }
    clouds.push_back(&store[i]);
  }
  const float ratios[4] = {0.3f, 0.1f, 0.06f, 0.03f};
  std::vector<std::vector<IMUST>> x_wins;
  std::vector<lvba_summary> sums;
  const int rc = lvba_b200::run_window_stage(clouds, xs, 3, 1.0, ratios, x_wins, &sums);
  if (rc == LVBA_ERR_NO_DEVICE) return 2;
  if (rc != LVBA_OK) { std::printf("window stage error %d: %s\n", rc, lvba_last_error()); return 1; }
  if (x_wins.size() != 2 || x_wins[0].size() != 3 || x_wins[1].size() != 3) return 1;
  int solved = 0;
  for (const auto& sm : sums) if (sm.termination != LVBA_TERM_SKIPPED && sm.cost_last <= sm.cost_first) ++solved;
  std::printf("window stage ok: %zu windows, %d solved, costs %.3e -> %.3e\n", x_wins.size(), solved, sums[0].cost_first, sums[0].cost_last);
  if (solved != 2) return 1;
  if (!with_anchors) return 0;
  // all of runWindowBA: anchors (B6) for both windows, every frame attached to its anchor
  lvba_b200::WindowBAResult<std::vector<IMUST>> wba;
  const int rb = lvba_b200::run_window_ba(clouds, xs, 3, 1.0, ratios, 0.1, /*use_window_ba_rel=*/true, wba);
  if (rb != LVBA_OK) { std::printf("window BA error %d: %s\n", rb, lvba_last_error()); return 1; }
  size_t pts = 0;
  for (const auto& c : wba.anchor_clouds) pts += c.points.size();
  bool attached = wba.anchor_poses.size() == 2 && wba.anchor_clouds.size() == 2 && pts > 1000 && pts < 6 * 6000;
  for (int i = 0; i < W; ++i) attached = attached && wba.anchor_index_per_frame[i] == i / 3;
  attached = attached && std::fabs(wba.rel_poses_to_anchor[0].p(0)) < 1e-12 && std::fabs(wba.rel_poses_to_anchor[1].p(0) - 0.3) < 0.05;
  std::printf("window BA ok: %zu anchors, %zu anchor points, frames attached %d\n", wba.anchor_poses.size(), pts, (int)attached);
  return attached ? 0 : 1;
}

// B4: three frames of a wall 4 m in front of a forward-looking camera -> every covered pixel reads ~4 m
struct IMUST_T { M3 R; V3 p; double t; };
static int depth_case() {
  const int F = 3;
  std::vector<IMUST_T> xs(F);
  std::vector<Cloud> store(F);
  std::vector<Cloud*> clouds;
  unsigned s = 4242;
  auto rnd = [&] { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  for (int i = 0; i < F; ++i) {
    xs[i].R = M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; xs[i].p = V3{{0.0, 0.1 * i, 0.0}}; xs[i].t = 10.0 + 0.1 * i;
    for (int k = 0; k < 20000; ++k) { PointXYZINormal p{}; p.x = (float)(6.0 * rnd()); p.y = (float)(6.0 * rnd() - 0.1 * i); p.z = 4.0f; store[i].points.push_back(p); }
    clouds.push_back(&store[i]);
  }
  lvba_b200::DepthRenderer dr;
  int rc = dr.buildGridMapFromOptimized(clouds, xs);
  if (rc == LVBA_ERR_NO_DEVICE) return 2;
  if (rc != LVBA_OK) { std::printf("depth grid error %d: %s\n", rc, lvba_last_error()); return 1; }
  std::vector<M3> Rcw = {M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}, M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}};
  std::vector<V3> tcw = {V3{{0, 0, 0}}, V3{{0, 0, 1.0}}};
  std::vector<double> ts = {10.1, 99.0};                                       // the second image has no frame in its window
  std::vector<float> depth;
  rc = dr.generateDepthWithVoxel(Rcw, tcw, ts, 100, 100, 80, 60, 0, 0, 0, 0, 160, 120, depth);
  if (rc != LVBA_OK) { std::printf("depth render error %d: %s\n", rc, lvba_last_error()); return 1; }
  int filled = 0, wrong = 0, second = 0;
  for (int i = 0; i < 160 * 120; ++i) { if (depth[i] != 0) { ++filled; if (std::fabs(depth[i] - 4.0f) > 1e-6f) ++wrong; } if (depth[160 * 120 + i] != 0) ++second; }
  std::printf("depth ok: %d pixels filled, %d off the wall, %d in the empty-window image\n", filled, wrong, second);
  return (filled > 10000 && wrong == 0 && second == 0) ? 0 : 1;
}

// B7: a ring of cameras looking at a few points -> build_tracks_and_fuse_3d with mock sift::Keypoint / match tables
struct Keypoint { float x, y; };
static int fuse_case() {
  const int N = 6, NP = 12;
  const double fx = 646.78472, fy = 646.65775, cx = 313.456795, cy = 261.399612;
  std::vector<M3> Rcw(N); std::vector<V3> tcw(N);
  for (int i = 0; i < N; ++i) { Rcw[i] = M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; tcw[i] = V3{{-0.6 * i, 0.0, 0.0}}; }     // camera centres at x = 0.6 i
  std::vector<std::vector<Keypoint>> kps(N);
  std::vector<std::vector<std::pair<int, int>>> matches((size_t)N * (N - 1) / 2);
  std::vector<double> kpX; std::vector<uint8_t> kpV;
  unsigned s = 99;
  auto rnd = [&] { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  std::vector<std::array<double, 3>> pts(NP);
  for (auto& p : pts) p = {1.5 + 2.0 * rnd(), 1.0 * rnd(), 6.0 + 2.0 * rnd()};
  std::vector<std::vector<int>> kp_of(NP, std::vector<int>(N, -1));
  for (int i = 0; i < N; ++i)
    for (int p = 0; p < NP; ++p) {
      const double X = pts[p][0] - 0.6 * i, Y = pts[p][1], Z = pts[p][2];
      kp_of[p][i] = (int)kps[i].size();
      kps[i].push_back(Keypoint{(float)(fx * X / Z + cx), (float)(fy * Y / Z + cy)});
    }
  for (int i = 0; i < N; ++i)                                                    // depth candidates in (image, keypoint) order
    for (int p = 0; p < NP; ++p) { for (int d = 0; d < 3; ++d) kpX.push_back(pts[p][d] + 0.005 * rnd()); kpV.push_back(p % 4 == 3 ? 0 : 1); }
  for (int i = 0; i < N - 1; ++i)
    for (int j = i + 1; j < N; ++j)
      for (int p = 0; p < NP; ++p) matches[(size_t)i * N - (size_t)i * (i + 1) / 2 + (j - i - 1)].push_back({kp_of[p][i], kp_of[p][j]});
  std::vector<lvba_b200::FusedTrack> tracks;
  lvba_fuse_summary fs{};
  lvba_fuse_opts fo;
  lvba_fuse_default_opts(&fo);
  fo.map_order = LVBA_FUSE_ORDER_ASCENDING;     // this case predates the container-order mode (the default since); that mode has its own device tests
  const int rc = lvba_b200::build_tracks_and_fuse_3d(kps, matches, Rcw, tcw, fx, fy, cx, cy, 0.0, 0.0, 0.0, 0.0, kpX, kpV, tracks, &fo, &fs);
  if (rc == LVBA_ERR_NO_DEVICE) return 2;
  if (rc != LVBA_OK) { std::printf("fuse error %d: %s\n", rc, lvba_last_error()); return 1; }
  double worst = 0;
  for (const auto& t : tracks) {
    if ((int)t.observations.size() != N || (int)t.inlier_indices.size() < 3) { std::printf("fuse: bad track shape\n"); return 1; }
    const int p = t.observations[0].second;                                       // keypoint index == point index in image 0
    for (int d = 0; d < 3; ++d) worst = std::fmax(worst, std::fabs(t.Xw_fused[d] - pts[p][d]));
  }
  std::printf("fuse ok: %zu tracks (%lld depth, %lld triangulated), worst |X - truth| = %.2e\n", tracks.size(), (long long)fs.n_depth_selected,
              (long long)fs.n_tri_selected, worst);
  return ((int)tracks.size() == NP && worst < 0.05) ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "fuse") return fuse_case();             // B7, run by tests/test_zz_fuse_gpu.py
  if (argc > 1 && std::string(argv[1]) == "windows") return window_stage_case(false);   // B3 + B1 batched, run by tests/test_zz_voxel_gpu.py
  if (argc > 1 && std::string(argv[1]) == "windowba") return window_stage_case(true);   // ... and the anchors (B6), tests/test_zz_offline_gpu.py
  if (argc > 1 && std::string(argv[1]) == "depth") return depth_case();           // B4, run by tests/test_zz_depth_gpu.py
  if (argc > 1 && std::string(argv[1]) == "surfmap") return surf_map_case();     // B3, run by tests/test_zz_voxel_gpu.py
  const int W = 4;
  std::vector<IMUST> xs(W);
  for (int i = 0; i < W; ++i) { xs[i].R = M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; xs[i].p = V3{{0.5 * i, 0.01 * i, 0}}; }
  std::vector<std::vector<PointCluster>> store(6, std::vector<PointCluster>(W));
  VOX_HESS vh; vh.win_size = W;
  unsigned s = 12345;
  auto rnd = [&] { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  for (int a = 0; a < 6; ++a) {
    const double n[3] = {a % 3 == 0 ? 1.0 : 0.0, a % 3 == 1 ? 1.0 : 0.0, a % 3 == 2 ? 1.0 : 0.0};
    for (int i = 0; i < W; ++i) for (int k = 0; k < 20; ++k) {
      double w[3] = {3 + rnd(), 2 + rnd(), 1 + rnd()};
      for (int d = 0; d < 3; ++d) if (n[d] == 1.0) w[d] = 2.0 + 0.01 * rnd();
      const double b[3] = {w[0] - 0.5 * i, w[1], w[2]};          // ground-truth body frame (true p = (0.5 i, 0, 0))
      store[a][i].push(b);
    }
    vh.plvec_voxels.push_back(&store[a]);
  }
  lvba_summary sum{};
  const int rc = lvba_b200::damping_iter(xs, vh, nullptr, &sum);
  if (rc == LVBA_ERR_NO_DEVICE) { std::printf("no device: %s\n", lvba_last_error()); return 2; }
  if (rc != LVBA_OK) { std::printf("error %d: %s\n", rc, lvba_last_error()); return 1; }
  std::printf("ok: %d iterations, cost %.3e -> %.3e\n", sum.iterations, sum.cost_first, sum.cost_last);
  if (!(sum.cost_last <= sum.cost_first)) return 1;
  // the same window twice through the batched entry point (runWindowBA): both copies must land on the single-call result
  std::vector<IMUST> xa(W), xb(W);
  for (int i = 0; i < W; ++i) { xa[i].R = M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; xa[i].p = V3{{0.5 * i, 0.01 * i, 0}}; xb[i] = xa[i]; }
  lvba_b200::WindowBatch<std::vector<IMUST>> batch;
  batch.add(xa, vh);
  batch.add(xb, vh);
  std::vector<lvba_summary> sums;
  const int rb = batch.solve(/*min_voxels_per_pose=*/1, nullptr, &sums);
  if (rb != LVBA_OK) { std::printf("batch error %d: %s\n", rb, lvba_last_error()); return 1; }
  double worst = 0;
  for (int i = 0; i < W; ++i) for (int d = 0; d < 3; ++d) {
    worst = std::fmax(worst, std::fabs(xa[i].p(d) - xs[i].p(d)));
    worst = std::fmax(worst, std::fabs(xb[i].p(d) - xs[i].p(d)));
  }
  std::printf("batch ok: %d windows, %d / %d iterations, max |p - single| = %.2e\n", batch.size(), sums[0].iterations, sums[1].iterations, worst);
  return (worst < 1e-9 && sums[0].iterations == sum.iterations) ? 0 : 1;
}
