// fuse_api.cuh — boundary B7: LvbaSystem::BuildTracksAndFuse3D (reference src/lvba_system.cpp:921-1263)
// as ONE call: pairwise matches + keypoints + per-keypoint depth candidates (boundary B4, lvba_depth_backproject) + camera
// poses in, fused 3-D landmarks with their observation / inlier lists out — the CSR lvba_visual_lm takes.  The match graph and
// the retry rounds run on the host (fuse_pipeline.h documents which reference lines they mirror and the one freedom the
// reference leaves open: the order of its three unordered_map loops — lvba_fuse_opts::map_order); the per-component work
// (gating, view-angle filter, DLT, reprojection tests, choice) is one device item per component.  No host path.
#pragma once
#include "fuse_pipeline.h"
#include "voxel_api.cuh"

struct lvba_track_set {
  lvba::fuse::Result res;
  lvba_fuse_summary sum{};
};

extern "C" {

void lvba_fuse_default_opts(lvba_fuse_opts* o) {
  if (!o) return;
  o->obser_thr = 3;                   // include/lvba_system.h:139
  o->min_view_angle_deg = 8.0;        // track_fusion/min_view_angle, src/lvba_system.cpp:129
  o->reproj_mean_thr_px = 3.0;        // track_fusion/reproj_mean_thr, :130
  o->depth_gate_m = 0.12;             // :1050
  o->device = -1;
  o->map_order = LVBA_FUSE_ORDER_LIBSTDCXX;   // what a g++ build of the reference does; verified against the reference's own source (tests/test_zzz_ref_gpu.py)
}

int lvba_tracks_fuse_create(int32_t n_images, const int64_t* kp_ptr, const float* kp_uv, int64_t n_matches, const int32_t* match_img_a,
                            const int32_t* match_kp_a, const int32_t* match_img_b, const int32_t* match_kp_b, const double* cams,
                            const double intr[8], const double* kp_Xw, const uint8_t* kp_valid, const lvba_fuse_opts* opts,
                            lvba_track_set** out, lvba_fuse_summary* summary) LVBA_ABI_BEGIN {
  using namespace lvba;
  if (!out) return fail(LVBA_ERR_INVALID_ARG, "out is null");
  *out = nullptr;
  if (n_images <= 0 || n_matches < 0 || !kp_ptr || !cams || !intr) return fail(LVBA_ERR_INVALID_ARG, "null argument or bad count");
  if (kp_ptr[0] != 0) return fail(LVBA_ERR_INVALID_ARG, "kp_ptr[0] != 0");
  for (int32_t i = 0; i < n_images; ++i)
    if (kp_ptr[i + 1] < kp_ptr[i]) return fail(LVBA_ERR_INVALID_ARG, "kp_ptr not monotone at image %d", i);
  const int64_t n_kp = kp_ptr[n_images];
  if (n_kp > 0 && (!kp_uv || !kp_Xw || !kp_valid)) return fail(LVBA_ERR_INVALID_ARG, "null keypoint arrays");
  if (n_matches > 0 && (!match_img_a || !match_kp_a || !match_img_b || !match_kp_b)) return fail(LVBA_ERR_INVALID_ARG, "null match arrays");
  lvba_fuse_opts o;
  if (opts) o = *opts; else lvba_fuse_default_opts(&o);
  if (o.obser_thr < 1 || !(o.reproj_mean_thr_px >= 0) || !(o.depth_gate_m >= 0) || !std::isfinite(o.min_view_angle_deg))
    return fail(LVBA_ERR_INVALID_ARG, "bad fusion options");
  if (o.map_order != LVBA_FUSE_ORDER_ASCENDING && o.map_order != LVBA_FUSE_ORDER_LIBSTDCXX)
    return fail(LVBA_ERR_INVALID_ARG, "map_order %d is neither LVBA_FUSE_ORDER_ASCENDING nor LVBA_FUSE_ORDER_LIBSTDCXX", (int)o.map_order);
  for (int64_t k = 0; k < (int64_t)n_images * 12; ++k)
    if (!std::isfinite(cams[k])) return fail(LVBA_ERR_INVALID_ARG, "non-finite camera entry %lld", (long long)k);
  for (int k = 0; k < 8; ++k)
    if (!std::isfinite(intr[k])) return fail(LVBA_ERR_INVALID_ARG, "non-finite intrinsic %d", k);
  for (int64_t g = 0; g < n_kp; ++g)
    if (kp_valid[g] && !(std::isfinite(kp_Xw[3 * g]) && std::isfinite(kp_Xw[3 * g + 1]) && std::isfinite(kp_Xw[3 * g + 2])))
      return fail(LVBA_ERR_INVALID_ARG, "keypoint %lld: valid depth candidate with a non-finite coordinate", (long long)g);
  LVBA_TRY(select_device(o.device));
  const double t0 = wall_ms();
  std::unique_ptr<lvba_track_set> h(new lvba_track_set());
  CudaExec ex;
  StreamDrain drain(nullptr);
  fuse::Params prm{o.obser_thr, std::cos(o.min_view_angle_deg * M_PI / 180.0), o.reproj_mean_thr_px, o.depth_gate_m, (int)o.map_order};
  const int rc = fuse::run(ex, n_images, kp_ptr, kp_uv, n_matches, match_img_a, match_kp_a, match_img_b, match_kp_b, cams, intr, kp_Xw,
                           kp_valid, prm, h->res);
  if (rc != LVBA_OK) return rc;
  lvba_fuse_summary& s = h->sum;
  s.n_keypoints = n_kp; s.n_components = h->res.n_components; s.n_candidates = h->res.n_candidates;
  s.n_tracks = (int64_t)h->res.tracks.size(); s.n_depth_selected = h->res.n_depth; s.n_tri_selected = h->res.n_tri;
  s.n_rounds = h->res.n_rounds; s.n_attempts = h->res.n_attempts;
  s.n_obs = 0; s.n_inliers = 0;
  for (const auto& t : h->res.tracks) { s.n_obs += (int64_t)t.img.size(); for (uint8_t f : t.inlier) s.n_inliers += f; }
  s.kernel_launches = ex.launches;
  s.ms_total = wall_ms() - t0;
  if (summary) *summary = s;
  *out = h.release();
  return LVBA_OK;
} LVBA_ABI_END("lvba_tracks_fuse_create")

int lvba_tracks_fuse_summary(const lvba_track_set* s, lvba_fuse_summary* summary) LVBA_ABI_BEGIN {
  if (!s || !summary) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  *summary = s->sum;
  return LVBA_OK;
} LVBA_ABI_END("lvba_tracks_fuse_summary")

int lvba_tracks_fuse_export(lvba_track_set* s, int64_t* obs_ptr, int32_t* obs_img, int32_t* obs_kp, uint8_t* obs_inlier, double* Xw,
                            uint8_t* source, double* mean_reproj) LVBA_ABI_BEGIN {
  if (!s || !obs_ptr) return lvba::fail(LVBA_ERR_INVALID_ARG, "null argument");
  int64_t o = 0;
  obs_ptr[0] = 0;
  for (size_t i = 0; i < s->res.tracks.size(); ++i) {
    const auto& t = s->res.tracks[i];
    for (size_t q = 0; q < t.img.size(); ++q, ++o) {
      if (obs_img) obs_img[o] = t.img[q];
      if (obs_kp) obs_kp[o] = t.kp[q];
      if (obs_inlier) obs_inlier[o] = t.inlier[q];
    }
    obs_ptr[i + 1] = o;
    if (Xw) for (int q = 0; q < 3; ++q) Xw[3 * i + q] = t.Xw[q];
    if (source) source[i] = t.source;
    if (mean_reproj) mean_reproj[i] = t.mean;
  }
  return LVBA_OK;
} LVBA_ABI_END("lvba_tracks_fuse_export")

int lvba_tracks_fuse_destroy(lvba_track_set* s) LVBA_ABI_BEGIN {
  delete s;
  return LVBA_OK;
} LVBA_ABI_END("lvba_tracks_fuse_destroy")

}  // extern "C"
