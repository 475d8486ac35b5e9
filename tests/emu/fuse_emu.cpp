// fuse_emu.cpp — TEST INFRASTRUCTURE (never part of liblvba_b200.so): the track fusion of
// global-lvba_b200/csrc/fuse_pipeline.h (host component builder + rounds + the per-component device functor) run with the
// sequential host execution policy.  tests/test_fuse_emu.py compares it with oracle/fuse_oracle.py.
#include <cstring>
#include <vector>

#define LVBA_HOST_ONLY 1
#include "host_exec.h"
#include "../../global-lvba_b200/csrc/fuse_pipeline.h"

using namespace lvba;

static fuse::Result g_res;

// runs the stage; returns the number of tracks (negative: error).  counts: [0] observations of all tracks, [1] components,
// [2] candidates, [3] rounds, [4] attempts, [5] depth selected, [6] triangulation selected
extern "C" long long fuse_emu_run(int n_images, const long long* kp_ptr, const float* kp_uv, long long n_matches, const int* ma_img, const int* ma_kp,
                                  const int* mb_img, const int* mb_kp, const double* cams, const double* intr, const double* kp_Xw,
                                  const unsigned char* kp_valid, int obser_thr, double min_view_angle_deg, double reproj_thr, double depth_gate,
                                  long long* counts, int map_order) {
  HostExec ex;
  fuse::Params prm{obser_thr, std::cos(min_view_angle_deg * M_PI / 180.0), reproj_thr, depth_gate, map_order};
  static_assert(sizeof(long long) == sizeof(int64_t), "");
  const int rc = fuse::run(ex, n_images, (const int64_t*)kp_ptr, kp_uv, n_matches, ma_img, ma_kp, mb_img, mb_kp, cams, intr, kp_Xw, kp_valid, prm, g_res);
  if (rc != 0) return -1;
  long long n_obs = 0;
  for (const auto& t : g_res.tracks) n_obs += (long long)t.img.size();
  counts[0] = n_obs; counts[1] = g_res.n_components; counts[2] = g_res.n_candidates; counts[3] = g_res.n_rounds; counts[4] = g_res.n_attempts;
  counts[5] = g_res.n_depth; counts[6] = g_res.n_tri;
  return (long long)g_res.tracks.size();
}
extern "C" void fuse_emu_export(long long* obs_ptr, int* obs_img, int* obs_kp, unsigned char* obs_inlier, double* Xw, unsigned char* source,
                                double* mean, long long* seed) {
  long long o = 0;
  obs_ptr[0] = 0;
  for (size_t i = 0; i < g_res.tracks.size(); ++i) {
    const auto& t = g_res.tracks[i];
    for (size_t q = 0; q < t.img.size(); ++q, ++o) { obs_img[o] = t.img[q]; obs_kp[o] = t.kp[q]; obs_inlier[o] = t.inlier[q]; }
    obs_ptr[i + 1] = o;
    for (int q = 0; q < 3; ++q) Xw[3 * i + q] = t.Xw[q];
    source[i] = t.source; mean[i] = t.mean; seed[i] = t.seed;
  }
}

// stl_bucket_count of fuse_pipeline.h on its own (tests/test_ref_system_pin.py holds it against the real container's bucket_count())
extern "C" unsigned emu_stl_bucket_count(unsigned n) {
  int np = 0;
  const uint32_t* pr = fuse::stl_prime_table(&np);
  return fuse::stl_bucket_count(n, pr, np);
}
