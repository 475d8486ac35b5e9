// track_emu.cpp — TEST INFRASTRUCTURE: the per-track functors of global-lvba_b200/csrc/track_pipeline.h run by plain loops
// for tests/test_track_emu.py (comparison with a numpy restatement).  Never part of the product.
#include "../../global-lvba_b200/csrc/track_pipeline.h"
#include "host_exec.h"

extern "C" {
int emu_tracks_triangulate(int64_t n, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams, const double* cams,
                           const double* intr, double* Xw, double* mean, int32_t* count, uint8_t* ok) {
  lvba::track::TriangulateF f{obs_ptr, obs_cam, obs_uv, n_cams, cams, {}, Xw, mean, count, ok};
  for (int q = 0; q < 8; ++q) f.intr[q] = intr[q];
  HostExec().for_each(n, f);
  return 0;
}
int emu_tracks_mean_reproj(int64_t n, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams, const double* cams,
                           const double* intr, const double* Xw, int32_t min_count, double* mean, int32_t* count, uint8_t* ok) {
  lvba::track::MeanReprojF f{obs_ptr, obs_cam, obs_uv, n_cams, cams, {}, Xw, min_count, mean, count, ok};
  for (int q = 0; q < 8; ++q) f.intr[q] = intr[q];
  HostExec().for_each(n, f);
  return 0;
}
}
