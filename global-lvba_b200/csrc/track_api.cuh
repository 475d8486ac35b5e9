// track_api.cuh — boundary B5 (second half): DLT triangulation and mean reprojection error of many tracks
// at once on the device.  Passes: track_pipeline.h; execution policy: CudaExec (voxel_api.cuh).  No host path.
#pragma once
#include "track_pipeline.h"
#include "voxel_api.cuh"

namespace lvba {

inline int tracks_check(int64_t n_tracks, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams,
                        const double* cams, const double* intr) {
  if (n_tracks < 0 || n_cams < 0 || !obs_ptr || !intr || (n_cams > 0 && !cams)) return fail(LVBA_ERR_INVALID_ARG, "null argument or negative count");
  if (obs_ptr[0] != 0) return fail(LVBA_ERR_INVALID_ARG, "obs_ptr[0] != 0");
  for (int64_t t = 0; t < n_tracks; ++t)
    if (obs_ptr[t + 1] < obs_ptr[t]) return fail(LVBA_ERR_INVALID_ARG, "obs_ptr not monotone at track %lld", (long long)t);
  if (obs_ptr[n_tracks] > 0 && (!obs_cam || !obs_uv)) return fail(LVBA_ERR_INVALID_ARG, "null observation arrays");
  for (int64_t k = 0; k < (int64_t)n_cams * 12; ++k)
    if (!std::isfinite(cams[k])) return fail(LVBA_ERR_INVALID_ARG, "non-finite camera entry %lld", (long long)k);
  for (int k = 0; k < 8; ++k)
    if (!std::isfinite(intr[k])) return fail(LVBA_ERR_INVALID_ARG, "non-finite intrinsic %d", k);
  return LVBA_OK;
}

// shared driver: upload, one pass, download
template <bool kTriangulate>
inline int tracks_run(int64_t n_tracks, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams, const double* cams,
                      const double* intr, int32_t device, const double* Xw_in, int32_t min_count, double* Xw_out, double* mean, int32_t* count,
                      uint8_t* ok) {
  LVBA_TRY(tracks_check(n_tracks, obs_ptr, obs_cam, obs_uv, n_cams, cams, intr));
  if (n_tracks > 0 && (!mean || !count || !ok || (kTriangulate ? !Xw_out : !Xw_in))) return fail(LVBA_ERR_INVALID_ARG, "null output / input array");
  LVBA_TRY(select_device(device));
  if (n_tracks == 0) return LVBA_OK;
  CudaExec ex;
  const int64_t n_obs = obs_ptr[n_tracks];
  DevBuf<int64_t> d_ptr;
  DevBuf<int32_t> d_cam, d_count;
  DevBuf<float> d_uv;
  DevBuf<double> d_cams, d_X, d_mean;
  DevBuf<uint8_t> d_ok;
  StreamDrain drain(nullptr);            // NULL stream
  LVBA_TRY(d_ptr.upload(obs_ptr, (size_t)n_tracks + 1, ex.stream));
  LVBA_TRY(d_cam.upload(obs_cam, (size_t)n_obs, ex.stream));
  LVBA_TRY(d_uv.upload(obs_uv, (size_t)n_obs * 2, ex.stream));
  LVBA_TRY(d_cams.upload(cams, (size_t)n_cams * 12, ex.stream));
  LVBA_TRY(d_mean.alloc((size_t)n_tracks)); LVBA_TRY(d_count.alloc((size_t)n_tracks)); LVBA_TRY(d_ok.alloc((size_t)n_tracks));
  if (kTriangulate) {
    LVBA_TRY(d_X.alloc((size_t)n_tracks * 3));
    track::TriangulateF f{d_ptr.p, d_cam.p, d_uv.p, n_cams, d_cams.p, {}, d_X.p, d_mean.p, d_count.p, d_ok.p};
    for (int q = 0; q < 8; ++q) f.intr[q] = intr[q];
    LVBA_TRY(ex.for_each(n_tracks, f));
    LVBA_CUDA(cudaMemcpyAsync(Xw_out, d_X.p, (size_t)n_tracks * 3 * sizeof(double), cudaMemcpyDeviceToHost, ex.stream));
  } else {
    LVBA_TRY(d_X.upload(Xw_in, (size_t)n_tracks * 3, ex.stream));
    track::MeanReprojF f{d_ptr.p, d_cam.p, d_uv.p, n_cams, d_cams.p, {}, d_X.p, min_count, d_mean.p, d_count.p, d_ok.p};
    for (int q = 0; q < 8; ++q) f.intr[q] = intr[q];
    LVBA_TRY(ex.for_each(n_tracks, f));
  }
  LVBA_CUDA(cudaMemcpyAsync(mean, d_mean.p, (size_t)n_tracks * sizeof(double), cudaMemcpyDeviceToHost, ex.stream));
  LVBA_CUDA(cudaMemcpyAsync(count, d_count.p, (size_t)n_tracks * sizeof(int32_t), cudaMemcpyDeviceToHost, ex.stream));
  LVBA_CUDA(cudaMemcpyAsync(ok, d_ok.p, (size_t)n_tracks, cudaMemcpyDeviceToHost, ex.stream));
  LVBA_CUDA(cudaStreamSynchronize(ex.stream));
  return LVBA_OK;
}

}  // namespace lvba

extern "C" {

int lvba_tracks_triangulate(int64_t n_tracks, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams,
                            const double* cams, const double intr[8], int32_t device, double* Xw, double* mean_reproj, int32_t* count,
                            uint8_t* ok) {
  return lvba::tracks_run<true>(n_tracks, obs_ptr, obs_cam, obs_uv, n_cams, cams, intr, device, nullptr, 4, Xw, mean_reproj, count, ok);
}

int lvba_tracks_mean_reproj(int64_t n_tracks, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams,
                            const double* cams, const double intr[8], int32_t device, const double* Xw, int32_t min_count,
                            double* mean_reproj, int32_t* count, uint8_t* ok) LVBA_ABI_BEGIN {
  if (min_count < 0) return lvba::fail(LVBA_ERR_INVALID_ARG, "min_count < 0");
  return lvba::tracks_run<false>(n_tracks, obs_ptr, obs_cam, obs_uv, n_cams, cams, intr, device, Xw, min_count, nullptr, mean_reproj, count, ok);
} LVBA_ABI_END("lvba_tracks_mean_reproj")

}  // extern "C"
