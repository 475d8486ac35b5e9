// anchor_api.cuh — boundary B6: the anchor clouds of runWindowBA (transform every scan into its window's anchor frame, merge,
// keep per leaf voxel the point closest to the voxel centre) on the device.  Passes: anchor_pipeline.h.  No host path.
#pragma once
#include "anchor_pipeline.h"
#include "voxel_api.cuh"

struct lvba_anchor_clouds {
  lvba::anchor::AnchorClouds<lvba::CudaExec> ac;
  int device = 0;
  double ms_device = 0.0;
  ~lvba_anchor_clouds() { cudaStreamSynchronize(ac.ex.stream); }     // members are parked in the pool after this body: they must be idle
};

extern "C" {

int lvba_anchor_clouds_create(int32_t n_windows, const int32_t* win_ptr, const int64_t* scan_ptr, const float* xyz, int32_t xyz_stride_floats,
                              const double* rel_poses, double leaf, int32_t device, lvba_anchor_clouds** out, int64_t* n_points_out) LVBA_ABI_BEGIN {
  if (!out) return lvba::fail(LVBA_ERR_INVALID_ARG, "null output handle");
  *out = nullptr;
  if (n_windows <= 0 || !win_ptr || !scan_ptr) return lvba::fail(LVBA_ERR_INVALID_ARG, "n_windows=%d must be positive, win_ptr / scan_ptr non-null", n_windows);
  if (win_ptr[0] != 0) return lvba::fail(LVBA_ERR_INVALID_ARG, "win_ptr[0] must be 0");
  for (int w = 0; w < n_windows; ++w)
    if (win_ptr[w + 1] < win_ptr[w]) return lvba::fail(LVBA_ERR_INVALID_ARG, "win_ptr must be non-decreasing");
  const int32_t S = win_ptr[n_windows];
  if (xyz_stride_floats < 3) return lvba::fail(LVBA_ERR_INVALID_ARG, "xyz_stride %d < 3 floats", xyz_stride_floats);
  if (!(leaf >= 0.0) || !std::isfinite(leaf)) return lvba::fail(LVBA_ERR_INVALID_ARG, "leaf must be finite and >= 0");
  if (scan_ptr[0] != 0) return lvba::fail(LVBA_ERR_INVALID_ARG, "scan_ptr[0] != 0");
  for (int32_t j = 0; j < S; ++j)
    if (scan_ptr[j + 1] < scan_ptr[j]) return lvba::fail(LVBA_ERR_INVALID_ARG, "scan_ptr not monotone at scan %d", j);
  const int64_t N = scan_ptr[S];
  if (N >= (int64_t)0xfffffff0ll) return lvba::fail(LVBA_ERR_UNSUPPORTED, "%lld points: more than 2^32 per call", (long long)N);
  if ((N > 0 && !xyz) || (S > 0 && !rel_poses)) return lvba::fail(LVBA_ERR_INVALID_ARG, "null xyz / rel_poses");
  for (int64_t k = 0; k < (int64_t)S * 12; ++k)
    if (!std::isfinite(rel_poses[k])) return lvba::fail(LVBA_ERR_INVALID_ARG, "non-finite pose entry %lld", (long long)k);
  LVBA_TRY(lvba::select_device(device));
  std::unique_ptr<lvba_anchor_clouds> h(new lvba_anchor_clouds());
  cudaGetDevice(&h->device);
  lvba::CudaExec& ex = h->ac.ex;
  lvba::DevBuf<float> d_xyz;
  lvba::DevBuf<int64_t> d_scan;
  lvba::DevBuf<double> d_rel;
  lvba::DevBuf<int32_t> d_win;
  lvba::StreamDrain drain(nullptr);      // every path of this handle runs on the NULL stream
  std::vector<float> packed;
  const float* src = xyz;
  if (xyz_stride_floats != 3 && N > 0) {
    packed.resize((size_t)N * 3);
    float* dst = packed.data();
    const int stride = xyz_stride_floats;
    lvba::parallel_chunks(N, 1 << 16, [=](int64_t a, int64_t b, int) {
      for (int64_t i = a; i < b; ++i) { dst[3 * i] = xyz[i * stride]; dst[3 * i + 1] = xyz[i * stride + 1]; dst[3 * i + 2] = xyz[i * stride + 2]; }
    });
    src = packed.data();
  }
  LVBA_TRY(d_xyz.upload(src, (size_t)N * 3, ex.stream));
  LVBA_TRY(d_scan.upload(scan_ptr, (size_t)S + 1, ex.stream));
  LVBA_TRY(d_rel.upload(rel_poses, (size_t)S * 12, ex.stream));
  LVBA_TRY(d_win.upload(win_ptr, (size_t)n_windows + 1, ex.stream));
  lvba::EventPair ev;
  LVBA_TRY(ev.create());
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  LVBA_CUDA(cudaEventRecord(e0, ex.stream));
  const int rc = h->ac.build(d_xyz.p, d_scan.p, d_rel.p, d_win.p, S, n_windows, N, leaf);
  if (rc != LVBA_OK) {
    if (h->ac.error[0]) return lvba::fail(rc, "%s", h->ac.error);
    return rc;
  }
  LVBA_CUDA(cudaEventRecord(e1, ex.stream));
  LVBA_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  LVBA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  ex.temp.release();
  h->ms_device = ms;
  if (n_points_out) *n_points_out = h->ac.n_out;
  *out = h.release();
  return LVBA_OK;
} LVBA_ABI_END("lvba_anchor_clouds_create")

int lvba_anchor_clouds_export(lvba_anchor_clouds* a, int64_t* cloud_ptr, float* xyz, double* ms_device) LVBA_ABI_BEGIN {
  if (!a) return lvba::fail(LVBA_ERR_INVALID_ARG, "null handle");
  LVBA_CUDA(cudaSetDevice(a->device));
  cudaStream_t s = a->ac.ex.stream;
  if (cloud_ptr) LVBA_CUDA(cudaMemcpyAsync(cloud_ptr, a->ac.cloud_ptr.p, (size_t)(a->ac.n_windows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  if (xyz && a->ac.n_out) LVBA_CUDA(cudaMemcpyAsync(xyz, a->ac.out.p, (size_t)a->ac.n_out * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  LVBA_CUDA(cudaStreamSynchronize(s));
  if (ms_device) *ms_device = a->ms_device;
  return LVBA_OK;
} LVBA_ABI_END("lvba_anchor_clouds_export")

int lvba_anchor_clouds_destroy(lvba_anchor_clouds* a) LVBA_ABI_BEGIN {
  if (!a) return LVBA_OK;
  cudaSetDevice(a->device);
  delete a;
  return LVBA_OK;
} LVBA_ABI_END("lvba_anchor_clouds_destroy")

}  // extern "C"
