// depth_emu.cpp — TEST INFRASTRUCTURE: the depth-rendering pipeline of global-lvba_b200/csrc/depth_pipeline.h run with the
// sequential host policy, compared with oracle/depth_oracle.py by tests/test_depth_emu.py.  Never part of the product.
#include "../../global-lvba_b200/csrc/depth_pipeline.h"
#include "host_exec.h"

using Grid = lvba::depth::DepthGrid<HostExec>;

extern "C" {

int emu_depth_grid_create(int32_t F, const int64_t* scan_ptr, const float* xyz, const double* poses, const double* frame_ts,
                          double voxel_size, void** out, int64_t* n_voxels, int64_t* n_pairs) {
  Grid* g = new Grid();
  const int rc = g->build(xyz, scan_ptr, poses, frame_ts, F, scan_ptr[F], voxel_size);
  if (rc != 0) { delete g; return rc; }
  *out = g; *n_voxels = g->n_voxels; *n_pairs = g->n_pairs;
  return 0;
}

int emu_depth_render(void* h, int64_t n_img, const double* cams, const double* image_ts, double half_window, const double* intr,
                     int32_t width, int32_t height, float* depth, int64_t* work) {
  Grid* g = (Grid*)h;
  const int rc = g->render(n_img, cams, image_ts, half_window, intr, width, height, depth);
  if (work) { work[0] = g->last_pairs; work[1] = g->last_chunks; }
  return rc;
}

int emu_depth_backproject(void* h, int64_t n_img, const float* depth, const double* cams, const double* intr, int32_t width, int32_t height,
                          const int64_t* kp_ptr, const float* kp_uv, double* Xw, uint8_t* valid) {
  return ((Grid*)h)->backproject(n_img, depth, cams, intr, width, height, kp_ptr, kp_ptr[n_img] - kp_ptr[0], kp_uv, Xw, valid);
}

int emu_depth_grid_destroy(void* h) { delete (Grid*)h; return 0; }

}  // extern "C"
