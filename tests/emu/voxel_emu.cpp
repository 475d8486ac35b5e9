// voxel_emu.cpp — TEST INFRASTRUCTURE: runs the voxel-map pipeline of global-lvba_b200/csrc/voxel_pipeline.h with a
// sequential host execution policy, so that the logic every CUDA pass executes (keys, stable sort order, segment sums,
// node states, emission order, plane lookup) is checked against oracle/voxel_oracle.py on a machine without a GPU.
// Built by tests/test_voxel_emu.py with g++ -ffp-contract=off; never part of liblvba_b200.so — the product
// instantiates the pipeline with the CUDA policy only (voxel_api.cuh) and has no host path.
#include <cstring>

#include "../../global-lvba_b200/csrc/voxel_pipeline.h"
#include "host_exec.h"

namespace {

using Map = lvba::vox::VoxelMap<HostExec>;

}  // namespace

// memcpy with a null source is undefined even for zero bytes (empty maps have unallocated buffers)
static void copy_n_bytes(void* dst, const void* src, size_t n) { if (n) std::memcpy(dst, src, n); }

extern "C" {

int emu_voxel_map_create(int32_t W, const int64_t* scan_ptr, const float* xyz, const double* poses, double voxel_size,
                         const float* eigen_ratio, int32_t layer_limit, int32_t min_points, void** out) {
  Map* m = new Map();
  lvba::vox::VoxParams prm{voxel_size, {eigen_ratio[0], eigen_ratio[1], eigen_ratio[2], eigen_ratio[3]}, layer_limit, min_points};
  const int rc = m->build(xyz, scan_ptr, poses, W, scan_ptr[W], prm);
  if (rc != 0) { delete m; return rc; }
  *out = m;
  return 0;
}

int emu_voxel_map_create_windows(int32_t n_windows, const int32_t* win_ptr, const int64_t* scan_ptr, const float* xyz, const double* poses,
                                 double voxel_size, const float* eigen_ratio, int32_t layer_limit, int32_t min_points, void** out) {
  Map* m = new Map();
  lvba::vox::VoxParams prm{voxel_size, {eigen_ratio[0], eigen_ratio[1], eigen_ratio[2], eigen_ratio[3]}, layer_limit, min_points};
  const int W = win_ptr[n_windows];
  const int rc = m->build(xyz, scan_ptr, poses, W, scan_ptr[W], prm, win_ptr, n_windows);
  if (rc != 0) { delete m; return rc; }
  *out = m;
  return 0;
}

int emu_voxel_map_windows(void* h, int32_t* vox_window) {
  Map* m = (Map*)h;
  copy_n_bytes(vox_window, m->vox_window.p, (size_t)m->V * sizeof(int32_t));
  return 0;
}

int emu_voxel_map_sizes(void* h, int64_t* n_voxels, int64_t* nnz, int64_t* n_nodes) {
  Map* m = (Map*)h;
  *n_voxels = m->V; *nnz = m->nnz;
  for (int L = 0; L < 3; ++L) n_nodes[L] = L < m->n_layers ? m->layer[L].n_nodes : 0;
  return 0;
}

int emu_voxel_map_export(void* h, int64_t* vox_ptr, int32_t* pose_idx, double* clusters, int64_t* root_key, int8_t* path,
                         double* centre, double* normal, double* eigenvalues) {
  Map* m = (Map*)h;
  copy_n_bytes(vox_ptr, m->vox_ptr.p, (size_t)(m->V + 1) * sizeof(int64_t));
  copy_n_bytes(pose_idx, m->vox_pose.p, (size_t)m->nnz * sizeof(int32_t));
  copy_n_bytes(clusters, m->vox_cluster.p, (size_t)m->nnz * 10 * sizeof(double));
  copy_n_bytes(root_key, m->vox_root.p, (size_t)m->V * 3 * sizeof(int64_t));
  copy_n_bytes(path, m->vox_path.p, (size_t)m->V * 3);
  copy_n_bytes(centre, m->vox_centre.p, (size_t)m->V * 3 * sizeof(double));
  copy_n_bytes(normal, m->vox_direct.p, (size_t)m->V * 3 * sizeof(double));
  copy_n_bytes(eigenvalues, m->vox_eig.p, (size_t)m->V * 3 * sizeof(double));
  return 0;
}

int emu_voxel_map_lookup(void* h, int64_t n, const double* X, double* plane_nd) { return ((Map*)h)->lookup(n, X, plane_nd); }

int emu_voxel_map_destroy(void* h) { delete (Map*)h; return 0; }

}  // extern "C"
