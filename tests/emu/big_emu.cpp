// big_emu.cpp — TEST INFRASTRUCTURE: the big-voxel passes of global-lvba_b200/csrc/lidar_big.h (path A for voxels seen from
// more poses than a batch CTA holds) run by plain loops over ALL voxels of a problem, so that tests/test_big_voxel_emu.py can
// compare residual, g and every Hessian block with oracle/lidar_oracle.py without a GPU.  Never part of the product.
#include <vector>

#include "../../global-lvba_b200/csrc/lidar_big.h"
#include "host_exec.h"

extern "C" double emu_big_accumulate(int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx, const double* clusters, const double* poses,
                                     const int* first, const long long* row_start, double* H, double* g, int residual_only) {
  using namespace lvba::big;
  std::vector<int64_t> pair_ptr((size_t)V + 1, 0);
  for (int64_t a = 0; a < V; ++a) { const int64_t K = vox_ptr[a + 1] - vox_ptr[a]; pair_ptr[a + 1] = pair_ptr[a] + K * (K - 1) / 2; }
  View bv{V, vox_ptr, pose_idx, clusters, pair_ptr.data(), first, row_start};
  std::vector<double> params((size_t)V * kParams), res((size_t)V), feat((size_t)vox_ptr[V] * kFeat);
  ParamsF pf{bv, poses, params.data(), res.data()};
  HostExec ex;
  ex.for_each(V, pf);
  double sum = 0.0;
  for (double r : res) sum += r;
  if (residual_only) return sum;
  SlotsF sf{bv, poses, params.data(), feat.data(), H, g};
  ex.for_each(vox_ptr[V], sf);
  PairsF qf{bv, params.data(), feat.data(), H};
  ex.for_each(pair_ptr[V], qf);
  return sum;
}
