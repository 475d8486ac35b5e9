// anchor_emu.cpp — TEST INFRASTRUCTURE: the anchor-cloud pipeline of global-lvba_b200/csrc/anchor_pipeline.h run with the
// sequential host policy, for tests/test_anchor_emu.py.  Never part of the product.
#include <cstring>

#include "../../global-lvba_b200/csrc/anchor_pipeline.h"
#include "host_exec.h"

using AC = lvba::anchor::AnchorClouds<HostExec>;

// memcpy with a null source is undefined even for zero bytes (empty maps have unallocated buffers)
static void copy_n_bytes(void* dst, const void* src, size_t n) { if (n) std::memcpy(dst, src, n); }

extern "C" {
int emu_anchor_create(int32_t n_windows, const int32_t* win_ptr, const int64_t* scan_ptr, const float* xyz, const double* rel, double leaf,
                      void** out, int64_t* n_out) {
  AC* a = new AC();
  const int S = win_ptr[n_windows];
  const int rc = a->build(xyz, scan_ptr, rel, win_ptr, S, n_windows, scan_ptr[S], leaf);
  if (rc != 0) { delete a; return rc; }
  *out = a; *n_out = a->n_out;
  return 0;
}
int emu_anchor_export(void* h, int64_t* cloud_ptr, float* xyz) {
  AC* a = (AC*)h;
  copy_n_bytes(cloud_ptr, a->cloud_ptr.p, (size_t)(a->n_windows + 1) * sizeof(int64_t));
  copy_n_bytes(xyz, a->out.p, (size_t)a->n_out * 3 * sizeof(float));
  return 0;
}
int emu_anchor_destroy(void* h) { delete (AC*)h; return 0; }
}
