// anchor_pipeline.h — anchor clouds of runWindowBA (boundary B6): the tail of the window loop, src/lvba_system.cpp:284-301
//   for every scan j of a window:  tmp = *pl_win[j];  pl_transform(tmp, rel_j);  *merged += tmp;      (include/BALM/tools.hpp:385-395)
//   down_sampling_voxel2(*merged, anchor_leaf);                                                        (include/BALM/tools.hpp:301-359)
// pl_transform stores rel.R * p + rel.p back into the float members; down_sampling_voxel2 keeps, per voxel of edge `leaf`, the
// ORIGINAL point closest to the voxel centre (squared distance in double, strict '<': the first of equally close points in
// cloud order wins) and returns them in unordered_map order.  Here: transform + key + distance per point, one stable radix
// sort by (window, voxel key), a sequential arg-min per voxel run (stable order = cloud order, so the same point wins), and the
// survivors come out ordered by (window, voxel key).  Same Exec-policy scheme as voxel_pipeline.h.
#pragma once
#include "voxel_pipeline.h"

namespace lvba {
namespace anchor {

using vox::add_;
using vox::KeyPacking;
using vox::mul_;

struct AnchorPointF {        // one item per point
  const float* xyz; const int64_t* scan_ptr; const double* rel; const int32_t* win_ptr; int S; int n_windows; double leaf;
  float* pt; double* d2; int32_t* win_of; int32_t* kx; int32_t* ky; int32_t* kz; int32_t* bad;
  LVBA_HD void operator()(int64_t i) const {
    int lo = 0, hi = S;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (scan_ptr[mid] <= i) lo = mid; else hi = mid; }
    const int s = lo;
    lo = 0; hi = n_windows;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (win_ptr[mid] <= s) lo = mid; else hi = mid; }
    win_of[i] = lo;
    double w[3];
    vox::world_point(rel + 12 * (int64_t)s, xyz + 3 * i, w);                  // pvec = xx.R * pvec + xx.p   (tools.hpp:389-390)
    const float p[3] = {(float)w[0], (float)w[1], (float)w[2]};              // ap.x = pvec[0]  (:391-393)
    int64_t k[3] = {0, 0, 0};
    bool ok = true;
    double dd = 0.0;
    for (int a = 0; a < 3; ++a) {
      pt[3 * i + a] = p[a];
      ok = vox::root_key_axis((double)p[a], leaf, &k[a]) && ok;              // :317-327
      const double c = mul_(add_((double)k[a], 0.5), leaf);                   // voxel centre :332-334
      const double d = vox::sub_((double)p[a], c);
      dd = add_(dd, mul_(d, d));                                              // d2 = dx*dx + dy*dy + dz*dz  (:339)
    }
    if (!ok) { *bad = 1; k[0] = k[1] = k[2] = 0; }
    d2[i] = dd;
    kx[i] = (int32_t)k[0]; ky[i] = (int32_t)k[1]; kz[i] = (int32_t)k[2];
  }
};
struct AnchorKeyF {
  const int32_t* kx; const int32_t* ky; const int32_t* kz; const int32_t* win_of; KeyPacking pk; uint64_t* key; uint32_t* idx;
  LVBA_HD void operator()(int64_t i) const {
    const int64_t k[3] = {kx[i], ky[i], kz[i]};
    key[i] = pk.with_window(pk.pack(k), (uint32_t)win_of[i]); idx[i] = (uint32_t)i;
  }
};
struct RunHeadF {            // over [0, n] (terminator 0)
  const uint64_t* key; int64_t n; uint32_t* flag;
  LVBA_HD void operator()(int64_t r) const { flag[r] = (r < n && (r == 0 || key[r] != key[r - 1])) ? 1u : 0u; }
};
struct RunStartF {           // over [0, n]
  const uint32_t* flag; const uint32_t* pos; int64_t n; uint32_t* start;
  LVBA_HD void operator()(int64_t r) const { if (r == n) start[pos[n]] = (uint32_t)n; else if (flag[r]) start[pos[r]] = (uint32_t)r; }
};
struct BestInVoxelF {        // one item per voxel run: `if (!best.inited || d2 < best.best_d2)` over the points in cloud order (:341-346)
  const uint32_t* start; const uint32_t* idx; const uint64_t* key; const double* d2; const float* pt; KeyPacking pk; float* out; int32_t* out_win;
  LVBA_HD void operator()(int64_t v) const {
    uint32_t best = idx[start[v]];
    double bd = d2[best];
    for (uint32_t r = start[v] + 1; r < start[v + 1]; ++r) { const uint32_t i = idx[r]; if (d2[i] < bd) { bd = d2[i]; best = i; } }
    for (int a = 0; a < 3; ++a) out[3 * v + a] = pt[3 * (int64_t)best + a];
    out_win[v] = (int32_t)pk.window_of(key[start[v]]);
  }
};
struct WindowStartF {        // over [0, n_windows]: first output point of every window (outputs are window-major)
  const int32_t* out_win; int64_t n_out; int64_t* cloud_ptr;
  LVBA_HD void operator()(int64_t w) const {
    int64_t lo = 0, hi = n_out;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (out_win[mid] < (int32_t)w) lo = mid + 1; else hi = mid; }
    cloud_ptr[w] = lo;
  }
};
struct PassThroughF {        // leaf < 0.001: no down-sampling (:303), the transformed points in cloud order
  const float* pt; float* out;
  LVBA_HD void operator()(int64_t i) const { out[i] = pt[i]; }
};
struct ScanWindowPtrF {      // cloud_ptr[w] = scan_ptr[win_ptr[w]] for the pass-through case
  const int64_t* scan_ptr; const int32_t* win_ptr; int64_t* cloud_ptr;
  LVBA_HD void operator()(int64_t w) const { cloud_ptr[w] = scan_ptr[win_ptr[w]]; }
};

template <class Exec>
struct AnchorClouds {
  Exec ex;
  int n_windows = 0;
  int64_t n_out = 0;
  typename Exec::template Buf<float> out;            // [n_out * 3] anchor-frame points, window-major
  typename Exec::template Buf<int64_t> cloud_ptr;    // [n_windows + 1]
  const char* error = "";

  // xyz [N*3], scan_ptr [S+1], rel [S*12] (pose of every scan in its anchor frame), win_ptr [n_windows+1] over scans
  int build(const float* xyz, const int64_t* scan_ptr, const double* rel, const int32_t* win_ptr, int S, int n_windows_, int64_t N, double leaf) {
    n_windows = n_windows_; n_out = 0;
    LVBA_VOX_TRY(cloud_ptr.alloc((size_t)n_windows + 1));
    LVBA_VOX_TRY(ex.fill_zero(cloud_ptr.p, (size_t)n_windows + 1));
    if (N == 0) return ex.sync();
    typename Exec::template Buf<float> pt;
    typename Exec::template Buf<double> d2;
    typename Exec::template Buf<int32_t> win_of, kx, ky, kz, bad;
    LVBA_VOX_TRY(pt.alloc((size_t)N * 3)); LVBA_VOX_TRY(d2.alloc((size_t)N)); LVBA_VOX_TRY(win_of.alloc((size_t)N));
    LVBA_VOX_TRY(kx.alloc((size_t)N)); LVBA_VOX_TRY(ky.alloc((size_t)N)); LVBA_VOX_TRY(kz.alloc((size_t)N)); LVBA_VOX_TRY(bad.alloc(1));
    LVBA_VOX_TRY(ex.fill_zero(bad.p, 1));
    const bool sample = !(leaf < 0.001);
    LVBA_VOX_TRY(ex.for_each(N, AnchorPointF{xyz, scan_ptr, rel, win_ptr, S, n_windows, sample ? leaf : 1.0, pt.p, d2.p, win_of.p, kx.p, ky.p, kz.p, bad.p}));
    if (!sample) {
      n_out = N;
      LVBA_VOX_TRY(out.alloc((size_t)N * 3));
      LVBA_VOX_TRY(ex.for_each(N * 3, PassThroughF{pt.p, out.p}));
      LVBA_VOX_TRY(ex.for_each((int64_t)n_windows + 1, ScanWindowPtrF{scan_ptr, win_ptr, cloud_ptr.p}));
      return ex.sync();
    }
    int32_t h_bad = 0;
    LVBA_VOX_TRY(ex.fetch(&h_bad, bad.p, 1));
    if (h_bad) { error = "a transformed point is non-finite or more than 2^30 leaf voxels from the anchor"; return vox::kErrInvalid; }
    KeyPacking pk{{0, 0, 0}, {0, 0, 0}, 0, 0};
    const int32_t* axis[3] = {kx.p, ky.p, kz.p};
    for (int a = 0; a < 3; ++a) {
      int32_t mn = 0, mx = 0;
      LVBA_VOX_TRY(ex.min_max(axis[a], N, &mn, &mx));
      pk.mn[a] = mn;
      pk.bits[a] = vox::bit_length((uint64_t)((int64_t)mx - (int64_t)mn));
    }
    pk.root_bits = pk.bits[0] + pk.bits[1] + pk.bits[2];
    pk.win_bits = n_windows > 1 ? vox::bit_length((uint64_t)(n_windows - 1)) : 0;
    if (pk.key_bits() > 62) { error = "leaf voxel keys (and window index) span more than 62 bits"; return vox::kErrUnsupported; }
    typename Exec::template Buf<uint64_t> key, key_s;
    typename Exec::template Buf<uint32_t> idx, idx_s, flag, pos, start;
    LVBA_VOX_TRY(key.alloc((size_t)N)); LVBA_VOX_TRY(key_s.alloc((size_t)N)); LVBA_VOX_TRY(idx.alloc((size_t)N)); LVBA_VOX_TRY(idx_s.alloc((size_t)N));
    LVBA_VOX_TRY(ex.for_each(N, AnchorKeyF{kx.p, ky.p, kz.p, win_of.p, pk, key.p, idx.p}));
    LVBA_VOX_TRY(ex.sort_pairs(key.p, key_s.p, idx.p, idx_s.p, N, pk.key_bits() > 0 ? pk.key_bits() : 1));
    LVBA_VOX_TRY(flag.alloc((size_t)N + 1)); LVBA_VOX_TRY(pos.alloc((size_t)N + 1));
    LVBA_VOX_TRY(ex.for_each(N + 1, RunHeadF{key_s.p, N, flag.p}));
    LVBA_VOX_TRY(ex.exclusive_scan(flag.p, pos.p, N + 1));
    uint32_t nv = 0;
    LVBA_VOX_TRY(ex.fetch(&nv, pos.p + N, 1));
    n_out = nv;
    LVBA_VOX_TRY(start.alloc((size_t)nv + 1));
    LVBA_VOX_TRY(ex.for_each(N + 1, RunStartF{flag.p, pos.p, N, start.p}));
    typename Exec::template Buf<int32_t> out_win;
    LVBA_VOX_TRY(out.alloc((size_t)nv * 3)); LVBA_VOX_TRY(out_win.alloc((size_t)nv));
    LVBA_VOX_TRY(ex.for_each((int64_t)nv, BestInVoxelF{start.p, idx_s.p, key_s.p, d2.p, pt.p, pk, out.p, out_win.p}));
    LVBA_VOX_TRY(ex.for_each((int64_t)n_windows + 1, WindowStartF{out_win.p, (int64_t)nv, cloud_ptr.p}));
    return ex.sync();
  }
};

}  // namespace anchor
}  // namespace lvba
