// voxel_pipeline.h — the adaptive voxel map (SURVEY.md §8 boundary B3: cut_voxel -> recut -> tras_opt, plus the plane
// lookup of recompute_local_planes) as a sequence of data-parallel passes over flat arrays.
//
// The reference grows an unordered_map of octrees point by point (include/BALM/bavoxel.hpp:799-836, 420-464).  Here the
// octant path of a point is a pure function of its world position, so every point gets its whole key up front and the
// tree becomes three sorted arrays of nodes (one per layer):
//
//   1  point keys       world = R p + t, root key per axis, the two octants; min/max per axis   (the only pass over xyz besides the sums)
//   2  packed keys      key2 = (packed root key) << 6 | octant1 << 3 | octant2                  (17 B in, 8 B out per point)
//   3  per layer L      STABLE radix sort of the points by key2 >> 3(2-L): inside a node the points keep the caller's
//                       order (scan by scan, point by point), which is the order the reference pushes them in, so the
//                       sequential per-(node, pose) sums below reproduce its PointCluster sums term by term;
//                       run heads -> (node, pose) segments and nodes; segment sums; node plane test, which needs the
//                       parent's state from layer L-1 (a node only exists in the reference if its parent was split)
//   4  emission         plane nodes seen from >= 2 poses (push_voxel :45-54), ordered by (root key, path), as the CSR
//                       cluster layout lvba_lidar_lm takes
//   5  lookup           findCorrespondPoint :320-333 as binary searches over the per-layer node arrays
//
// The passes are written against an `Exec` policy (buffers, for_each, stable radix sort, scan, min/max): voxel_api.cuh
// supplies the CUDA one (grid-stride kernels + cub), tests/emu/voxel_emu.cpp a sequential host one so that the
// pipeline's logic is checked against oracle/voxel_oracle.py on machines without a GPU.  The product only ever
// instantiates the CUDA policy.
#pragma once
#include <stdint.h>

#include "voxel_math.h"

namespace lvba {
namespace vox {

enum NodeState : uint8_t {
  NS_NONE = 0,    // parent was not split: the reference never creates this node
  NS_MID = 1,     // fewer than min_ps points, or not a plane at layer_limit   (MID_NODE, :429-433, :446-452)
  NS_PLANE = 2,   // (:435-443)
  NS_SPLIT = 3    // not a plane, children exist                               (:453-456)
};

constexpr int kMaxLayers = 3;                 // layer_limit <= 2
constexpr int kErrOk = 0, kErrInvalid = -1, kErrUnsupported = -4;

// lower_bound over an ascending uint64 array; returns n when every element is < key
LVBA_HD int64_t lower_bound_u64(const uint64_t* a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
LVBA_HD int64_t find_u64(const uint64_t* a, int64_t n, uint64_t key) {
  const int64_t i = lower_bound_u64(a, n, key);
  return (i < n && a[i] == key) ? i : -1;
}

// How the three signed root-key axes are packed into one unsigned sort key (order-preserving: x major, then y, z — the
// order of VOXEL_LOC::operator<, tools.hpp:42-46).
struct KeyPacking {
  int32_t mn[3];
  int bits[3];
  int root_bits;
  int win_bits;    // > 0: a window index sits above the root key (one independent map per window, see VoxelMap::build)
  LVBA_HD int key_bits() const { return root_bits + win_bits; }
  LVBA_HD uint64_t with_window(uint64_t root, uint32_t w) const { return ((uint64_t)w << root_bits) | root; }
  LVBA_HD uint32_t window_of(uint64_t key) const { return (uint32_t)(key >> root_bits); }
  LVBA_HD uint64_t pack(const int64_t k[3]) const {
    return ((uint64_t)(k[0] - mn[0]) << (bits[1] + bits[2])) | ((uint64_t)(k[1] - mn[1]) << bits[2]) | (uint64_t)(k[2] - mn[2]);
  }
  LVBA_HD bool contains(const int64_t k[3]) const {
    for (int a = 0; a < 3; ++a)
      if (k[a] < mn[a] || (k[a] - mn[a]) >> bits[a]) return false;
    return true;
  }
  LVBA_HD void unpack(uint64_t r, int64_t k[3]) const {        // of the root part; window bits above it are ignored
    k[2] = (int64_t)(r & (((uint64_t)1 << bits[2]) - 1)) + mn[2];
    r >>= bits[2];
    k[1] = (int64_t)(r & (((uint64_t)1 << bits[1]) - 1)) + mn[1];
    r >>= bits[1];
    k[0] = (int64_t)(r & (((uint64_t)1 << bits[0]) - 1)) + mn[0];
  }
};

// Read-only view of one layer's node table (device pointers under the CUDA policy).
struct LayerView {
  int64_t n_nodes;
  const uint64_t* node_key;   // ascending; layer L key = packed root << 3L | octants
  const uint32_t* node_seg;   // [n_nodes + 1] first (node, pose) segment
  const uint8_t* state;
  const double* centre;       // [n_nodes * 3] merged world-frame centroid (judge_eigen's `center`)
  const double* direct;       // [n_nodes * 3] eigenvector of the smallest eigenvalue (`direct`)
  const double* eig;          // [n_nodes * 3] eigenvalues, ascending (`value_vector`)
  const int32_t* seg_pose;    // [n_segs]
  const double* seg_cluster;  // [n_segs * 10] body-frame PointCluster of that (node, pose)
};

// ================================================================ pass functors
struct PointKeysF {          // pass 1: everything that needs the point itself — its pose, root key and the two octants
  const float* xyz; const int64_t* scan_ptr; const double* poses; int W; double voxel_size;
  int32_t* pose_of; int32_t* kx; int32_t* ky; int32_t* kz; uint8_t* oct; int32_t* bad;
  LVBA_HD void operator()(int64_t i) const {
    int lo = 0, hi = W;                                   // last j with scan_ptr[j] <= i  (empty scans are skipped)
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (scan_ptr[mid] <= i) lo = mid; else hi = mid; }
    pose_of[i] = lo;
    double w[3];
    world_point(poses + 12 * (int64_t)lo, xyz + 3 * i, w);
    int64_t k[3] = {0, 0, 0};
    bool ok = true;
    for (int a = 0; a < 3; ++a) ok = root_key_axis(w[a], voxel_size, &k[a]) && ok;
    if (!ok) { *bad = 1; k[0] = k[1] = k[2] = 0; }
    kx[i] = (int32_t)k[0]; ky[i] = (int32_t)k[1]; kz[i] = (int32_t)k[2];
    float c0[3], c1[3];                                   // octant path: a function of the world point and the root key only
    int b[3];
    for (int a = 0; a < 3; ++a) c0[a] = root_centre_axis(k[a], voxel_size);
    const int o1 = octant(w, c0, b);
    child_centre(c0, b, root_quater(voxel_size), c1);
    const int o2 = octant(w, c1, b);
    oct[i] = (uint8_t)(o1 << 3 | o2);
  }
};

struct PackKeysF {           // pass 2 (after the key range is known): no access to the points any more
  const int32_t* pose_of; const int32_t* kx; const int32_t* ky; const int32_t* kz; const uint8_t* oct;
  KeyPacking pk; uint64_t* key2; const int32_t* win_ptr; int n_windows;
  LVBA_HD void operator()(int64_t i) const {
    uint32_t win = 0;
    if (n_windows > 0) {                                  // window of the scan: last w with win_ptr[w] <= scan
      int lo = 0, hi = n_windows;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (win_ptr[mid] <= pose_of[i]) lo = mid; else hi = mid; }
      win = (uint32_t)lo;
    }
    const int64_t k[3] = {kx[i], ky[i], kz[i]};
    key2[i] = (pk.with_window(pk.pack(k), win) << 6) | (uint64_t)oct[i];
  }
};

struct LayerKeysF {          // pass 3a; `sub` lists the points that take part in this layer (null: all of them)
  const uint64_t* key2; const uint32_t* sub; int shift; uint64_t* key; uint32_t* idx;
  LVBA_HD void operator()(int64_t r) const { const uint32_t i = sub ? sub[r] : (uint32_t)r; key[r] = key2[i] >> shift; idx[r] = i; }
};

struct HeadFlagsF {          // pass 3b, over [0, N] — element N is the terminator (flag 0)
  const uint64_t* key; const uint32_t* idx; const int32_t* pose_of; int64_t N; uint32_t* seg_flag; uint32_t* node_flag;
  LVBA_HD void operator()(int64_t r) const {
    uint32_t nf = 0, sf = 0;
    if (r < N) {
      nf = (r == 0 || key[r] != key[r - 1]) ? 1u : 0u;
      sf = (nf || pose_of[idx[r]] != pose_of[idx[r - 1]]) ? 1u : 0u;
    }
    seg_flag[r] = sf; node_flag[r] = nf;
  }
};

struct ScatterHeadsF {       // pass 3c, over [0, N]
  const uint64_t* key; const uint32_t* idx; const int32_t* pose_of; int64_t N;
  const uint32_t* seg_flag; const uint32_t* node_flag; const uint32_t* seg_pos; const uint32_t* node_pos;
  uint32_t* seg_start; int32_t* seg_pose; uint64_t* node_key; uint32_t* node_seg;
  LVBA_HD void operator()(int64_t r) const {
    if (r == N) { seg_start[seg_pos[N]] = (uint32_t)N; node_seg[node_pos[N]] = seg_pos[N]; return; }
    if (seg_flag[r]) { seg_start[seg_pos[r]] = (uint32_t)r; seg_pose[seg_pos[r]] = pose_of[idx[r]]; }
    if (node_flag[r]) { node_key[node_pos[r]] = key[r]; node_seg[node_pos[r]] = seg_pos[r]; }
  }
};

struct SplitFlagF {          // pass 3f, over [0, n] (terminator 0): does the point's node split?  (its children exist then)
  const uint32_t* node_flag; const uint32_t* node_pos; const uint8_t* state; int64_t n; uint32_t* flag;
  LVBA_HD void operator()(int64_t r) const { flag[r] = (r < n && state[node_pos[r] + node_flag[r] - 1u] == NS_SPLIT) ? 1u : 0u; }
};
struct SplitCompactF {       // pass 3g: the points of split nodes, in this layer's sorted order (equal next-layer keys keep the caller's order)
  const uint32_t* flag; const uint32_t* pos; const uint32_t* idx; uint32_t* next;
  LVBA_HD void operator()(int64_t r) const { if (flag[r]) next[pos[r]] = idx[r]; }
};

struct SegmentSumF {         // pass 3d: PointCluster::push over the segment's points in the caller's order
  const float* xyz; const uint32_t* idx; const uint32_t* seg_start; double* seg_cluster;
  LVBA_HD void operator()(int64_t s) const {
    double c[10];
    cluster_zero(c);
    for (uint32_t r = seg_start[s]; r < seg_start[s + 1]; ++r) cluster_push(c, xyz + 3 * (int64_t)idx[r]);
    for (int k = 0; k < 10; ++k) seg_cluster[10 * s + k] = c[k];
  }
};

struct NodeTestF {           // pass 3e: recut :420-464 for one node
  int layer; VoxParams prm; const double* poses;
  int64_t n_parent; const uint64_t* parent_key; const uint8_t* parent_state;
  const uint64_t* node_key; const uint32_t* node_seg; const int32_t* seg_pose; const double* seg_cluster;
  uint8_t* state; double* centre; double* direct; double* eig;
  LVBA_HD void operator()(int64_t m) const {
    for (int k = 0; k < 3; ++k) { centre[3 * m + k] = 0.0; direct[3 * m + k] = 0.0; eig[3 * m + k] = 0.0; }
    if (layer > 0) {
      const int64_t p = find_u64(parent_key, n_parent, node_key[m] >> 3);
      if (p < 0 || parent_state[p] != NS_SPLIT) { state[m] = NS_NONE; return; }
    }
    const uint32_t a = node_seg[m], b = node_seg[m + 1];
    double n_points = 0.0;
    for (uint32_t s = a; s < b; ++s) n_points += seg_cluster[10 * (int64_t)s + 9];
    if (n_points < (double)prm.min_points) { state[m] = NS_MID; return; }
    double Pm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, vm[3] = {0, 0, 0}, Nm = 0.0;
    for (uint32_t s = a; s < b; ++s)                      // ascending pose index, as judge_eigen's loop over the window
      cluster_transform_add(seg_cluster + 10 * (int64_t)s, poses + 12 * (int64_t)seg_pose[s], Pm, vm, &Nm);
    double c[3], d[3], lam[3];
    const bool plane = plane_test(Pm, vm, Nm, prm.eigen_ratio[layer], c, d, lam);
    for (int k = 0; k < 3; ++k) { centre[3 * m + k] = c[k]; direct[3 * m + k] = d[k]; eig[3 * m + k] = lam[k]; }
    state[m] = plane ? NS_PLANE : (layer == prm.layer_limit ? NS_MID : NS_SPLIT);
  }
};

struct EmitFlagF {           // pass 4a, over [0, n_nodes] (terminator 0): tras_opt + push_voxel
  const uint8_t* state; const uint32_t* node_seg; int64_t n_nodes; uint32_t* flag;
  LVBA_HD void operator()(int64_t m) const {
    flag[m] = (m < n_nodes && state[m] == NS_PLANE && node_seg[m + 1] - node_seg[m] >= 2u) ? 1u : 0u;
  }
};
struct EmitScatterF {        // pass 4b
  const uint32_t* flag; const uint32_t* pos; const uint64_t* node_key; int layer; int64_t base;
  uint64_t* ext_key; uint32_t* ref;
  LVBA_HD void operator()(int64_t m) const {
    if (!flag[m]) return;
    const int64_t q = base + pos[m];
    ext_key[q] = node_key[m] << (3 * (2 - layer));
    ref[q] = ((uint32_t)layer << 30) | (uint32_t)m;
  }
};
struct VoxelCountF {         // pass 4c, over [0, V] (terminator 0)
  const uint32_t* ref; int64_t V; LayerView lay[kMaxLayers]; int64_t* count;
  LVBA_HD void operator()(int64_t v) const {
    if (v == V) { count[v] = 0; return; }
    const LayerView& L = lay[ref[v] >> 30];
    const uint32_t m = ref[v] & 0x3fffffffu;
    count[v] = (int64_t)(L.node_seg[m + 1] - L.node_seg[m]);
  }
};
struct VoxelFillF {          // pass 4d: per plane voxel — key, path, centre, normal, eigenvalues
  const uint32_t* ref; const uint64_t* ext_key; LayerView lay[kMaxLayers]; KeyPacking pk;
  int64_t* root_key; int8_t* path; double* centre; double* direct; double* eig; int32_t* window;
  LVBA_HD void operator()(int64_t v) const {
    window[v] = (int32_t)pk.window_of(ext_key[v] >> 6);
    const int layer = (int)(ref[v] >> 30);
    const LayerView& L = lay[layer];
    const uint32_t m = ref[v] & 0x3fffffffu;
    int64_t k3[3];
    pk.unpack(ext_key[v] >> 6, k3);
    for (int k = 0; k < 3; ++k) {
      root_key[3 * v + k] = k3[k];
      centre[3 * v + k] = L.centre[3 * (int64_t)m + k];
      direct[3 * v + k] = L.direct[3 * (int64_t)m + k];
      eig[3 * v + k] = L.eig[3 * (int64_t)m + k];
    }
    path[3 * v] = (int8_t)layer;
    path[3 * v + 1] = layer >= 1 ? (int8_t)((ext_key[v] >> 3) & 7) : (int8_t)-1;
    path[3 * v + 2] = layer >= 2 ? (int8_t)(ext_key[v] & 7) : (int8_t)-1;
  }
};
// pass 4e: one item per (voxel, pose) SLOT of the output CSR — pose index and the 10-double cluster record of the slot.
// (One item per voxel copying its whole segment — a street voxel is seen from ~50-80 poses — ran as 27 CTAs waiting on one
// dependent load after the other: 445 us of the 1.27 ms map build at 5 M points, ncu long_scoreboard 98 %,
// profiles/r02_setup_ncu_summary.md.)  The voxel of slot q is the last v with vox_ptr[v] <= q.
struct SlotFillF {
  const uint32_t* ref; LayerView lay[kMaxLayers]; const int64_t* vox_ptr; int64_t V; int32_t* pose_idx; double* clusters;
  LVBA_HD void operator()(int64_t q) const {
    int64_t lo = 0, hi = V;                                 // vox_ptr[lo] <= q < vox_ptr[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (vox_ptr[mid] <= q) lo = mid; else hi = mid;
    }
    const LayerView& L = lay[ref[lo] >> 30];
    const uint32_t m = ref[lo] & 0x3fffffffu;
    const int64_t s = (int64_t)L.node_seg[m] + (q - vox_ptr[lo]);
    pose_idx[q] = L.seg_pose[s];
    for (int k = 0; k < 10; ++k) clusters[10 * q + k] = L.seg_cluster[10 * s + k];
  }
};

// pass 5 — src/lvba_system.cpp:1529-1566 (recompute_local_planes) with OCTO_TREE_NODE::findCorrespondPoint (:320-333):
// plane_nd = (n, d) with n = direct / |direct| and d = -n . center of the PLANE node the point falls in, zeros otherwise.
struct PlaneLookupF {
  const double* X; double voxel_size; int layer_limit; int n_layers; KeyPacking pk; LayerView lay[kMaxLayers]; double* plane_nd;
  LVBA_HD void operator()(int64_t i) const {
    double* out = plane_nd + 4 * i;
    out[0] = out[1] = out[2] = out[3] = 0.0;
    const double* x = X + 3 * i;
    if (!(fabs(x[0]) <= 1.79769313486231570e308 && fabs(x[1]) <= 1.79769313486231570e308 && fabs(x[2]) <= 1.79769313486231570e308)) return;
    int64_t k[3];
    for (int a = 0; a < 3; ++a) {                          // :1540-1543 (float arithmetic; same value as cut_voxel's)
      float loc = (float)(x[a] / voxel_size);
      if (loc < 0) loc = fadd_(loc, -1.0f);
      if (!(loc > -(float)kKeyLimit && loc < (float)kKeyLimit)) return;
      k[a] = (int64_t)loc;
    }
    if (n_layers < 1 || !pk.contains(k)) return;
    uint64_t key = pk.pack(k);
    int64_t m = find_u64(lay[0].node_key, lay[0].n_nodes, key);
    if (m < 0) return;                                     // surf_map.find(key) == end()
    float c[3];
    for (int a = 0; a < 3; ++a) c[a] = root_centre_axis(k[a], voxel_size);
    float quater = root_quater(voxel_size);
    int layer = 0;
    while (lay[layer].state[m] == NS_SPLIT && layer < layer_limit && layer + 1 < n_layers) {
      int b[3];
      const int o = octant(x, c, b);
      const uint64_t child = key << 3 | (uint64_t)o;
      const int64_t mc = find_u64(lay[layer + 1].node_key, lay[layer + 1].n_nodes, child);
      if (mc < 0) break;                                   // leaves[leafnum] == nullptr -> this
      float cc[3];
      child_centre(c, b, quater, cc);
      for (int a = 0; a < 3; ++a) c[a] = cc[a];
      quater = quater / 2.0f;
      key = child; m = mc; ++layer;
    }
    const LayerView& L = lay[layer];
    if (L.state[m] != NS_PLANE) return;
    const double* d = L.direct + 3 * m;
    const double* ce = L.centre + 3 * m;
    const double nrm = sqrt(dot3_(d, d));
    if (!(nrm >= 1e-6) || !(fabs(ce[0]) <= 1.79769313486231570e308 && fabs(ce[1]) <= 1.79769313486231570e308 && fabs(ce[2]) <= 1.79769313486231570e308)) return;
    const double n[3] = {d[0] / nrm, d[1] / nrm, d[2] / nrm};
    out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
    out[3] = -dot3_(n, ce);
  }
};

// ================================================================ the map
template <class Exec>
struct Layer {
  int64_t n_nodes = 0, n_segs = 0;
  typename Exec::template Buf<uint64_t> node_key;
  typename Exec::template Buf<uint32_t> node_seg;
  typename Exec::template Buf<uint8_t> state;
  typename Exec::template Buf<double> centre, direct, eig;
  typename Exec::template Buf<int32_t> seg_pose;
  typename Exec::template Buf<double> seg_cluster;
  LayerView view() const {
    return LayerView{n_nodes, node_key.p, node_seg.p, state.p, centre.p, direct.p, eig.p, seg_pose.p, seg_cluster.p};
  }
};

inline int bit_length(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

#define LVBA_VOX_TRY(call) do { int rc__ = (call); if (rc__ != 0) return rc__; } while (0)

template <class Exec>
struct VoxelMap {
  Exec ex;
  VoxParams prm{};
  int W = 0;
  int64_t N = 0;
  KeyPacking pk{};
  int n_layers = 0;
  Layer<Exec> layer[kMaxLayers];
  // emitted plane voxels, ordered by (root key, path)
  int64_t V = 0, nnz = 0;
  typename Exec::template Buf<int64_t> vox_ptr;
  typename Exec::template Buf<int32_t> vox_pose;
  typename Exec::template Buf<double> vox_cluster;
  typename Exec::template Buf<int64_t> vox_root;
  typename Exec::template Buf<int8_t> vox_path;
  typename Exec::template Buf<double> vox_centre, vox_direct, vox_eig;
  typename Exec::template Buf<int32_t> vox_window;   // [V] window of every voxel (all 0 for a single map)
  int n_windows = 0;                                  // 0: one map over all scans
  const char* error = "";

  // xyz [N*3], scan_ptr [W+1], poses [W*12]: pointers the Exec's passes can dereference (device memory under CUDA).
  // win_ptr [n_windows+1] (Exec-dereferenceable) splits the scans into consecutive windows, each of which gets its OWN map
  // (the surf_map that runWindowBA builds per window, src/lvba_system.cpp:247-258): the window index becomes the top bits of
  // every key, so voxels never merge across windows, pose indices stay global, and the emitted CSR is the input of
  // lvba_lidar_lm_batch as it stands.  n_windows = 0: one map.
  int build(const float* xyz, const int64_t* scan_ptr, const double* poses, int W_, int64_t N_, const VoxParams& prm_,
            const int32_t* win_ptr = nullptr, int n_windows_ = 0) {
    W = W_; N = N_; prm = prm_; n_windows = n_windows_;
    n_layers = 0; V = 0; nnz = 0;
    pk = KeyPacking{{0, 0, 0}, {0, 0, 0}, 0, 0};
    pk.win_bits = n_windows > 1 ? bit_length((uint64_t)(n_windows - 1)) : 0;
    if (N > 0) {
      typename Exec::template Buf<int32_t> pose_of;
      typename Exec::template Buf<uint64_t> key2;
      LVBA_VOX_TRY(pose_of.alloc((size_t)N));
      LVBA_VOX_TRY(key2.alloc((size_t)N));
      {
        typename Exec::template Buf<int32_t> kx, ky, kz, bad;
        typename Exec::template Buf<uint8_t> oct;
        LVBA_VOX_TRY(kx.alloc((size_t)N)); LVBA_VOX_TRY(ky.alloc((size_t)N)); LVBA_VOX_TRY(kz.alloc((size_t)N));
        LVBA_VOX_TRY(oct.alloc((size_t)N));
        LVBA_VOX_TRY(bad.alloc(1));
        LVBA_VOX_TRY(ex.fill_zero(bad.p, 1));
        LVBA_VOX_TRY(ex.for_each(N, PointKeysF{xyz, scan_ptr, poses, W, prm.voxel_size, pose_of.p, kx.p, ky.p, kz.p, oct.p, bad.p}));
        int32_t h_bad = 0;
        LVBA_VOX_TRY(ex.fetch(&h_bad, bad.p, 1));
        if (h_bad) { error = "a point is non-finite or more than 2^30 root voxels from the origin"; return kErrInvalid; }
        const int32_t* axis[3] = {kx.p, ky.p, kz.p};
        for (int a = 0; a < 3; ++a) {
          int32_t mn = 0, mx = 0;
          LVBA_VOX_TRY(ex.min_max(axis[a], N, &mn, &mx));
          pk.mn[a] = mn;
          pk.bits[a] = bit_length((uint64_t)((int64_t)mx - (int64_t)mn));
        }
        pk.root_bits = pk.bits[0] + pk.bits[1] + pk.bits[2];
        if (pk.key_bits() + 6 > 62) { error = "root voxel keys (and window index) span more than 56 bits"; return kErrUnsupported; }
        LVBA_VOX_TRY(ex.for_each(N, PackKeysF{pose_of.p, kx.p, ky.p, kz.p, oct.p, pk, key2.p, win_ptr, n_windows}));
      }
      // Only the points of nodes that split take part in the next layer (the reference hands exactly those to cut_func,
      // :453-456); layer 0 takes them all.
      typename Exec::template Buf<uint32_t> sub, next;
      int64_t n_sub = N;
      for (int L = 0; L <= prm.layer_limit && n_sub > 0; ++L) {
        int64_t n_next = 0;
        LVBA_VOX_TRY(build_layer(L, xyz, poses, pose_of.p, key2.p, L == 0 ? nullptr : sub.p, n_sub, L < prm.layer_limit ? &next : nullptr, &n_next));
        n_layers = L + 1;
        sub.swap(next);
        n_sub = n_next;
      }
    }
    return emit();
  }

  // sub [n] (null = all N points, in the caller's order): the points of this layer.  next / n_next (optional): the points of
  // the nodes that split here, for the next layer.
  int build_layer(int L, const float* xyz, const double* poses, const int32_t* pose_of, const uint64_t* key2, const uint32_t* sub,
                  int64_t N, typename Exec::template Buf<uint32_t>* next, int64_t* n_next) {
    Layer<Exec>& Y = layer[L];
    typename Exec::template Buf<uint64_t> kin, kout;
    typename Exec::template Buf<uint32_t> vin, idx, seg_flag, node_flag, seg_pos, node_pos, seg_start;
    LVBA_VOX_TRY(kin.alloc((size_t)N)); LVBA_VOX_TRY(kout.alloc((size_t)N));
    LVBA_VOX_TRY(vin.alloc((size_t)N)); LVBA_VOX_TRY(idx.alloc((size_t)N));
    LVBA_VOX_TRY(ex.for_each(N, LayerKeysF{key2, sub, 3 * (2 - L), kin.p, vin.p}));
    const int end_bit = pk.key_bits() + 3 * L;
    LVBA_VOX_TRY(ex.sort_pairs(kin.p, kout.p, vin.p, idx.p, N, end_bit > 0 ? end_bit : 1));
    LVBA_VOX_TRY(seg_flag.alloc((size_t)N + 1)); LVBA_VOX_TRY(node_flag.alloc((size_t)N + 1));
    LVBA_VOX_TRY(seg_pos.alloc((size_t)N + 1)); LVBA_VOX_TRY(node_pos.alloc((size_t)N + 1));
    LVBA_VOX_TRY(ex.for_each(N + 1, HeadFlagsF{kout.p, idx.p, pose_of, N, seg_flag.p, node_flag.p}));
    LVBA_VOX_TRY(ex.exclusive_scan(seg_flag.p, seg_pos.p, N + 1));
    LVBA_VOX_TRY(ex.exclusive_scan(node_flag.p, node_pos.p, N + 1));
    uint32_t n_segs = 0, n_nodes = 0;
    LVBA_VOX_TRY(ex.fetch(&n_segs, seg_pos.p + N, 1));
    LVBA_VOX_TRY(ex.fetch(&n_nodes, node_pos.p + N, 1));
    Y.n_segs = n_segs; Y.n_nodes = n_nodes;
    if (n_nodes >= (1u << 30)) { error = "more than 2^30 nodes in one layer"; return kErrUnsupported; }
    LVBA_VOX_TRY(seg_start.alloc((size_t)n_segs + 1));
    LVBA_VOX_TRY(Y.seg_pose.alloc(n_segs)); LVBA_VOX_TRY(Y.seg_cluster.alloc((size_t)n_segs * 10));
    LVBA_VOX_TRY(Y.node_key.alloc(n_nodes)); LVBA_VOX_TRY(Y.node_seg.alloc((size_t)n_nodes + 1));
    LVBA_VOX_TRY(Y.state.alloc(n_nodes));
    LVBA_VOX_TRY(Y.centre.alloc((size_t)n_nodes * 3)); LVBA_VOX_TRY(Y.direct.alloc((size_t)n_nodes * 3)); LVBA_VOX_TRY(Y.eig.alloc((size_t)n_nodes * 3));
    LVBA_VOX_TRY(ex.for_each(N + 1, ScatterHeadsF{kout.p, idx.p, pose_of, N, seg_flag.p, node_flag.p, seg_pos.p, node_pos.p,
                                                   seg_start.p, Y.seg_pose.p, Y.node_key.p, Y.node_seg.p}));
    LVBA_VOX_TRY(ex.for_each(n_segs, SegmentSumF{xyz, idx.p, seg_start.p, Y.seg_cluster.p}));
    const bool has_parent = L > 0;
    LVBA_VOX_TRY(ex.for_each(n_nodes, NodeTestF{L, prm, poses, has_parent ? layer[L - 1].n_nodes : 0,
                                                 has_parent ? layer[L - 1].node_key.p : nullptr,
                                                 has_parent ? layer[L - 1].state.p : nullptr,
                                                 Y.node_key.p, Y.node_seg.p, Y.seg_pose.p, Y.seg_cluster.p,
                                                 Y.state.p, Y.centre.p, Y.direct.p, Y.eig.p}));
    if (next) {                                            // seg_flag / seg_pos are free again: reuse them for the compaction
      LVBA_VOX_TRY(ex.for_each(N + 1, SplitFlagF{node_flag.p, node_pos.p, Y.state.p, N, seg_flag.p}));
      LVBA_VOX_TRY(ex.exclusive_scan(seg_flag.p, seg_pos.p, N + 1));
      uint32_t n = 0;
      LVBA_VOX_TRY(ex.fetch(&n, seg_pos.p + N, 1));
      *n_next = n;
      LVBA_VOX_TRY(next->alloc(n));
      LVBA_VOX_TRY(ex.for_each(N, SplitCompactF{seg_flag.p, seg_pos.p, idx.p, next->p}));
    }
    return 0;
  }

  int emit() {
    typename Exec::template Buf<uint32_t> flag[kMaxLayers], pos[kMaxLayers];
    int64_t base[kMaxLayers + 1] = {0, 0, 0, 0};
    for (int L = 0; L < n_layers; ++L) {
      const int64_t M = layer[L].n_nodes;
      LVBA_VOX_TRY(flag[L].alloc((size_t)M + 1)); LVBA_VOX_TRY(pos[L].alloc((size_t)M + 1));
      LVBA_VOX_TRY(ex.for_each(M + 1, EmitFlagF{layer[L].state.p, layer[L].node_seg.p, M, flag[L].p}));
      LVBA_VOX_TRY(ex.exclusive_scan(flag[L].p, pos[L].p, M + 1));
      uint32_t e = 0;
      LVBA_VOX_TRY(ex.fetch(&e, pos[L].p + M, 1));
      base[L + 1] = base[L] + e;
    }
    V = base[n_layers];
    LVBA_VOX_TRY(vox_ptr.alloc((size_t)V + 1));
    typename Exec::template Buf<uint64_t> ext_in, ext;
    typename Exec::template Buf<uint32_t> ref_in, ref;
    LVBA_VOX_TRY(ext_in.alloc((size_t)V)); LVBA_VOX_TRY(ext.alloc((size_t)V));
    LVBA_VOX_TRY(ref_in.alloc((size_t)V)); LVBA_VOX_TRY(ref.alloc((size_t)V));
    for (int L = 0; L < n_layers; ++L)
      LVBA_VOX_TRY(ex.for_each(layer[L].n_nodes, EmitScatterF{flag[L].p, pos[L].p, layer[L].node_key.p, L, base[L], ext_in.p, ref_in.p}));
    if (V > 0) LVBA_VOX_TRY(ex.sort_pairs(ext_in.p, ext.p, ref_in.p, ref.p, V, pk.key_bits() + 6));
    typename Exec::template Buf<int64_t> count;
    LVBA_VOX_TRY(count.alloc((size_t)V + 1));
    VoxelCountF cf{ref.p, V, {}, count.p};
    for (int L = 0; L < kMaxLayers; ++L) cf.lay[L] = L < n_layers ? layer[L].view() : LayerView{};
    LVBA_VOX_TRY(ex.for_each(V + 1, cf));
    LVBA_VOX_TRY(ex.exclusive_scan(count.p, vox_ptr.p, V + 1));
    LVBA_VOX_TRY(ex.fetch(&nnz, vox_ptr.p + V, 1));
    LVBA_VOX_TRY(vox_pose.alloc((size_t)nnz)); LVBA_VOX_TRY(vox_cluster.alloc((size_t)nnz * 10));
    LVBA_VOX_TRY(vox_root.alloc((size_t)V * 3)); LVBA_VOX_TRY(vox_path.alloc((size_t)V * 3));
    LVBA_VOX_TRY(vox_centre.alloc((size_t)V * 3)); LVBA_VOX_TRY(vox_direct.alloc((size_t)V * 3)); LVBA_VOX_TRY(vox_eig.alloc((size_t)V * 3));
    LVBA_VOX_TRY(vox_window.alloc((size_t)V));
    VoxelFillF ff{ref.p, ext.p, {}, pk, vox_root.p, vox_path.p, vox_centre.p, vox_direct.p, vox_eig.p, vox_window.p};
    SlotFillF sf{ref.p, {}, vox_ptr.p, V, vox_pose.p, vox_cluster.p};
    for (int L = 0; L < kMaxLayers; ++L) { ff.lay[L] = cf.lay[L]; sf.lay[L] = cf.lay[L]; }
    LVBA_VOX_TRY(ex.for_each(V, ff));
    LVBA_VOX_TRY(ex.for_each(nnz, sf));
    return ex.sync();
  }

  // X [n*3] -> plane_nd [n*4], pointers the Exec's passes can dereference
  int lookup(int64_t n, const double* X, double* plane_nd) {
    if (pk.win_bits > 0) { error = "plane lookup needs a single map, not a windowed one"; return kErrUnsupported; }
    PlaneLookupF f{X, prm.voxel_size, prm.layer_limit, n_layers, pk, {}, plane_nd};
    for (int L = 0; L < kMaxLayers; ++L) f.lay[L] = L < n_layers ? layer[L].view() : LayerView{};
    LVBA_VOX_TRY(ex.for_each(n, f));
    return ex.sync();
  }
};

}  // namespace vox
}  // namespace lvba
