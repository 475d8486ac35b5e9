// fuse_pipeline.h — track fusion: LvbaSystem::BuildTracksAndFuse3D (reference src/lvba_system.cpp:921-1263), the stage that
// turns pairwise keypoint matches + depth images into the 3-D landmarks and observation lists the visual BA (boundary B2)
// consumes.  SURVEY.md section 8(f) N3.
//
//   :930-954   adjacency of the match graph                      -> host (build_adjacency: same push_back order)
//   :967-988   connected components by BFS from every unvisited keypoint in (image, keypoint) order
//                                                                -> host (bfs_component: same queue discipline, same member order)
//   :989-1003  drop components with < obser_thr members / images; first observation of every image
//   :1016-1103 depth candidate: per-member back-projected points (boundary B4, lvba_depth_backproject), anchor = first valid
//              member, members within 0.12 m of it, first such member per image, mean, greedy view-angle filter, mean
//              reprojection error                                 |
//   :1106-1159 triangulation candidate: DLT over the first member of every image, greedy view-angle filter around the seed,
//              DLT over the survivors                             |> device, ONE item per component (FuseF)
//   :1161-1199 choice between the two candidates, finite / non-zero test  |
//   :1199      a component that fails is released (obs_to_track = -1) and is therefore tried AGAIN from its next member in
//              scan order as BFS seed (another member order: another anchor, other first observations)
//                                                                -> host loop over rounds (fuse_api.cuh)
//
// The reference iterates three std::unordered_map<int,int> (image id -> member) whose order is unspecified; the sums over
// them (mean of the depth points, A^T A of the DLT, mean reprojection error) depend on it in the last bits and the GREEDY
// view-angle filter depends on it outright.  Two orders are implemented (Params::map_order, lvba_fuse_opts::map_order):
//   0  ASCENDING image id — independent of any C++ library;
//   1  the order GNU libstdc++'s std::unordered_map<int,int> iterates in after the reference's reserve() calls and insertions
//      (stl_map_order below: published bucket-count table, identity hash, insert-at-front-of-list / front-of-bucket) — what a g++
//      build of the reference does, and the default of the ABI.  With it the stage reproduces the reference's own BuildTracksAndFuse3D track for track
//      (tests/test_ref_system_pin.py: the reference's source compiled where it lies vs this pipeline through the host policy).
// Everything else (BFS order, first-per-image rule, thresholds) is the reference's in both.
//
// Same Exec-policy scheme as voxel_pipeline.h / depth_pipeline.h / track_pipeline.h: the functor below runs as a grid-stride
// kernel on the device and as a plain loop in tests/emu/fuse_emu.cpp.
#pragma once
#include <algorithm>
#include <deque>
#include <utility>
#include <vector>

#include "track_pipeline.h"

namespace lvba {
namespace fuse {

using depth::finite_;

struct Params {
  int obser_thr;             // lvba_system.h:139 (3)
  double cos_min_view;       // cos(track_fusion/min_view_angle), src/lvba_system.cpp:129 (8 degrees)
  double reproj_thr;         // track_fusion/reproj_mean_thr, :130 (3 px)
  double depth_gate;         // :1050 (0.12 m)
  int map_order;             // 0: ascending image id, 1: GNU libstdc++ unordered_map order (see the header comment)
};

// ------------------------------------------------------------------------------------------------ container order (map_order = 1)
// std::__detail::_Prime_rehash_policy::_M_next_bkt of GNU libstdc++ (read out of GCC 13.3's library by calling it for every n): bucket count after
// reserve(n) with max_load_factor 1.  Table entries as the library publishes them (up to 2.0e9: far beyond any component).
inline const uint32_t* stl_prime_table(int* n) {
  static const uint32_t primes[] = {
      17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127, 137, 139, 149, 157, 167, 179, 193, 199, 211, 227, 241,
      257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541, 577, 619, 661, 709, 761, 823, 887, 953, 1031, 1109, 1193, 1289, 1381, 1493, 1613, 1741,
      1879, 2029, 2179, 2357, 2549, 2753, 2971, 3209, 3469, 3739, 4027, 4349, 4703, 5087, 5503, 5953, 6427, 6949, 7517, 8123, 8783, 9497, 10273, 11113,
      12011, 12983, 14033, 15173, 16411, 17749, 19183, 20753, 22447, 24281, 26267, 28411, 30727, 33223, 35933, 38873, 42043, 45481, 49201, 53201, 57557,
      62233, 67307, 72817, 78779, 85229, 92203, 99733, 107897, 116731, 126271, 136607, 147793, 159871, 172933, 187091, 202409, 218971, 236897, 256279,
      277261, 299951, 324503, 351061, 379787, 410857, 444487, 480881, 520241, 562841, 608903, 658753, 712697, 771049, 834181, 902483, 976369, 1056323,
      1142821, 1236397, 1337629, 1447153, 1565659, 1693859, 1832561, 1982627, 2144977, 2320627, 2510653, 2716249, 2938679, 3179303, 3439651, 3721303,
      4026031, 4355707, 4712381, 5098259, 5515729, 5967347, 6456007, 6984629, 7556579, 8175383, 8844859, 9569143, 10352717, 11200489, 12117689, 13109983,
      14183539, 15345007, 16601593, 17961079, 19431899, 21023161, 22744717, 24607243, 26622317, 28802401, 31160981, 33712729, 36473443, 39460231,
      42691603, 46187573, 49969847, 54061849, 58488943, 63278561, 68460391, 74066549, 80131819, 86693767, 93793069, 101473717, 109783337, 118773397,
      128499677, 139022417, 150406843, 162723577, 176048909, 190465427, 206062531, 222936881, 241193053, 260944219, 282312799, 305431229, 330442829,
      357502601, 386778277, 418451333, 452718089, 489790921, 529899637, 573292817, 620239453, 671030513, 725980837, 785430967, 849749479, 919334987,
      994618837, 1076067617, 1164186217, 1259520799, 1362662261, 1474249943, 1594975441, 1725587117, 1866894511, 2019773507};
  *n = (int)(sizeof(primes) / sizeof(primes[0]));
  return primes;
}
// `primes` / `n_primes`: the table above, in the memory space of the caller (host: stl_prime_table; device: a copy uploaded by run())
LVBA_HD uint32_t stl_bucket_count(uint32_t n, const uint32_t* primes, int n_primes) {
  const unsigned char fast[14] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13};
  if (n == 0) return 2;               // reserve(0) = rehash(0): the container asks for room for one more element, _M_next_bkt(1) = 2
  if (n < 14) return fast[n];
  int lo = 0, hi = n_primes;
  while (lo < hi) { const int mid = (lo + hi) / 2; if (primes[mid] < n) lo = mid + 1; else hi = mid; }      // first entry >= n
  return lo < n_primes ? primes[lo] : primes[n_primes - 1];
}
// Iteration order of the map after reserve(reserve_n) and the insertion of n DISTINCT non-negative keys key(0..n-1) in this order:
// out[0..n) = indices into the insertion sequence, front of the container first.  libstdc++'s _Hashtable keeps ONE singly linked list; a
// node whose bucket (key mod bucket count: std::hash<int> is the identity) is empty goes to the front of the list, a node whose bucket is
// occupied goes to the front of that bucket's run.  No rehash afterwards: the reference never inserts more keys than it reserved.
template <class KeyOf>
LVBA_HD void stl_map_order(int reserve_n, KeyOf key, int n, int32_t* out, const uint32_t* primes, int n_primes) {
  const uint32_t nb = stl_bucket_count((uint32_t)(reserve_n < 0 ? 0 : reserve_n), primes, n_primes);
  for (int i = 0; i < n; ++i) {
    const uint32_t b = (uint32_t)key(i) % nb;
    int pos = 0;
    bool found = false;
    for (int q = 0; q < i; ++q)
      if ((uint32_t)key(out[q]) % nb == b) { pos = q; found = true; break; }
    if (!found) pos = 0;
    for (int q = i; q > pos; --q) out[q] = out[q - 1];
    out[pos] = i;
  }
}

// ------------------------------------------------------------------------------------------------ host: match graph
struct Graph {
  std::vector<int64_t> kp_ptr;                 // [N+1]
  std::vector<int64_t> adj_ptr;                // [n_kp+1]
  std::vector<int64_t> adj;                    // neighbours as global keypoint ids, in the reference's push_back order
  std::vector<int32_t> img_of;                 // [n_kp]
};
// matches must come in the order the reference visits them: image pairs (i < j) by ascending i then j, the matches of a pair
// in their stored order (:937-953).  A match with an out-of-range keypoint is skipped (:944-947).
inline void build_adjacency(int32_t n_images, const int64_t* kp_ptr, int64_t n_matches, const int32_t* ma_img, const int32_t* ma_kp,
                            const int32_t* mb_img, const int32_t* mb_kp, Graph& G) {
  G.kp_ptr.assign(kp_ptr, kp_ptr + n_images + 1);
  const int64_t n_kp = kp_ptr[n_images];
  G.img_of.resize((size_t)n_kp);
  for (int32_t i = 0; i < n_images; ++i) for (int64_t g = kp_ptr[i]; g < kp_ptr[i + 1]; ++g) G.img_of[(size_t)g] = i;
  auto gid = [&](int32_t im, int32_t k) -> int64_t {
    if (im < 0 || im >= n_images || k < 0 || k >= kp_ptr[im + 1] - kp_ptr[im]) return -1;
    return kp_ptr[im] + k;
  };
  std::vector<int64_t> deg((size_t)n_kp + 1, 0);
  for (int64_t m = 0; m < n_matches; ++m) {
    const int64_t a = gid(ma_img[m], ma_kp[m]), b = gid(mb_img[m], mb_kp[m]);
    if (a < 0 || b < 0) continue;
    ++deg[(size_t)a + 1]; ++deg[(size_t)b + 1];
  }
  for (int64_t g = 0; g < n_kp; ++g) deg[(size_t)g + 1] += deg[(size_t)g];
  G.adj_ptr = deg;
  G.adj.assign((size_t)deg[(size_t)n_kp], 0);
  std::vector<int64_t> fill(G.adj_ptr.begin(), G.adj_ptr.end() - 1);
  for (int64_t m = 0; m < n_matches; ++m) {
    const int64_t a = gid(ma_img[m], ma_kp[m]), b = gid(mb_img[m], mb_kp[m]);
    if (a < 0 || b < 0) continue;
    G.adj[(size_t)fill[(size_t)a]++] = b;
    G.adj[(size_t)fill[(size_t)b]++] = a;
  }
}
// BFS from `seed` over the keypoints whose mark is 0 (:968-988): members in dequeue order; marks the members with `stamp`
inline void bfs_component(const Graph& G, int64_t seed, std::vector<int32_t>& mark, int32_t stamp, std::vector<int64_t>& members) {
  members.clear();
  std::deque<int64_t> q;
  q.push_back(seed);
  mark[(size_t)seed] = stamp;
  while (!q.empty()) {
    const int64_t cur = q.front(); q.pop_front();
    members.push_back(cur);
    for (int64_t e = G.adj_ptr[(size_t)cur]; e < G.adj_ptr[(size_t)cur + 1]; ++e) {
      const int64_t nb = G.adj[(size_t)e];
      if (mark[(size_t)nb] == 0) { mark[(size_t)nb] = stamp; q.push_back(nb); }
    }
  }
}

struct KeyOfPairs {                            // key functor of stl_map_order over (image, position) pairs (host side)
  const std::pair<int32_t, int32_t>* p;
  LVBA_HD int operator()(int i) const { return (int)p[i].first; }
};

// one round of candidates (CSR) for the device
struct Batch {
  std::vector<int64_t> ptr{0};                 // [n+1] members
  std::vector<int32_t> img, kp;                // members in BFS order
  std::vector<int32_t> rank;                   // member -> index of its image in the component's ascending image list
  std::vector<int64_t> uptr{0};                // [n+1] images
  std::vector<int32_t> upos;                   // per image (ascending id): position of its FIRST member in the component (:996-1000)
  std::vector<int32_t> uord;                   // per image slot: the rank visited k-th by `for (auto& kv : unique_id)` (identity when map_order = 0)
  int map_order = 0;
  void clear() { ptr.assign(1, 0); img.clear(); kp.clear(); rank.clear(); uptr.assign(1, 0); upos.clear(); uord.clear(); }
  int64_t size() const { return (int64_t)ptr.size() - 1; }
  // returns the number of distinct images
  int add(const Graph& G, const std::vector<int64_t>& members) {
    std::vector<std::pair<int32_t, int32_t>> first;                       // (image, first position)
    for (size_t t = 0; t < members.size(); ++t) first.emplace_back(G.img_of[(size_t)members[t]], (int32_t)t);
    std::stable_sort(first.begin(), first.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<std::pair<int32_t, int32_t>> uniq;
    for (const auto& f : first) if (uniq.empty() || uniq.back().first != f.first) uniq.push_back(f);
    for (size_t t = 0; t < members.size(); ++t) {
      const int64_t g = members[t];
      const int32_t im = G.img_of[(size_t)g];
      img.push_back(im); kp.push_back((int32_t)(g - G.kp_ptr[(size_t)im]));
      const auto it = std::lower_bound(uniq.begin(), uniq.end(), std::make_pair(im, (int32_t)-1));
      rank.push_back((int32_t)(it - uniq.begin()));
    }
    for (const auto& u : uniq) upos.push_back(u.second);
    if (map_order == 1) {
      // unique_id.reserve(component.size()); keys = images in the order their first member appears (:994-999)
      std::vector<std::pair<int32_t, int32_t>> by_pos(uniq.begin(), uniq.end());
      std::sort(by_pos.begin(), by_pos.end(), [](const auto& x, const auto& y) { return x.second < y.second; });
      std::vector<int32_t> ord(by_pos.size());
      int n_primes = 0;
      const uint32_t* primes = stl_prime_table(&n_primes);
      stl_map_order((int)members.size(), KeyOfPairs{by_pos.data()}, (int)by_pos.size(), ord.data(), primes, n_primes);
      for (int32_t i : ord) {
        const auto it = std::lower_bound(uniq.begin(), uniq.end(), std::make_pair(by_pos[(size_t)i].first, (int32_t)-1));
        uord.push_back((int32_t)(it - uniq.begin()));
      }
    } else {
      for (size_t u = 0; u < uniq.size(); ++u) uord.push_back((int32_t)u);
    }
    ptr.push_back((int64_t)img.size());
    uptr.push_back((int64_t)upos.size());
    return (int)uniq.size();
  }
};

// ------------------------------------------------------------------------------------------------ device: one component
struct View {
  int64_t n_comp;
  const int64_t* ptr; const int32_t* img; const int32_t* kp; const int32_t* rank;
  const int64_t* uptr; const int32_t* upos; const int32_t* uord;
  const int64_t* kp_ptr; const float* kp_uv;           // keypoints of all images
  const double* kp_Xw; const uint8_t* kp_valid;        // depth candidate of every keypoint (lvba_depth_backproject, :1020-1038)
  int n_cams; const double* cams; double intr[8];      // cams: [n][12] = Rcw row-major, tcw
  Params prm;
  // scratch
  int32_t* best;             // [images of the batch] first gated member per image (depth path)
  int32_t* sel;              // [members of the batch] positions of the currently selected members
  double* dirs;              // [images of the batch][3]
  int32_t* ins;              // [images of the batch] map_order = 1: image ranks in the order best_id received them
  int32_t* cand;             // [images of the batch] map_order = 1: members in the order `for (auto& kv : best_id)` visits them
  const uint32_t* primes; int n_primes;     // map_order = 1: libstdc++'s bucket-count table (stl_prime_table) in the executing memory space
  // out
  uint8_t* status;           // per component: 0 dropped, 1 depth candidate, 2 triangulation candidate
  double* Xw;                // [n][3]
  double* mean;              // mean reprojection error of the chosen candidate
  uint8_t* inlier;           // per member: 1 = in Track::inlier_indices
};

// camera centre direction test shared by both candidates (:1069-1095, :1124-1150): members `cand[0..nc)` in the caller's order,
// 3-D point of member position t given by `pt(t)`; survivors are written to sel[0..*ns)
template <class PointOf>
LVBA_HD int view_angle_filter(const View& v, int64_t a, const int32_t* cand, int nc, PointOf pt, double* dirs, int32_t* sel) {
  int ns = 0;
  for (int u = 0; u < nc; ++u) {
    const int t = cand[u];
    if (t < 0) continue;
    const int cam = v.img[a + t];
    if (cam < 0 || cam >= v.n_cams) continue;
    const double* P = v.cams + 12 * (int64_t)cam;
    double Cw[3];                                                      // -Rcw^T tcw
    for (int i = 0; i < 3; ++i) Cw[i] = -(P[i] * P[9] + P[3 + i] * P[10] + P[6 + i] * P[11]);
    double X[3];
    pt(t, X);
    double d[3] = {X[0] - Cw[0], X[1] - Cw[1], X[2] - Cw[2]};
    const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (nrm < 1e-6) continue;
    d[0] /= nrm; d[1] /= nrm; d[2] /= nrm;
    double min_dot = 1.0;
    for (int q = 0; q < ns; ++q) {
      const double dot = d[0] * dirs[3 * q] + d[1] * dirs[3 * q + 1] + d[2] * dirs[3 * q + 2];
      if (dot < min_dot) min_dot = dot;
    }
    if (ns == 0 || min_dot <= v.prm.cos_min_view) {
      dirs[3 * ns] = d[0]; dirs[3 * ns + 1] = d[1]; dirs[3 * ns + 2] = d[2];
      sel[ns++] = t;
    }
  }
  return ns;
}

// ComputeMeanReproj (:8-50) over the selected member positions
LVBA_HD bool mean_reproj_sel(const View& v, int64_t a, const int32_t* sel, int ns, const double* X, int min_count, double* mean, int* cnt_out) {
  double sum = 0.0;
  int cnt = 0;
  for (int q = 0; q < ns; ++q) {
    const int t = sel[q];
    const int cam = v.img[a + t];
    if (cam < 0 || cam >= v.n_cams) continue;
    const int64_t g = v.kp_ptr[cam] + v.kp[a + t];
    double uh, vh;
    if (!track::project_world(v.cams + 12 * (int64_t)cam, v.intr, X, &uh, &vh)) continue;
    const double du = uh - (double)v.kp_uv[2 * g], dv = vh - (double)v.kp_uv[2 * g + 1];
    sum += sqrt(du * du + dv * dv);
    ++cnt;
  }
  *cnt_out = cnt;
  if (cnt < min_count) return false;
  *mean = sum / (double)cnt;
  return finite_(*mean);
}

// TriangulateTrackDLT (:52-111) over the selected member positions
LVBA_HD bool dlt_sel(const View& v, int64_t a, const int32_t* sel, int ns, double* X, double* mean, int* cnt) {
  if (ns < 4) return false;
  double AtA[16];
  for (int i = 0; i < 16; ++i) AtA[i] = 0.0;
  int rows = 0;
  for (int q = 0; q < ns; ++q) {
    const int t = sel[q];
    const int cam = v.img[a + t];
    if (cam < 0 || cam >= v.n_cams) continue;
    const int64_t g = v.kp_ptr[cam] + v.kp[a + t];
    double x, y;
    if (!depth::undistort_pixel(v.intr, (double)v.kp_uv[2 * g], (double)v.kp_uv[2 * g + 1], &x, &y)) continue;
    const double* P = v.cams + 12 * (int64_t)cam;
    double ru[4], rv[4];
    for (int k = 0; k < 3; ++k) { ru[k] = x * P[6 + k] - P[k]; rv[k] = y * P[6 + k] - P[3 + k]; }
    ru[3] = x * P[11] - P[9]; rv[3] = y * P[11] - P[10];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) AtA[4 * i + j] += ru[i] * ru[j];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) AtA[4 * i + j] += rv[i] * rv[j];
    rows += 2;
  }
  if (rows < 8) return false;
  double Xh[4];
  track::smallest_eigvec4(AtA, Xh);
  if (fabs(Xh[3]) < 1e-12) return false;
  X[0] = Xh[0] / Xh[3]; X[1] = Xh[1] / Xh[3]; X[2] = Xh[2] / Xh[3];
  if (!(finite_(X[0]) && finite_(X[1]) && finite_(X[2]))) return false;
  return mean_reproj_sel(v, a, sel, ns, X, 4, mean, cnt);
}

struct FuseF {
  View v;
  LVBA_HD void operator()(int64_t c) const {
    const int64_t a = v.ptr[c], ua = v.uptr[c];
    const int K = (int)(v.ptr[c + 1] - a), U = (int)(v.uptr[c + 1] - ua);
    const int thr = v.prm.obser_thr;
    int32_t* best = v.best + ua;
    int32_t* sel = v.sel + a;
    double* dirs = v.dirs + 3 * ua;
    for (int t = 0; t < K; ++t) v.inlier[a + t] = 0;
    v.status[c] = 0; v.mean[c] = 0.0; v.Xw[3 * c] = v.Xw[3 * c + 1] = v.Xw[3 * c + 2] = 0.0;
    if (K < thr || U < thr) return;                                         // :989, :1001
    auto gid = [&](int t) -> int64_t { return v.kp_ptr[v.img[a + t]] + v.kp[a + t]; };
    // ------------------------------------------------ depth candidate (:1016-1103)
    bool depth_ok = false;
    double Xd[3] = {0, 0, 0}, mean_d = 0.0;
    int n_valid = 0, anchor = -1;
    for (int t = 0; t < K; ++t)
      if (v.kp_valid[gid(t)]) { ++n_valid; if (anchor < 0) anchor = t; }
    int ns_depth = 0;
    if (n_valid >= thr) {
      const double* Xa = v.kp_Xw + 3 * gid(anchor);
      for (int u = 0; u < U; ++u) best[u] = -1;
      for (int t = 0; t < K; ++t) {                                         // first gated member of every image, in member order
        const int64_t g = gid(t);
        if (!v.kp_valid[g]) continue;
        const double* X = v.kp_Xw + 3 * g;
        const double d0 = X[0] - Xa[0], d1 = X[1] - Xa[1], d2 = X[2] - Xa[2];
        if (!(sqrt(d0 * d0 + d1 * d1 + d2 * d2) < v.prm.depth_gate)) continue;
        if (best[v.rank[a + t]] < 0) best[v.rank[a + t]] = t;
      }
      int nbest = 0;
      const int32_t* cand = best;                                           // the members the two loops below visit, and how many slots
      int ncand = U;
      if (v.prm.map_order == 1) {
        // best_id.reserve(inliers.size()), keys inserted in inlier (= member) order (:1051-1056); the loops of :1057 and :1069 then run in
        // the container's order
        int32_t* ins = v.ins + ua;
        int32_t* cnd = v.cand + ua;
        int n_inl = 0, n_ins = 0;
        for (int u = 0; u < U; ++u) best[u] = -1;
        for (int t = 0; t < K; ++t) {
          const int64_t g = gid(t);
          if (!v.kp_valid[g]) continue;
          const double* X = v.kp_Xw + 3 * g;
          const double d0 = X[0] - Xa[0], d1 = X[1] - Xa[1], d2 = X[2] - Xa[2];
          if (!(sqrt(d0 * d0 + d1 * d1 + d2 * d2) < v.prm.depth_gate)) continue;
          ++n_inl;
          if (best[v.rank[a + t]] < 0) { best[v.rank[a + t]] = t; ins[n_ins++] = v.rank[a + t]; }
        }
        stl_map_order(n_inl, [&](int i) { return (int)v.img[a + best[ins[i]]]; }, n_ins, cnd, v.primes, v.n_primes);
        for (int k = 0; k < n_ins; ++k) cnd[k] = best[ins[cnd[k]]];
        cand = cnd; ncand = n_ins;
      }
      for (int u = 0; u < ncand; ++u)
        if (cand[u] >= 0) { const double* X = v.kp_Xw + 3 * gid(cand[u]); Xd[0] += X[0]; Xd[1] += X[1]; Xd[2] += X[2]; ++nbest; }
      if (nbest >= thr) {
        Xd[0] /= (double)nbest; Xd[1] /= (double)nbest; Xd[2] /= (double)nbest;
        ns_depth = view_angle_filter(v, a, cand, ncand, [&](int t, double* X) { const double* p = v.kp_Xw + 3 * gid(t); X[0] = p[0]; X[1] = p[1]; X[2] = p[2]; }, dirs, sel);
        if (ns_depth >= thr) {
          int cnt = 0;
          depth_ok = mean_reproj_sel(v, a, sel, ns_depth, Xd, thr, &mean_d, &cnt) && mean_d <= v.prm.reproj_thr;
        }
      }
    }
    if (depth_ok) for (int q = 0; q < ns_depth; ++q) v.inlier[a + sel[q]] = 1;          // provisional: the depth candidate's survivors
    // ------------------------------------------------ triangulation candidate (:1106-1159)
    bool tri_ok = false;
    double Xt[3] = {0, 0, 0}, mean_t = 0.0;
    int ns_tri = 0;
    if (U >= 4) {
      double seed[3], m0; int c0;
      // all images, first member each, ascending image id: best[] is free again (the depth survivors are flagged in inlier[])
      for (int u = 0; u < U; ++u) best[u] = v.upos[ua + v.uord[ua + u]];   // `for (auto& kv : unique_id)` (:1124): ascending, or the container's order
      if (dlt_sel(v, a, best, U, seed, &m0, &c0)) {
        ns_tri = view_angle_filter(v, a, best, U, [&](int, double* X) { X[0] = seed[0]; X[1] = seed[1]; X[2] = seed[2]; }, dirs, sel);
        int c1;
        if (ns_tri >= 4 && dlt_sel(v, a, sel, ns_tri, Xt, &mean_t, &c1)) tri_ok = mean_t <= v.prm.reproj_thr;
      }
    }
    // ------------------------------------------------ choice (:1161-1199)
    int pick = 0;
    if (depth_ok && tri_ok) pick = (mean_t < mean_d) ? 2 : 1;
    else if (tri_ok) pick = 2;
    else if (depth_ok) pick = 1;
    if (pick == 2) {                                                        // the triangulation's survivors replace the depth ones
      for (int t = 0; t < K; ++t) v.inlier[a + t] = 0;
      for (int q = 0; q < ns_tri; ++q) v.inlier[a + sel[q]] = 1;
    }
    const double* X = pick == 2 ? Xt : Xd;
    if (pick != 0) {
      const bool fin = finite_(X[0]) && finite_(X[1]) && finite_(X[2]);
      const bool zero = fabs(X[0]) <= 1e-12 && fabs(X[1]) <= 1e-12 && fabs(X[2]) <= 1e-12;       // Eigen isZero(1e-12)
      if (!fin || zero) pick = 0;
    }
    if (pick == 0) { for (int t = 0; t < K; ++t) v.inlier[a + t] = 0; return; }
    v.status[c] = (uint8_t)pick;
    v.Xw[3 * c] = X[0]; v.Xw[3 * c + 1] = X[1]; v.Xw[3 * c + 2] = X[2];
    v.mean[c] = pick == 2 ? mean_t : mean_d;
  }
};

// ------------------------------------------------------------------------------------------------ the whole stage
struct Track {
  int64_t seed;                                // global keypoint id of the BFS seed that produced the track (reference track order)
  std::vector<int32_t> img, kp;                // Track::observations (the whole component, BFS order)
  std::vector<uint8_t> inlier;                 // Track::inlier_indices as flags over the observations
  double Xw[3];
  double mean;
  uint8_t source;                              // 1 depth candidate, 2 triangulation
};
struct Result {
  std::vector<Track> tracks;                   // in the order the reference appends them (by seed)
  int64_t n_components = 0, n_candidates = 0, n_rounds = 0, n_attempts = 0, n_depth = 0, n_tri = 0;
};

#ifndef LVBA_FUSE_TRY
#define LVBA_FUSE_TRY(call) do { const int rc__ = (call); if (rc__ != 0) return rc__; } while (0)
#endif

// kp_* / cams are host arrays; everything the device needs is uploaded here.  Rounds: every component is tried from its lowest
// keypoint as BFS seed; the ones that fail are tried again from their next keypoint (what the reference's scan does after
// releasing them, :1199), until they succeed or run out of seeds.
template <class Exec>
inline int run(Exec& ex, int32_t n_images, const int64_t* kp_ptr, const float* kp_uv, int64_t n_matches, const int32_t* ma_img,
               const int32_t* ma_kp, const int32_t* mb_img, const int32_t* mb_kp, const double* cams, const double* intr,
               const double* kp_Xw, const uint8_t* kp_valid, const Params& prm, Result& R) {
  R = Result();
  Graph G;
  build_adjacency(n_images, kp_ptr, n_matches, ma_img, ma_kp, mb_img, mb_kp, G);
  const int64_t n_kp = kp_ptr[n_images];
  typename Exec::template Buf<int64_t> d_kp_ptr, d_ptr, d_uptr;
  typename Exec::template Buf<float> d_uv;
  typename Exec::template Buf<double> d_Xw_in, d_cams, d_dirs, d_X, d_mean;
  typename Exec::template Buf<uint8_t> d_valid, d_status, d_inlier;
  typename Exec::template Buf<uint32_t> d_primes;
  typename Exec::template Buf<int32_t> d_img, d_kp, d_rank, d_upos, d_uord, d_best, d_sel, d_ins, d_cand;
  LVBA_FUSE_TRY(d_kp_ptr.alloc((size_t)n_images + 1)); LVBA_FUSE_TRY(ex.put(d_kp_ptr.p, kp_ptr, (size_t)n_images + 1));
  LVBA_FUSE_TRY(d_uv.alloc((size_t)std::max<int64_t>(n_kp, 1) * 2)); LVBA_FUSE_TRY(ex.put(d_uv.p, kp_uv, (size_t)n_kp * 2));
  LVBA_FUSE_TRY(d_Xw_in.alloc((size_t)std::max<int64_t>(n_kp, 1) * 3)); LVBA_FUSE_TRY(ex.put(d_Xw_in.p, kp_Xw, (size_t)n_kp * 3));
  LVBA_FUSE_TRY(d_valid.alloc((size_t)std::max<int64_t>(n_kp, 1))); LVBA_FUSE_TRY(ex.put(d_valid.p, kp_valid, (size_t)n_kp));
  LVBA_FUSE_TRY(d_cams.alloc((size_t)std::max(n_images, 1) * 12)); LVBA_FUSE_TRY(ex.put(d_cams.p, cams, (size_t)n_images * 12));
  int n_primes = 0;
  const uint32_t* h_primes = stl_prime_table(&n_primes);
  LVBA_FUSE_TRY(d_primes.alloc((size_t)n_primes)); LVBA_FUSE_TRY(ex.put(d_primes.p, h_primes, (size_t)n_primes));
  LVBA_FUSE_TRY(ex.sync());

  struct Pending { std::vector<int64_t> sorted; int attempt; };
  std::vector<Pending> pend;                   // components of the current round, in batch order
  std::vector<std::vector<int64_t>> order;     // their members in BFS order
  std::vector<int32_t> mark((size_t)n_kp, 0);
  std::vector<int64_t> members;
  Batch B;
  B.map_order = prm.map_order;
  // round 0: the scan of :964-966
  for (int64_t g = 0; g < n_kp; ++g) {
    if (mark[(size_t)g] != 0) continue;
    bfs_component(G, g, mark, 1, members);
    ++R.n_components;
    if ((int)members.size() < prm.obser_thr) continue;                      // :989 (retries cannot change the size)
    const int U = B.add(G, members);
    if (U < prm.obser_thr) {                                                // :1001: drop it from the batch again
      B.ptr.pop_back(); B.uptr.pop_back();
      B.img.resize((size_t)B.ptr.back()); B.kp.resize((size_t)B.ptr.back()); B.rank.resize((size_t)B.ptr.back()); B.upos.resize((size_t)B.uptr.back());
      B.uord.resize((size_t)B.uptr.back());
      continue;
    }
    ++R.n_candidates;
    Pending p; p.sorted = members; std::sort(p.sorted.begin(), p.sorted.end()); p.attempt = 0;
    pend.push_back(std::move(p));
    order.push_back(members);
  }
  std::vector<uint8_t> h_status, h_inlier;
  std::vector<double> h_X, h_mean;
  while (B.size() > 0) {
    const int64_t nc = B.size(), nm = B.ptr.back(), nu = B.uptr.back();
    ++R.n_rounds; R.n_attempts += nc;
    LVBA_FUSE_TRY(d_ptr.alloc((size_t)nc + 1)); LVBA_FUSE_TRY(d_uptr.alloc((size_t)nc + 1));
    LVBA_FUSE_TRY(d_img.alloc((size_t)nm)); LVBA_FUSE_TRY(d_kp.alloc((size_t)nm)); LVBA_FUSE_TRY(d_rank.alloc((size_t)nm)); LVBA_FUSE_TRY(d_upos.alloc((size_t)nu));
    LVBA_FUSE_TRY(d_uord.alloc((size_t)nu)); LVBA_FUSE_TRY(d_ins.alloc((size_t)nu)); LVBA_FUSE_TRY(d_cand.alloc((size_t)nu));
    LVBA_FUSE_TRY(d_best.alloc((size_t)nu)); LVBA_FUSE_TRY(d_sel.alloc((size_t)nm)); LVBA_FUSE_TRY(d_dirs.alloc((size_t)nu * 3));
    LVBA_FUSE_TRY(d_status.alloc((size_t)nc)); LVBA_FUSE_TRY(d_X.alloc((size_t)nc * 3)); LVBA_FUSE_TRY(d_mean.alloc((size_t)nc)); LVBA_FUSE_TRY(d_inlier.alloc((size_t)nm));
    LVBA_FUSE_TRY(ex.put(d_ptr.p, B.ptr.data(), (size_t)nc + 1)); LVBA_FUSE_TRY(ex.put(d_uptr.p, B.uptr.data(), (size_t)nc + 1));
    LVBA_FUSE_TRY(ex.put(d_img.p, B.img.data(), (size_t)nm)); LVBA_FUSE_TRY(ex.put(d_kp.p, B.kp.data(), (size_t)nm));
    LVBA_FUSE_TRY(ex.put(d_rank.p, B.rank.data(), (size_t)nm)); LVBA_FUSE_TRY(ex.put(d_upos.p, B.upos.data(), (size_t)nu));
    LVBA_FUSE_TRY(ex.put(d_uord.p, B.uord.data(), (size_t)nu));
    View v{nc, d_ptr.p, d_img.p, d_kp.p, d_rank.p, d_uptr.p, d_upos.p, d_uord.p, d_kp_ptr.p, d_uv.p, d_Xw_in.p, d_valid.p, n_images, d_cams.p, {}, prm,
           d_best.p, d_sel.p, d_dirs.p, d_ins.p, d_cand.p, d_primes.p, n_primes, d_status.p, d_X.p, d_mean.p, d_inlier.p};
    for (int q = 0; q < 8; ++q) v.intr[q] = intr[q];
    LVBA_FUSE_TRY(ex.for_each(nc, FuseF{v}));
    h_status.resize((size_t)nc); h_inlier.resize((size_t)nm); h_X.resize((size_t)nc * 3); h_mean.resize((size_t)nc);
    LVBA_FUSE_TRY(ex.fetch(h_status.data(), d_status.p, (size_t)nc));
    LVBA_FUSE_TRY(ex.fetch(h_inlier.data(), d_inlier.p, (size_t)nm));
    LVBA_FUSE_TRY(ex.fetch(h_X.data(), d_X.p, (size_t)nc * 3));
    LVBA_FUSE_TRY(ex.fetch(h_mean.data(), d_mean.p, (size_t)nc));
    // successes become tracks; failures are released and re-entered from their next seed
    std::vector<Pending> next_pend;
    std::vector<std::vector<int64_t>> next_order;
    Batch NB;
    NB.map_order = prm.map_order;
    for (int64_t c = 0; c < nc; ++c) {
      Pending& p = pend[(size_t)c];
      if (h_status[(size_t)c] != 0) {
        Track t;
        t.seed = p.sorted[(size_t)p.attempt];
        const int64_t a = B.ptr[(size_t)c], b = B.ptr[(size_t)c + 1];
        t.img.assign(B.img.begin() + a, B.img.begin() + b); t.kp.assign(B.kp.begin() + a, B.kp.begin() + b);
        t.inlier.assign(h_inlier.begin() + a, h_inlier.begin() + b);
        for (int q = 0; q < 3; ++q) t.Xw[q] = h_X[(size_t)c * 3 + q];
        t.mean = h_mean[(size_t)c]; t.source = h_status[(size_t)c];
        if (t.source == 1) ++R.n_depth; else ++R.n_tri;
        R.tracks.push_back(std::move(t));
        continue;
      }
      if (p.attempt + 1 >= (int)p.sorted.size()) continue;                  // every member has been the seed once
      ++p.attempt;
      // the reference's scan reaches the next member only if it is still unmarked, i.e. not part of a track: always true here
      for (int64_t g : p.sorted) mark[(size_t)g] = 0;
      bfs_component(G, p.sorted[(size_t)p.attempt], mark, 1, members);
      NB.add(G, members);
      next_order.push_back(members);
      next_pend.push_back(std::move(p));
    }
    B = std::move(NB); pend = std::move(next_pend); order = std::move(next_order);
  }
  std::stable_sort(R.tracks.begin(), R.tracks.end(), [](const Track& x, const Track& y) { return x.seed < y.seed; });
  return 0;
}

}  // namespace fuse
}  // namespace lvba
