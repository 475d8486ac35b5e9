// nd_plan.h — symbolic plan of the substructured ("nested dissection") block LDL^T of a block-banded pose system.
//
// The reference hands the pose system to Eigen::SimplicialLDLT (include/BALM/bavoxel.hpp:695-710) / the Ceres
// DENSE_SCHUR Cholesky (src/lvba_system.cpp:1573-1575): one sequential elimination.  A right-looking LDL^T of a banded
// matrix is a chain of n strictly sequential pivot columns (envelope.cuh / factor_la.cuh: one SM per chain), so the
// chain is cut instead ("substructuring"):
//
//   rows      |-- I_0 --|S_1|-- I_1 --|S_2|-- I_2 --| ... |S_{p-1}|-- I_{p-1} --|
//
// * separator S_j starts at row s_j and is as wide as the band there (w_j = last[s_j - 1] + 1 - s_j): no row beyond it
//   couples to a row before it, so the interiors I_c are mutually independent once the separators are taken out;
// * every interior is eliminated by its own CTA (leaf), every separator by a node of a balanced binary tree over
//   1..p-1: a node's pivot rows are coupled to at most two not-yet-eliminated separators, its BOUNDARY (a on the left,
//   c on the right) — the nearest ancestors on either side;
// * eliminating the pivot rows K of a node leaves the Schur update  U = -E^T K^-1 E  on its boundary rows, where E are
//   the boundary's columns of the pivot rows; with K = L D L^T that is  Z = L^-1 E  (forward substitution of 6(w_a+w_c)
//   right-hand sides, the "spike"), U = -Z^T D^-1 Z (a SYRK).  A parent assembles its pivot block, its E and the initial
//   U from the original matrix and its two children's U (multifrontal extend-add);
// * downwards x_K = L^-T D^-1 (w_K - Z x_boundary).
//
// The critical path is  (longest interior) + (tree depth) x (separator width)  pivot columns instead of n.
//
// This header is plain C++ (no CUDA): the plan is built once per structure on the host, checked on the CPU against
// dense solves through tests/emu/nd_emu.cpp, and uploaded by csrc/nd_solver.cuh.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace lvba {
namespace nd {

constexpr int kMaxSepWidth = 30;      // a separator is factorised by the register-window kernel (columns of <= 30 blocks)
constexpr int kMinInterior = 4;

struct Node {
  int kind = 0;              // 0: leaf (interior of a chunk), 1: separator
  int r0 = 0, npiv = 0;      // pivot rows [r0, r0 + npiv) of the global system
  int ntrail = 0;            // leaf: rows of the right boundary separator, carried as trailing rows of the banded factorisation
  int sa = 0, wa = 0;        // left boundary separator: first row, width (0: none)
  int sc = 0, wc = 0;        // right boundary separator
  int left = -1, right = -1; // children (separator nodes)
  int parent = -1;
  int level = 0;             // leaves 0; a separator node = 1 + max(level of its children)
  int ks = 0;                // right-hand sides of the spike: leaf 6 wa (the right boundary rides in the band), separator 6 (wa + wc)
  int nE = 0;                // leaf: leading pivot rows that couple to the left boundary
  int nb = 0;                // block rows of the update matrix U: wa + wc
  // offsets (in doubles) into the pooled buffers
  long long offU = 0, offu = 0, offZ = 0, offE = 0, offT = 0, offW = 0, offw = 0;
  int zrows = 0;             // rows of Z: npiv + ntrail
  int owner = 0;             // rank that eliminates this node; -1: a node of the top tree, eliminated by every rank (replicated)
};

struct Plan {
  int n = 0, p = 0, max_col = 0;
  std::vector<int> sep_start, sep_width;       // [p+1]; entries 1..p-1 used
  std::vector<Node> nodes;                     // leaves 0..p-1 (chunk order), then separator nodes
  std::vector<std::vector<int>> levels;        // node ids per level (level 0 = leaves)
  int root = -1;
  // leaf views of the envelope (indexed by GLOBAL row; a leaf's view starts at its r0)
  std::vector<int> first_rel, last_rel;
  std::vector<long long> rs_adj;
  long long sizeU = 0, sizeu = 0, sizeZ = 0, sizeE = 0, sizeT = 0, sizeW = 0, sizew = 0;
  int max_ks = 0, max_zrows_leaf = 0;
  // ---- multi-GPU: rank r eliminates the q = p / n_ranks consecutive chunks r q .. (r+1) q - 1 and the
  //      separators between them (a complete subtree); the n_ranks - 1 separators between rank ranges ("rank separators",
  //      rows owned by the rank on their left) form the top tree, eliminated redundantly by everybody after ONE all-gather of
  //      a fixed-size slot per rank: the update matrix of the rank's subtree root, and the rows of its rank separator
  int n_ranks = 1, q = 0;
  int local_levels = 0;                        // levels 0 .. local_levels-1 hold rank-owned nodes, the rest the top tree
  std::vector<int> rank_row_begin, rank_row_end;   // rows owned by rank r (its chunks, inner separators and its right rank separator)
  std::vector<int> rank_root;                  // node id of the subtree root of rank r
  long long slot = 0;                          // doubles per rank in the exchange region (at offset region0 of the U pool)
  long long region0 = 0, slotU = 0, slotu = 0, slotH = 0;
  int slot_rows = 0;                           // block rows of a rank separator carried in a slot (<= kMaxSepWidth)
};

// Builds the plan for `p_want` chunks; returns false (plan untouched apart from scratch) when the structure does not
// allow it: envelope wider than the separator kernel, chunks shorter than kMinInterior, a band that vanishes at a cut.
inline bool build_plan(int n, const int* first, const int* last, const long long* row_start, int max_col, int p_want, Plan& P,
                       int n_ranks = 1) {
  if (p_want < 2 || max_col > kMaxSepWidth || max_col < 1 || n < 2 * kMinInterior + max_col) return false;
  if (n_ranks < 1 || p_want % n_ranks != 0) return false;
  // ---- cuts: equal interiors, separators as wide as the band where they start
  std::vector<int> ss(p_want + 1, 0), sw(p_want + 1, 0), i0(p_want, 0), i1(p_want, 0);
  {
    long long est = (long long)n - (long long)(p_want - 1) * max_col;
    if (est < (long long)p_want * kMinInterior) return false;
    int cursor = 0;
    for (int c = 0; c + 1 < p_want; ++c) {
      const int remaining_chunks = p_want - c;
      const long long left = (long long)n - cursor - (long long)(remaining_chunks - 1) * max_col;
      int m = (int)std::max<long long>(kMinInterior, left / remaining_chunks);
      int s = cursor + m;
      if (s >= n - kMinInterior) return false;
      int w = last[s - 1] + 1 - s;
      if (w < 1 || w > kMaxSepWidth) return false;
      if (s + w > n - kMinInterior) return false;
      i0[c] = cursor; i1[c] = s; ss[c + 1] = s; sw[c + 1] = w;
      cursor = s + w;
    }
    i0[p_want - 1] = cursor; i1[p_want - 1] = n;
    if (n - cursor < kMinInterior) return false;
  }
  P = Plan();
  P.n = n; P.p = p_want; P.max_col = max_col;
  P.n_ranks = n_ranks; P.q = p_want / n_ranks;
  P.sep_start = ss; P.sep_width = sw;
  const int p = p_want;
  // ---- leaves
  P.nodes.resize((size_t)p);
  for (int c = 0; c < p; ++c) {
    Node& v = P.nodes[c];
    v.kind = 0; v.r0 = i0[c]; v.npiv = i1[c] - i0[c];
    if (c > 0) { v.sa = ss[c]; v.wa = sw[c]; }
    if (c + 1 < p) { v.sc = ss[c + 1]; v.wc = sw[c + 1]; v.ntrail = v.wc; }
    v.ks = 6 * v.wa;
    v.nb = v.wa + v.wc;
    v.zrows = v.npiv + v.ntrail;
    v.nE = 0;
    if (c > 0) v.nE = std::min(v.npiv, std::max(0, last[v.r0 - 1] - v.r0 + 1));
    v.level = 0;
    v.owner = c / P.q;
  }
  // ---- balanced tree over the separators lo..hi (1-based), children = sub-ranges or leaves
  struct Rec {
    Plan& P; const std::vector<int>& ss; const std::vector<int>& sw; int p;
    int build(int lo, int hi) {                          // returns node id covering separators lo..hi (lo <= hi)
      const int mid = (lo + hi) / 2;
      const int id = (int)P.nodes.size();
      P.nodes.push_back(Node());
      const int lc = (lo <= mid - 1) ? build(lo, mid - 1) : mid - 1;        // leaf chunk mid-1 sits left of separator mid
      const int rc = (mid + 1 <= hi) ? build(mid + 1, hi) : mid;            // leaf chunk mid sits right of it
      Node& v = P.nodes[id];
      v.kind = 1; v.r0 = ss[mid]; v.npiv = sw[mid]; v.ntrail = 0;
      if (lo - 1 >= 1) { v.sa = ss[lo - 1]; v.wa = sw[lo - 1]; }
      if (hi + 1 <= p - 1) { v.sc = ss[hi + 1]; v.wc = sw[hi + 1]; }
      v.left = lc; v.right = rc;
      P.nodes[lc].parent = id; P.nodes[rc].parent = id;
      v.level = 1 + std::max(P.nodes[lc].level, P.nodes[rc].level);
      v.ks = 6 * (v.wa + v.wc);
      v.nb = v.wa + v.wc;
      v.zrows = v.npiv;
      v.owner = ((lo - 1) / P.q == hi / P.q) ? (lo - 1) / P.q : -1;      // chunks lo-1 .. hi all belong to one rank?
      return id;
    }
  } rec{P, ss, sw, p};
  P.nodes.reserve((size_t)2 * p);
  P.root = rec.build(1, p - 1);
  int max_level = 0;
  for (const Node& v : P.nodes) max_level = std::max(max_level, v.level);
  P.levels.assign((size_t)max_level + 1, {});
  for (int id = 0; id < (int)P.nodes.size(); ++id) P.levels[P.nodes[id].level].push_back(id);
  // ---- rank subtree roots, row ownership, levels of the top tree
  P.rank_root.assign((size_t)n_ranks, -1);
  P.rank_row_begin.assign((size_t)n_ranks, 0); P.rank_row_end.assign((size_t)n_ranks, n);
  for (int id = 0; id < (int)P.nodes.size(); ++id) {
    const Node& v = P.nodes[id];
    if (v.owner < 0) continue;
    if (P.rank_root[v.owner] < 0 || v.level > P.nodes[P.rank_root[v.owner]].level) P.rank_root[v.owner] = id;
  }
  for (int r = 0; r < n_ranks; ++r) {
    P.rank_row_begin[r] = P.nodes[r * P.q].r0;
    if (r > 0) P.rank_row_begin[r] = P.rank_row_end[r - 1];
    P.rank_row_end[r] = (r + 1 < n_ranks) ? ss[(r + 1) * P.q] + sw[(r + 1) * P.q] : n;
  }
  P.local_levels = 1;
  for (const Node& v : P.nodes) if (v.owner >= 0) P.local_levels = std::max(P.local_levels, v.level + 1);
  for (const Node& v : P.nodes) if (v.owner < 0 && v.level < P.local_levels) return false;     // (cannot happen: equal subtrees)
  // ---- pooled buffers.  U and u share one pool (offU / offu are offsets into the same array); the update matrices of the
  //      rank roots live in the exchange region at its end: one fixed-size slot per rank = [U][u][rows of its rank separator][dadd]
  std::vector<char> is_root(P.nodes.size(), 0);
  if (n_ranks > 1) for (int r = 0; r < n_ranks; ++r) is_root[P.rank_root[r]] = 1;
  for (size_t id = 0; id < P.nodes.size(); ++id) {
    Node& v = P.nodes[id];
    if (!is_root[id]) { v.offU = P.sizeU; P.sizeU += (long long)v.nb * (v.nb + 1) / 2 * 36; }
    v.offZ = P.sizeZ; P.sizeZ += (long long)v.zrows * 6 * v.ks;
    if (v.kind == 0) {
      v.offE = P.sizeE; P.sizeE += (long long)v.nE * 6 * v.ks;
      v.offW = P.sizeW; P.sizeW += (long long)v.ntrail * v.ntrail * 36;
      v.offw = P.sizew; P.sizew += (long long)v.ntrail * 6;
      P.max_zrows_leaf = std::max(P.max_zrows_leaf, v.zrows);
    } else {
      v.offE = P.sizeE; P.sizeE += (long long)v.npiv * 6 * v.ks;
      v.offT = P.sizeT; P.sizeT += (long long)v.npiv * (v.npiv + 1) / 2 * 36;
    }
    P.max_ks = std::max(P.max_ks, v.ks);
  }
  for (size_t id = 0; id < P.nodes.size(); ++id) {
    Node& v = P.nodes[id];
    if (!is_root[id]) { v.offu = P.sizeU; P.sizeU += (long long)v.nb * 6; }
  }
  if (n_ranks > 1) {
    for (int r = 0; r < n_ranks; ++r) {
      const Node& v = P.nodes[P.rank_root[r]];
      P.slotU = std::max(P.slotU, (long long)v.nb * (v.nb + 1) / 2 * 36);
      P.slotu = std::max(P.slotu, (long long)v.nb * 6);
    }
    P.slot_rows = 0;
    for (int r = 0; r + 1 < n_ranks; ++r) P.slot_rows = std::max(P.slot_rows, sw[(r + 1) * P.q]);
    P.slotH = (long long)P.slot_rows * (max_col + 1) * 36;
    P.slot = P.slotU + P.slotu + P.slotH + (long long)P.slot_rows * 6;
    P.slot = (P.slot + 1) & ~1LL;
    P.region0 = (P.sizeU + 1) & ~1LL;
    for (int r = 0; r < n_ranks; ++r) {
      Node& v = P.nodes[P.rank_root[r]];
      v.offU = P.region0 + (long long)r * P.slot;
      v.offu = v.offU + P.slotU;
    }
    P.sizeU = P.region0 + (long long)n_ranks * P.slot;
  }
  P.sizeu = 0;       // (u lives in the U pool)
  // ---- leaf views: a leaf sees rows r0 .. r0+npiv+ntrail-1, columns >= r0 (the couplings to the left boundary are the
  //      spike's right-hand sides, not part of the banded factorisation)
  P.first_rel.assign((size_t)n, 0); P.last_rel.assign((size_t)n, 0); P.rs_adj.assign((size_t)n + 1, 0);
  for (int c = 0; c < p; ++c) {
    const Node& v = P.nodes[c];
    const int end = v.r0 + v.npiv + v.ntrail;
    for (int r = v.r0; r < end; ++r) {
      const int f = std::max(first[r], v.r0);
      P.first_rel[r] = f - v.r0;
      P.rs_adj[r] = row_start[r] + (f - first[r]);
      P.last_rel[r] = std::min(last[r], end - 1) - v.r0;
    }
    // the view must be a valid monotone envelope with columns of at most max_col blocks
    for (int r = v.r0 + 1; r < end; ++r)
      if (P.first_rel[r] < P.first_rel[r - 1]) return false;
  }
  P.rs_adj[n] = row_start[n];
  return true;
}

// largest number of chunks (a power of two is not required) not above p_want for which a plan exists
inline int choose_chunks(int n, const int* first, const int* last, const long long* row_start, int max_col, int p_want, Plan& P,
                         int n_ranks = 1) {
  for (int p = p_want - p_want % n_ranks; p >= 2 && p >= n_ranks; p -= n_ranks)
    if (build_plan(n, first, last, row_start, max_col, p, P, n_ranks)) return p;
  return 0;
}

}  // namespace nd
}  // namespace lvba
