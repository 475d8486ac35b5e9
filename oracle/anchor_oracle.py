"""CPU restatement of the anchor-cloud tail of runWindowBA — TEST INFRASTRUCTURE; nothing in the product imports it.
  rel poses              src/lvba_system.cpp:286-289   rel.R = anchor.R^T x.R ; rel.p = anchor.R^T (x.p - anchor.p)
  pl_transform           include/BALM/tools.hpp:385-395   p <- (float)(R p + t) per coordinate
  down_sampling_voxel2   include/BALM/tools.hpp:301-359   per voxel the original point closest to the voxel centre, first wins ties
The reference returns the survivors in unordered_map order; both restatements here return them sorted by voxel key.
PARITY: pinned against the reference's own source (include/BALM/tools.hpp compiled where it lies): the surviving float32 points equal bit for bit,
as a set (tests/golden/ref_balm.npz, ref_system.npz; tests/test_ref_pin.py, tests/test_ref_system_pin.py)."""
from __future__ import annotations

import numpy as np


def rel_poses(poses, win_ptr):
    """Pose of every scan in the frame of the first pose of its window (the anchor, :284)."""
    rel = np.zeros_like(poses)
    for w in range(len(win_ptr) - 1):
        a = win_ptr[w]
        Ra = poses[a, :9].reshape(3, 3); pa = poses[a, 9:]
        for j in range(win_ptr[w], win_ptr[w + 1]):
            R = poses[j, :9].reshape(3, 3)
            rel[j, :9] = (Ra.T @ R).ravel(); rel[j, 9:] = Ra.T @ (poses[j, 9:] - pa)
    return rel


def transform(scan, rel):
    p = np.asarray(scan, np.float32).reshape(-1, 3).astype(np.float64)
    R = rel[:9].reshape(3, 3); t = rel[9:]
    w = np.empty_like(p)
    for k in range(3):
        w[:, k] = ((R[k, 0] * p[:, 0] + R[k, 1] * p[:, 1]) + R[k, 2] * p[:, 2]) + t[k]
    return w.astype(np.float32)


def _keys_d2(pts, leaf):
    loc = (pts.astype(np.float64) / leaf).astype(np.float32)                   # :317-319
    loc = np.where(loc < 0, loc - np.float32(1.0), loc).astype(np.float32)    # :321-323
    key = np.trunc(loc).astype(np.int64)
    c = (key.astype(np.float64) + 0.5) * leaf                                  # :332-334
    d = pts.astype(np.float64) - c
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return key, d2


def anchor_clouds_literal(scans, rel, win_ptr, leaf):
    out = []
    for w in range(len(win_ptr) - 1):
        merged = [transform(scans[j], rel[j]) for j in range(win_ptr[w], win_ptr[w + 1])]
        merged = np.concatenate(merged) if merged else np.zeros((0, 3), np.float32)
        if leaf < 0.001:
            out.append(merged); continue
        key, d2 = _keys_d2(merged, leaf)
        best = {}
        for i in range(len(merged)):
            k = (int(key[i, 0]), int(key[i, 1]), int(key[i, 2]))
            if k not in best or d2[i] < best[k][0]:
                best[k] = (d2[i], i)
        out.append(np.array([merged[best[k][1]] for k in sorted(best)], np.float32).reshape(-1, 3))
    return out


def anchor_clouds(scans, rel, win_ptr, leaf):
    out = []
    for w in range(len(win_ptr) - 1):
        merged = [transform(scans[j], rel[j]) for j in range(win_ptr[w], win_ptr[w + 1])]
        merged = np.concatenate(merged) if merged else np.zeros((0, 3), np.float32)
        if leaf < 0.001 or len(merged) == 0:
            out.append(merged); continue
        key, d2 = _keys_d2(merged, leaf)
        order = np.lexsort((np.arange(len(merged)), d2, key[:, 2], key[:, 1], key[:, 0]))   # by key, then d2, then cloud order
        ks = key[order]
        head = np.concatenate([[True], np.any(ks[1:] != ks[:-1], axis=1)])
        out.append(merged[order[head]])
    return out
