"""CPU restatement of the reference's set-up stage — TEST INFRASTRUCTURE, groundwork
for a GPU voxelisation kernel; nothing in the product imports it.
PARITY: pinned against the reference's own source (include/BALM/bavoxel.hpp compiled where it lies, oracle/ref_driver.cpp): keys, octree paths,
layers, pose lists and — for the literal restatement — the cluster sums to the bit, plane look-ups, the LM on the map (tests/golden/ref_balm.npz,
tests/test_ref_pin.py).

Follows, line by line:
  cut_voxel            include/BALM/bavoxel.hpp:799-836   points -> root voxel (float division, negative fix, int64 key)
  OCTO_TREE_NODE       include/BALM/bavoxel.hpp:282-300   float voxel_center / quater_length
  recut                include/BALM/bavoxel.hpp:420-464   min_ps test, plane test, split into <= 8 leaves, layer_limit
  judge_eigen          include/BALM/bavoxel.hpp:335-352   merged world-frame covariance, lambda0/lambda2 <= eigen_ratio_array[layer]
  cut_func             include/BALM/bavoxel.hpp:391-418   octant by strict '>' against the float centre
  tras_opt/push_voxel  include/BALM/bavoxel.hpp:466-474, 45-54   PLANE nodes seen from >= 2 poses become BA voxels
  findCorrespondPoint  include/BALM/bavoxel.hpp:320-333 and recompute_local_planes, src/lvba_system.cpp:1529-1566
  PointCluster         include/BALM/tools.hpp:407-456

Two independent implementations that the tests compare:
  * voxelize()          vectorised: the octant path of a point is a pure function of its world position and root key,
                        so every point gets its (root, leaf1, leaf2) keys at once and nodes are group-by reductions
                        (this is the shape a GPU kernel would take);
  * voxelize_literal()  the reference's recursive tree with per-node point lists, for small inputs.
Output of both: plane voxels in the layout lvba_lidar_lm takes (CSR over voxels, ascending pose index inside a voxel,
10-double body-frame clusters) plus per-voxel metadata (root key, octant path, layer, centre, normal, eigenvalues);
voxels are ordered by (root key, path) because the reference's unordered_map order is unspecified.
"""
from __future__ import annotations

import numpy as np

EIGEN_RATIO_DEFAULT = (0.3, 0.1, 0.06, 0.03)     # bavoxel.hpp:17 (compiled-in)
LAYER_LIMIT = 2                                   # bavoxel.hpp:13
MIN_PS = 15                                       # bavoxel.hpp:24


def root_keys(world, voxel_size):
    """cut_voxel :808-816.  loc = (float)(pw / voxel_size); if (loc < 0) loc -= 1.0 ; key = (int64)loc (truncation)."""
    loc = (world / voxel_size).astype(np.float32)
    loc = np.where(loc < 0, (loc.astype(np.float64) - 1.0).astype(np.float32), loc)
    return np.trunc(loc).astype(np.int64)


def _root_centre(key, voxel_size):
    """:829-832 — (0.5 + key) * voxel_size in double, stored to the float member; quater_length = voxel_size / 4 (float)."""
    return ((0.5 + key.astype(np.float64)) * voxel_size).astype(np.float32), np.float32(voxel_size / 4.0)


def _octant(world, centre):
    """cut_func :399-403 — strict '>' of the double coordinate against the float centre; leafnum = 4x + 2y + z."""
    b = (world > centre.astype(np.float64)).astype(np.int64)
    return 4 * b[..., 0] + 2 * b[..., 1] + b[..., 2], b


def _child_centre(centre, bits, quater):
    """:407-410 — float arithmetic: centre + (2*bit - 1) * quater_length ; child quater = quater / 2."""
    return (centre + (2 * bits - 1).astype(np.float32) * quater).astype(np.float32), np.float32(quater / np.float32(2.0))


def _transform_cluster(P, v, N, R, t):
    """PointCluster::transform, tools.hpp:443-449."""
    Rv = R @ v
    rp = np.outer(Rv, t)
    return R @ P @ R.T + rp + rp.T + N * np.outer(t, t), Rv + N * t, N


def _plane_test(clusters_by_pose, poses, layer, eigen_ratio):
    """judge_eigen :335-352.  clusters_by_pose: {pose: (P, v, N)} body frame.  Returns (is_plane, centre, direct, eigenvalues)."""
    Pm = np.zeros((3, 3)); vm = np.zeros(3); Nm = 0
    for i in sorted(clusters_by_pose):
        P, v, N = clusters_by_pose[i]
        if N > 0:
            R = poses[i, :9].reshape(3, 3); t = poses[i, 9:]
            Pt, vt, _ = _transform_cluster(P, v, N, R, t)
            Pm += Pt; vm += vt; Nm += N
    c = vm / Nm
    lam, U = np.linalg.eigh(Pm / Nm - np.outer(c, c))
    ratio = lam[0] / lam[2]
    return bool(not (ratio > np.float32(eigen_ratio[layer]))), c, U[:, 0], lam


def _emit(out, key, path, layer, clusters_by_pose, plane):
    if sum(1 for c in clusters_by_pose.values() if c[2] != 0) < 2:          # push_voxel :47-52
        return
    out.append(dict(key=tuple(int(k) for k in key), path=tuple(path), layer=layer, clusters=clusters_by_pose,
                    centre=plane[1], direct=plane[2], eigenvalues=plane[3]))


def _pack(voxels):
    voxels = sorted(voxels, key=lambda v: (v["key"], v["path"]))
    vox_ptr = [0]; pose_idx = []; clusters = []
    for v in voxels:
        for i in sorted(v["clusters"]):
            P, vv, N = v["clusters"][i]
            if N == 0:
                continue
            pose_idx.append(i)
            clusters.append([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], vv[0], vv[1], vv[2], float(N)])
        vox_ptr.append(len(pose_idx))
    meta = dict(key=np.array([v["key"] for v in voxels], np.int64).reshape(-1, 3),
                path=[v["path"] for v in voxels], layer=np.array([v["layer"] for v in voxels], np.int32),
                centre=np.array([v["centre"] for v in voxels]).reshape(-1, 3),
                direct=np.array([v["direct"] for v in voxels]).reshape(-1, 3),
                eigenvalues=np.array([v["eigenvalues"] for v in voxels]).reshape(-1, 3))
    return (np.asarray(vox_ptr, np.int64), np.asarray(pose_idx, np.int32),
            np.asarray(clusters, np.float64).reshape(-1, 10), meta)


def _world(points, pose):
    R = pose[:9].reshape(3, 3); t = pose[9:]
    return points.astype(np.float64) @ R.T + t                             # :806-807 (float xyz cast to double)


def voxelize(scans, poses, voxel_size=1.0, eigen_ratio=EIGEN_RATIO_DEFAULT, layer_limit=LAYER_LIMIT, min_ps=MIN_PS):
    """scans: list of (n_i, 3) float32 body-frame points, one per pose; poses: (W, 12).  Vectorised restatement."""
    assert layer_limit <= 2
    W = len(scans)
    pts = np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in scans]) if W else np.zeros((0, 3), np.float32)
    pose_of = np.concatenate([np.full(len(s), i, np.int64) for i, s in enumerate(scans)]) if W else np.zeros(0, np.int64)
    world = np.concatenate([_world(np.asarray(s, np.float32).reshape(-1, 3), poses[i]) for i, s in enumerate(scans)]) if W else np.zeros((0, 3))
    key = root_keys(world, voxel_size)
    c0, q0 = _root_centre(key, voxel_size)
    o1, b1 = _octant(world, c0)
    c1, q1 = _child_centre(c0, b1, q0)
    o2, _ = _octant(world, c1)
    pd = pts.astype(np.float64)
    outer = np.einsum("ni,nj->nij", pd, pd)

    def groups(cols):
        """{node id tuple: {pose: (P, v, N)}} by lexicographic group-by over the given integer columns + pose."""
        if len(pd) == 0:
            return {}
        M = np.column_stack(cols + [pose_of])
        order = np.lexsort(M.T[::-1])
        Ms = M[order]
        start = np.concatenate([[True], np.any(Ms[1:] != Ms[:-1], axis=1)])
        idx = np.nonzero(start)[0]
        Ps = np.add.reduceat(outer[order], idx); vs = np.add.reduceat(pd[order], idx)
        Ns = np.diff(np.concatenate([idx, [len(order)]]))
        res = {}
        for r, (P, v, N) in zip(Ms[idx], zip(Ps, vs, Ns)):
            res.setdefault(tuple(int(x) for x in r[:-1]), {})[int(r[-1])] = (P, v, int(N))
        return res

    kx, ky, kz = key[:, 0], key[:, 1], key[:, 2]
    lvl0 = groups([kx, ky, kz])
    lvl1 = groups([kx, ky, kz, o1])
    lvl2 = groups([kx, ky, kz, o1, o2])
    children1 = {}
    for k in lvl1:
        children1.setdefault(k[:3], []).append(k)
    children2 = {}
    for k in lvl2:
        children2.setdefault(k[:4], []).append(k)
    out = []

    def recut(node_key, layer, table):
        cl = table[node_key]
        if sum(c[2] for c in cl.values()) < min_ps:                        # :427-433
            return
        plane = _plane_test(cl, poses, layer, eigen_ratio)
        if plane[0]:                                                       # :435-443
            _emit(out, node_key[:3], node_key[3:], layer, cl, plane)
            return
        if layer == layer_limit:                                           # :446-452
            return
        kids, nxt = (children1, lvl1) if layer == 0 else (children2, lvl2)
        for k in sorted(kids.get(node_key, [])):                           # :453-461
            recut(k, layer + 1, nxt)

    for k in sorted(lvl0):
        recut(k, 0, lvl0)
    return _pack(out)


class _Node:
    def __init__(self, W, centre, quater, layer):
        self.pts = [[] for _ in range(W)]
        self.centre = centre; self.quater = quater; self.layer = layer
        self.leaves = {}
        self.state = "UNKNOWN"; self.plane = None


def build_tree_literal(scans, poses, voxel_size=1.0, eigen_ratio=EIGEN_RATIO_DEFAULT, layer_limit=LAYER_LIMIT, min_ps=MIN_PS):
    """The reference's surf_map after cut_voxel + recut, literally: {root key: _Node} with per-node point lists, states
    UNKNOWN (split; children in .leaves) / MID_NODE / PLANE and, for every node that reached judge_eigen, .plane =
    (is_plane, center, direct, eigenvalues).  Small inputs only."""
    W = len(scans)
    roots = {}
    for i, s in enumerate(scans):                                          # cut_voxel, one call per scan (:254-257 of lvba_system.cpp)
        for p in np.asarray(s, np.float32).reshape(-1, 3):
            pw = _world(p[None, :], poses[i])[0]
            key = tuple(int(k) for k in root_keys(pw[None, :], voxel_size)[0])
            if key not in roots:
                c, q = _root_centre(np.array(key), voxel_size)
                roots[key] = _Node(W, c, q, 0)
            roots[key].pts[i].append(p)

    def recut(node):
        node.clusters = _clusters_of(node, W)
        if sum(c[2] for c in node.clusters.values()) < min_ps:             # :427-433
            node.state = "MID_NODE"
            return
        node.plane = _plane_test(node.clusters, poses, node.layer, eigen_ratio)
        if node.plane[0]:                                                  # :435-443
            node.state = "PLANE"
            return
        if node.layer == layer_limit:                                      # :446-452
            node.state = "MID_NODE"
            return
        for i in range(W):                                                 # cut_func per pose (:391-418)
            for p in node.pts[i]:
                pw = _world(np.asarray(p)[None, :], poses[i])[0]
                leaf, bits = _octant(pw, node.centre)
                leaf = int(leaf)
                if leaf not in node.leaves:
                    c, q = _child_centre(node.centre, bits, node.quater)
                    node.leaves[leaf] = _Node(W, c, q, node.layer + 1)
                node.leaves[leaf].pts[i].append(p)
        for leaf in sorted(node.leaves):
            recut(node.leaves[leaf])

    for key in sorted(roots):
        recut(roots[key])
    return roots


def _clusters_of(node, W):
    cl = {}
    for i in range(W):
        if node.pts[i]:
            a = np.asarray(node.pts[i], np.float64)
            cl[i] = (a.T @ a, a.sum(0), len(a))
    return cl


def voxelize_literal(scans, poses, voxel_size=1.0, eigen_ratio=EIGEN_RATIO_DEFAULT, layer_limit=LAYER_LIMIT, min_ps=MIN_PS):
    """tras_opt over the literal tree (:466-474): PLANE nodes, pushed when seen from >= 2 poses."""
    roots = build_tree_literal(scans, poses, voxel_size, eigen_ratio, layer_limit, min_ps)
    out = []

    def tras_opt(node, key, path):
        if node.state == "PLANE":
            _emit(out, key, path, node.layer, node.clusters, node.plane)
        else:
            for leaf in sorted(node.leaves):
                tras_opt(node.leaves[leaf], key, path + (leaf,))

    for key in sorted(roots):
        tras_opt(roots[key], key, ())
    return _pack(out)


def plane_lookup_literal(roots, X, voxel_size=1.0, layer_limit=LAYER_LIMIT):
    """recompute_local_planes, src/lvba_system.cpp:1529-1566, with OCTO_TREE_NODE::findCorrespondPoint (bavoxel.hpp:320-333):
    (n, d) of the PLANE node a world point falls in, zeros when there is none.  X: (n, 3).  Returns (n, 4)."""
    X = np.asarray(X, np.float64).reshape(-1, 3)
    out = np.zeros((len(X), 4))
    for pi, x in enumerate(X):
        if not np.all(np.isfinite(x)):                                     # :1532-1534
            continue
        loc = (x / voxel_size).astype(np.float32)                          # :1537-1541 (float; `-= 1.0f`)
        loc = np.where(loc < 0, loc - np.float32(1.0), loc).astype(np.float32)
        key = tuple(int(k) for k in np.trunc(loc).astype(np.int64))
        node = roots.get(key)
        if node is None:                                                   # :1546-1549
            continue
        while not (node.state == "PLANE" or node.layer >= layer_limit):    # findCorrespondPoint
            leaf, _ = _octant(x, node.centre)
            nxt = node.leaves.get(int(leaf))
            if nxt is None:
                break
            node = nxt
        if node.state != "PLANE":                                          # :1552-1554
            continue
        direct, centre = node.plane[2], node.plane[1]
        if not np.all(np.isfinite(direct)) or np.linalg.norm(direct) < 1e-6 or not np.all(np.isfinite(centre)):
            continue
        n = direct / np.linalg.norm(direct)                                # :1560-1563
        out[pi, :3] = n; out[pi, 3] = -n @ centre
    return out
