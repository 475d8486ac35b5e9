"""CPU restatement of the reference's track fusion — TEST INFRASTRUCTURE (only tests/ may import it):
    LvbaSystem::BuildTracksAndFuse3D   src/lvba_system.cpp:921-1263
with ComputeMeanReproj (:8-50) and TriangulateTrackDLT (:52-111) from oracle/track_oracle.py.

Literal where the reference is defined: adjacency in push_back order (:937-953), BFS with a FIFO queue from every unvisited
keypoint in (image, keypoint) order (:964-988), size / image-count gates (:989, :1001), first observation per image in member
order (:996-1000), depth candidate (:1016-1103), triangulation candidate (:1106-1159), choice (:1161-1199), release of a failed
component so that the scan tries it again from its next keypoint (:1199, :1203).
The reference iterates std::unordered_map<int,int> (image -> member) in three places; that order is unspecified and DECIDES the result (the
greedy view-angle filter keeps what it meets first).  `fuse(..., map_order=...)` takes it as a parameter: `libstdcxx_order` (default) = what GNU
libstdc++'s container does after the reference's reserve() / insert calls — LVBA_FUSE_ORDER_LIBSTDCXX, the default of the ABI; `ascending_order` =
ascending image id, the ABI's library-independent alternative (LVBA_FUSE_ORDER_ASCENDING).
PARITY: pinned against the reference's own source — under `libstdcxx_order` this function reproduces LvbaSystem::BuildTracksAndFuse3D, compiled
from src/lvba_system.cpp where it lies (oracle/ref_system_driver.cpp), track for track: tests/golden/ref_system.npz, tests/test_ref_system_pin.py;
`libstdcxx_order` itself is held against the real std::unordered_map there.
"""
from __future__ import annotations

from collections import deque

import numpy as np

from oracle import track_oracle as trk


_FAST_BKT = [2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13]
_PRIMES = [17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127, 137, 139, 149, 157, 167, 179, 193, 199,
           211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541, 577, 619, 661, 709, 761, 823, 887, 953, 1031, 1109, 1193, 1289,
           1381, 1493, 1613, 1741, 1879, 2029, 2179, 2357, 2549, 2753, 2971, 3209, 3469, 3739, 4027, 4349, 4703, 5087, 5503, 5953, 6427, 6949, 7517,
           8123, 8783, 9497, 10273, 11113, 12011, 12983, 14033, 15173, 16411, 17749, 19183, 20753, 22447, 24281, 26267, 28411, 30727, 33223, 35933,
           38873, 42043, 45481, 49201, 53201, 57557, 62233, 67307, 72817, 78779, 85229, 92203, 99733, 107897, 116731, 126271, 136607, 147793, 159871,
           172933, 187091, 202409, 218971, 236897, 256279, 277261, 299951, 324503, 351061, 379787, 410857, 444487, 480881, 520241, 562841, 608903,
           658753, 712697, 771049, 834181, 902483, 976369, 1056323, 1142821, 1236397, 1337629, 1447153, 1565659, 1693859, 1832561, 1982627, 2144977,
           2320627, 2510653, 2716249, 2938679, 3179303, 3439651, 3721303, 4026031, 4355707, 4712381, 5098259, 5515729, 5967347, 6456007, 6984629,
           7556579, 8175383, 8844859, 9569143, 10352717, 11200489, 12117689, 13109983, 14183539, 15345007, 16601593, 17961079, 19431899, 21023161,
           22744717, 24607243, 26622317, 28802401, 31160981, 33712729, 36473443, 39460231, 42691603, 46187573, 49969847, 54061849, 58488943,
           63278561, 68460391, 74066549, 80131819, 86693767, 93793069, 101473717, 109783337, 118773397, 128499677, 139022417, 150406843, 162723577,
           176048909, 190465427, 206062531, 222936881, 241193053, 260944219, 282312799, 305431229, 330442829, 357502601, 386778277, 418451333,
           452718089, 489790921, 529899637, 573292817, 620239453, 671030513, 725980837, 785430967, 849749479, 919334987, 994618837, 1076067617,
           1164186217, 1259520799, 1362662261, 1474249943, 1594975441, 1725587117, 1866894511, 2019773507]


def libstdcxx_order(reserve, keys):
    """Iteration order of a GNU libstdc++ std::unordered_map<int,int> after reserve(`reserve`) and the insertion of the distinct non-negative
    `keys` in this order — the order the reference's `for (auto& kv : map)` loops run in when it is built with g++ (the language leaves it
    open).  Restated from libstdc++'s published hashtable policy: reserve(n) gives the smallest bucket count >= n from the prime table (a
    fast table below 14), std::hash<int> is the identity, a node whose bucket is empty goes to the FRONT of the one global list, a node
    whose bucket is occupied goes to the front of that bucket's run.  No rehash happens afterwards: the reference never inserts more keys than
    it reserved.  Held against the real container on random inputs in tests/test_ref_system_pin.py."""
    n = int(reserve)
    if n == 0:
        nb = 2                                   # reserve(0) asks for room for one more element
    elif n < len(_FAST_BKT):
        nb = _FAST_BKT[n]
    else:
        nb = next(p for p in _PRIMES if p >= n)
    assert len(keys) <= max(nb, 1), "more keys than reserved: a rehash would reorder the list"
    order = []                                   # the global singly linked list, front first
    for k in keys:
        b = int(k) % nb
        pos = next((i for i, x in enumerate(order) if int(x) % nb == b), None)
        order.insert(0 if pos is None else pos, int(k))
    return order


def ascending_order(reserve, keys):
    """The library-independent alternative the ABI offers (LVBA_FUSE_ORDER_ASCENDING): images in ascending id."""
    return sorted(keys)


def fuse(kp_ptr, kp_uv, matches, cams, intr, kp_Xw, kp_valid, obser_thr=3, min_view_angle_deg=8.0, reproj_thr=3.0, depth_gate=0.12, map_order=None):
    """matches: (m, 4) int array (img_a, kp_a, img_b, kp_b) in the reference's visiting order.
    Returns a list of tracks {seed, obs (k,2), inlier (k,) bool, Xw, mean, source, kept (inlier positions in the order they were kept)} in the
    reference's track order.
    map_order(reserve, keys) -> keys in the order the reference's `for (auto& kv : unordered_map)` loops visit them (`keys` = image ids in insertion
    order, `reserve` = the argument of the map's reserve()).  None = `libstdcxx_order`, the order of a g++ build of the reference and the default of
    the ABI (LVBA_FUSE_ORDER_LIBSTDCXX): with it this function reproduces the reference's BuildTracksAndFuse3D track for track
    (tests/test_ref_system_pin.py).  `ascending_order` is the ABI's library-independent alternative."""
    if map_order is None:
        map_order = libstdcxx_order
    N = len(kp_ptr) - 1
    n_kp = int(kp_ptr[-1])
    img_of = np.repeat(np.arange(N), np.diff(kp_ptr))
    adj = [[] for _ in range(n_kp)]
    for ia, ka, ib, kb in np.asarray(matches, np.int64).reshape(-1, 4):
        if not (0 <= ia < N and 0 <= ib < N):
            continue
        if not (0 <= ka < kp_ptr[ia + 1] - kp_ptr[ia] and 0 <= kb < kp_ptr[ib + 1] - kp_ptr[ib]):
            continue
        a, b = int(kp_ptr[ia] + ka), int(kp_ptr[ib] + kb)
        adj[a].append(b); adj[b].append(a)
    cos_min = np.cos(min_view_angle_deg * np.pi / 180.0)
    state = np.full(n_kp, -1, np.int64)          # obs_to_track
    tracks = []
    cams = np.asarray(cams, np.float64).reshape(-1, 12)

    def centre(cam):
        R = cams[cam, :9].reshape(3, 3); t = cams[cam, 9:]
        return -R.T @ t

    def view_filter(cand, point_of):
        """cand: member positions in ascending image order; greedy filter (:1069-1095 / :1124-1150)"""
        kept, dirs = [], []
        for t in cand:
            cam = int(img_of[comp[t]])
            d = point_of(t) - centre(cam)
            nrm = np.linalg.norm(d)
            if nrm < 1e-6:
                continue
            d = d / nrm
            min_dot = min([float(d @ q) for q in dirs], default=1.0)
            if not dirs or min_dot <= cos_min:
                kept.append(t); dirs.append(d)
        return kept

    def sel_arrays(sel):
        g = [comp[t] for t in sel]
        return cams[[int(img_of[x]) for x in g]], np.asarray(kp_uv, np.float32).reshape(-1, 2)[g]

    for seed in range(n_kp):
        if state[seed] != -1:
            continue
        comp = []
        q = deque([seed]); state[seed] = -2
        while q:
            cur = q.popleft(); comp.append(cur)
            for nb in adj[cur]:
                if state[nb] == -1:
                    state[nb] = -2; q.append(nb)
        def release():
            for g in comp:
                state[g] = -1
        if len(comp) < obser_thr:
            release(); continue
        first = {}
        for t, g in enumerate(comp):
            first.setdefault(int(img_of[g]), t)
        if len(first) < obser_thr:
            release(); continue
        imgs = map_order(len(comp), list(first))                      # unique_id.reserve(component.size()), inserted in member order (:994-999)
        # ---- depth candidate
        depth_ok, Xd, mean_d, kept_d = False, np.zeros(3), np.inf, []
        valid = [t for t, g in enumerate(comp) if kp_valid[g]]
        if len(valid) >= obser_thr:
            Xa = np.asarray(kp_Xw[comp[valid[0]]], np.float64)
            best = {}
            for t in valid:
                if np.linalg.norm(np.asarray(kp_Xw[comp[t]], np.float64) - Xa) < depth_gate:
                    best.setdefault(int(img_of[comp[t]]), t)
            if len(best) >= obser_thr:
                n_inl = sum(1 for t in valid if np.linalg.norm(np.asarray(kp_Xw[comp[t]], np.float64) - Xa) < depth_gate)
                order = [best[i] for i in map_order(n_inl, list(best))]   # best_id.reserve(inliers.size()), inserted in inlier order (:1051-1056)
                Xd = np.zeros(3)
                for t in order:
                    Xd = Xd + np.asarray(kp_Xw[comp[t]], np.float64)
                Xd = Xd / float(len(order))
                kept_d = view_filter(order, lambda t: np.asarray(kp_Xw[comp[t]], np.float64))
                if len(kept_d) >= obser_thr:
                    cs, uv = sel_arrays(kept_d)
                    ok, m, _ = trk.mean_reproj(Xd, cs, uv, intr, obser_thr)
                    mean_d = m if ok else np.inf
                    depth_ok = ok and m <= reproj_thr
        # ---- triangulation candidate
        tri_ok, Xt, mean_t, kept_t = False, np.zeros(3), np.inf, []
        if len(first) >= 4:
            order = [first[i] for i in imgs]
            cs, uv = sel_arrays(order)
            ok, Xs, _, _ = trk.triangulate_dlt(cs, uv, intr)
            if ok:
                kept_t = view_filter(order, lambda t: Xs)
                if len(kept_t) >= 4:
                    cs, uv = sel_arrays(kept_t)
                    ok2, Xt, m2, _ = trk.triangulate_dlt(cs, uv, intr)
                    if ok2:
                        mean_t = m2
                        tri_ok = m2 <= reproj_thr
        if depth_ok and tri_ok:
            src = 2 if mean_t < mean_d else 1
        elif tri_ok:
            src = 2
        elif depth_ok:
            src = 1
        else:
            release(); continue
        X, mean, kept = (Xt, mean_t, kept_t) if src == 2 else (Xd, mean_d, kept_d)
        if not np.all(np.isfinite(X)) or np.all(np.abs(X) <= 1e-12):
            release(); continue
        tid = len(tracks)
        inl = np.zeros(len(comp), bool); inl[kept] = True
        obs = np.array([[int(img_of[g]), int(g - kp_ptr[img_of[g]])] for g in comp], np.int32)
        tracks.append(dict(seed=seed, obs=obs, inlier=inl, Xw=np.array(X, np.float64), mean=float(mean), source=src, kept=list(kept)))
        for g in comp:
            state[g] = tid
    return tracks
