"""Deterministic synthetic LiDAR-visual BA problems.

TEST INFRASTRUCTURE — not shipped code.  Only tests/, bench.py and
__graft_entry__.smoke() import this module.  It only *generates inputs*; it
contains none of the solver arithmetic.

The generator follows the distributions the survey fixes for the named configs
(A 50/2k/1k, B 500/50k/20k, C 2000/200k/100k, E 5000/500k/300k):

* trajectory: poses every 0.5 m of arc length on a closed 3-D Lissajous loop in
  a 200 x 200 x 10 m box, body x-axis along the tangent, roll/pitch N(0, 2deg);
  initial estimate = GT * Exp(N(0, 0.005 rad)), p + N(0, 0.03 m);
* plane voxels: centre pose c ~ U, K ~ U{2..12} distinct poses from
  [c-15, c+15]; plane centre = p_c + U(-10,10)^3; n_i ~ U{8..40} points per
  observing pose = in-plane U(-.25,.25)^2 + N(0, 0.01 m) normal noise, mapped
  to the GT body frame, rounded to float32 and accumulated into the BALM
  PointCluster (P = sum p p^T, v = sum p, N)   [reference: tools.hpp:407-466];
* tracks: centre camera c ~ U, L ~ U{3..8} distinct cameras from [c-10, c+10]
  that actually see the landmark (z >= 0.5 m, inside 640x512), landmark 3-20 m
  in front of camera c, pixel noise N(0, 0.5 px) stored as float32, landmark
  plane n ~ S^2, d = -n . X_gt, X0 = X_gt + N(0, 0.05 m).
  Intrinsics / extrinsics = reference config/config.yaml:1-20 (x scale 0.5,
  dataset_io.cpp:59-62); camera pose from body pose as lvba_system.cpp:486-505.

RNG: numpy Philox (counter based) keyed by seed = 20260923 + config index.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (index, poses, voxels, tracks)
    "A": (0, 50, 2_000, 1_000),
    "B": (1, 500, 50_000, 20_000),
    "C": (2, 2_000, 200_000, 100_000),
    "E": (4, 5_000, 500_000, 300_000),
}
BASE_SEED = 20260923

# config/config.yaml:5-12 scaled by 0.5 (dataset_io.cpp:59-62)
INTR = np.array([646.78472, 646.65775, 313.456795, 261.399612,
                 -0.076160, 0.123001, -0.00113, 0.000251], dtype=np.float64)
IMG_W, IMG_H = 640, 512
# config/config.yaml:15-20
RCL = np.array([[0.00610193, -0.999863, -0.0154172],
                [-0.00615449, 0.0153796, -0.999863],
                [0.999962, 0.00619598, -0.0060598]], dtype=np.float64)
PCL = np.array([0.0194384, 0.104689, -0.0251952], dtype=np.float64)
TIL = np.array([0.04165, 0.02326, -0.0284], dtype=np.float64)
SIGMA_PX = 0.5
SIGMA_PLANE = 0.01


def _orthonormalize(R):
    u, _, vt = np.linalg.svd(R)
    return u @ vt


def so3_exp(w):
    """Rodrigues, batched (n,3)->(n,3,3)."""
    w = np.atleast_2d(w)
    th = np.linalg.norm(w, axis=1)
    K = np.zeros((w.shape[0], 3, 3))
    safe = np.where(th > 0, th, 1.0)
    a = w / safe[:, None]
    K[:, 0, 1], K[:, 0, 2] = -a[:, 2], a[:, 1]
    K[:, 1, 0], K[:, 1, 2] = a[:, 2], -a[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -a[:, 1], a[:, 0]
    I = np.eye(3)[None]
    return I + np.sin(th)[:, None, None] * K + (1 - np.cos(th))[:, None, None] * (K @ K)


def rot_to_quat_wxyz(R):
    """Batched rotation matrix -> unit quaternion (w,x,y,z), w >= 0."""
    R = np.asarray(R)
    n = R.shape[0]
    q = np.empty((n, 4))
    for k in range(n):
        m = R[k]
        tr = m[0, 0] + m[1, 1] + m[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            q[k] = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q[k] = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q[k] = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q[k] = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    return q


def make_trajectory(n, rng):
    """GT body poses (R_wb, p) every 0.5 m along a closed Lissajous loop."""
    s = np.linspace(0.0, 2 * np.pi, 400_001)
    curve = np.stack([100 * np.sin(s), 100 * np.sin(2 * s + 0.7), 5 * np.sin(3 * s + 0.3)], 1)
    seg = np.linalg.norm(np.diff(curve, axis=0), axis=1)
    arclen = np.concatenate([[0.0], np.cumsum(seg)])
    total = arclen[-1]
    want = (np.arange(n) * 0.5) % total
    sp = np.interp(want, arclen, s)
    p = np.stack([100 * np.sin(sp), 100 * np.sin(2 * sp + 0.7), 5 * np.sin(3 * sp + 0.3)], 1)
    tan = np.stack([100 * np.cos(sp), 200 * np.cos(2 * sp + 0.7), 15 * np.cos(3 * sp + 0.3)], 1)
    yaw = np.arctan2(tan[:, 1], tan[:, 0])
    roll = rng.normal(0, np.deg2rad(2.0), n)
    pitch = rng.normal(0, np.deg2rad(2.0), n)
    cz, sz = np.cos(yaw), np.sin(yaw)
    cy, sy = np.cos(pitch), np.sin(pitch)
    cx, sx = np.cos(roll), np.sin(roll)
    Rz = np.zeros((n, 3, 3)); Ry = np.zeros((n, 3, 3)); Rx = np.zeros((n, 3, 3))
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = cz, -sz, sz, cz, 1
    Ry[:, 0, 0], Ry[:, 0, 2], Ry[:, 2, 0], Ry[:, 2, 2], Ry[:, 1, 1] = cy, sy, -sy, cy, 1
    Rx[:, 1, 1], Rx[:, 1, 2], Rx[:, 2, 1], Rx[:, 2, 2], Rx[:, 0, 0] = cx, -sx, sx, cx, 1
    R = Rz @ Ry @ Rx
    return R, p


def body_to_cam(R_wb, p):
    """(R_cw, t_cw) from the body pose, reference lvba_system.cpp:486-505."""
    Rli = np.eye(3)            # extrinsic_R = I
    tli = -Rli @ TIL
    Rci = RCL @ Rli
    tci = RCL @ tli + PCL
    R_cw = Rci[None] @ np.transpose(R_wb, (0, 2, 1))
    t_cw = tci[None] - np.einsum("nij,nj->ni", R_cw, p)
    return R_cw, t_cw


def project(R_cw, t_cw, X, intr=INTR):
    """Brown-Conrady projection (utils.hpp:61-111 forward model), batched."""
    Xc = np.einsum("...ij,...j->...i", R_cw, X) + t_cw
    z = Xc[..., 2]
    zs = np.where(np.abs(z) > 1e-12, z, 1e-12)
    xn, yn = Xc[..., 0] / zs, Xc[..., 1] / zs
    fx, fy, cx, cy, k1, k2, p1, p2 = intr
    r2 = xn * xn + yn * yn
    rad = 1 + k1 * r2 + k2 * r2 * r2
    xd = xn * rad + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
    yd = yn * rad + p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn
    return np.stack([fx * xd + cx, fy * yd + cy], -1), z


def _pick_distinct(rng, centre, half, n_items, counts):
    """For every row pick counts[r] distinct indices from [c-half, c+half] ∩ [0,n_items)."""
    width = 2 * half + 1
    cand = centre[:, None] + np.arange(-half, half + 1)[None, :]
    ok = (cand >= 0) & (cand < n_items)
    keys = rng.random(cand.shape)
    keys[~ok] = 2.0
    order = np.argsort(keys, axis=1)
    nvalid = ok.sum(1)
    counts = np.minimum(counts, nvalid)
    return cand, order, counts, width


def make_lidar(n_poses, n_vox, R_gt, p_gt, rng, k_lo=2, k_hi=12, chunk=60_000, half=15):
    """half: a voxel is seen from poses within +-half of its centre pose (15 = the band of the headline configs; larger
    values give the loop-closure couplings that widen the envelope)."""
    c = rng.integers(0, n_poses, n_vox)
    K = rng.integers(k_lo, k_hi + 1, n_vox)
    cand, order, K, _ = _pick_distinct(rng, c, half, n_poses, K)
    K = np.maximum(K, 2) if n_poses >= 2 else K
    vox_ptr = np.zeros(n_vox + 1, np.int64)
    np.cumsum(K, out=vox_ptr[1:])
    nnz = int(vox_ptr[-1])
    row = np.repeat(np.arange(n_vox), K)
    col = np.arange(nnz) - vox_ptr[row]
    pose_idx = np.take_along_axis(cand, order, 1)[row, col].astype(np.int32)
    # ascending pose order inside each voxel (reference iterates i = 0..W-1)
    key = row.astype(np.int64) * (n_poses + 1) + pose_idx
    pose_idx = pose_idx[np.argsort(key, kind="stable")]

    centre = p_gt[c] + rng.uniform(-10, 10, (n_vox, 3))
    nrm = rng.normal(size=(n_vox, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    helper = np.where(np.abs(nrm[:, :1]) < 0.9, np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]))
    e1 = np.cross(nrm, helper); e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(nrm, e1)

    npts = rng.integers(8, 41, nnz)
    clusters = np.empty((nnz, 10), np.float64)
    for s0 in range(0, nnz, chunk):
        s1 = min(nnz, s0 + chunk)
        m = s1 - s0
        vr = row[s0:s1]
        a = rng.uniform(-0.25, 0.25, (m, 40)); b = rng.uniform(-0.25, 0.25, (m, 40))
        e = rng.normal(0, 0.01, (m, 40))
        pw = (centre[vr][:, None, :] + a[..., None] * e1[vr][:, None, :]
              + b[..., None] * e2[vr][:, None, :] + e[..., None] * nrm[vr][:, None, :])
        pi = pose_idx[s0:s1]
        pb = np.einsum("nji,nkj->nki", R_gt[pi], pw - p_gt[pi][:, None, :])   # R^T (pw - p)
        pb = pb.astype(np.float32).astype(np.float64)
        mask = (np.arange(40)[None, :] < npts[s0:s1, None]).astype(np.float64)
        pbm = pb * mask[..., None]
        P = np.einsum("nki,nkj->nij", pbm, pb)
        clusters[s0:s1, 0] = P[:, 0, 0]; clusters[s0:s1, 1] = P[:, 0, 1]; clusters[s0:s1, 2] = P[:, 0, 2]
        clusters[s0:s1, 3] = P[:, 1, 1]; clusters[s0:s1, 4] = P[:, 1, 2]; clusters[s0:s1, 5] = P[:, 2, 2]
        clusters[s0:s1, 6:9] = pbm.sum(1)
        clusters[s0:s1, 9] = npts[s0:s1]
    return vox_ptr, pose_idx, clusters


def make_visual(n_cams, n_tracks, R_cw_gt, t_cw_gt, rng):
    intr = INTR
    obs_cam_rows, obs_uv_rows, L_all = [None] * n_tracks, [None] * n_tracks, np.zeros(n_tracks, np.int64)
    X_gt = np.zeros((n_tracks, 3))
    todo = np.arange(n_tracks)
    half = 10
    offs = np.arange(-half, half + 1)
    rounds = 0
    cam_lists = np.full((n_tracks, 8), -1, np.int64)
    uv_lists = np.zeros((n_tracks, 8, 2), np.float32)
    while todo.size:
        rounds += 1
        if rounds > 200:
            raise RuntimeError("track sampling did not converge")
        m = todo.size
        c = rng.integers(0, n_cams, m)
        L = rng.integers(3, 9, m)
        px = np.stack([rng.uniform(20, IMG_W - 20, m), rng.uniform(20, IMG_H - 20, m)], 1)
        depth = rng.uniform(3.0, 20.0, m)
        xn = (px[:, 0] - intr[2]) / intr[0]; yn = (px[:, 1] - intr[3]) / intr[1]
        Xc = np.stack([xn * depth, yn * depth, depth], 1)
        Xw = np.einsum("nji,nj->ni", R_cw_gt[c], Xc - t_cw_gt[c])
        cand = c[:, None] + offs[None, :]
        ok = (cand >= 0) & (cand < n_cams)
        candc = np.clip(cand, 0, n_cams - 1)
        uv, z = project(R_cw_gt[candc], t_cw_gt[candc], Xw[:, None, :])
        noise = rng.normal(0, 0.5, uv.shape)
        uvn = (uv + noise).astype(np.float32)
        vis = ok & (z >= 0.5) & (uv[..., 0] >= 0) & (uv[..., 0] < IMG_W) & (uv[..., 1] >= 0) & (uv[..., 1] < IMG_H)
        keys = rng.random(cand.shape); keys[~vis] = 2.0
        order = np.argsort(keys, axis=1)
        good = vis.sum(1) >= L
        gi = np.nonzero(good)[0]
        for r in gi:                       # small python loop over accepted rows only
            sel = np.sort(order[r, :L[r]])
            tid = todo[r]
            cam_lists[tid, :L[r]] = cand[r, sel]
            uv_lists[tid, :L[r]] = uvn[r, sel]
            L_all[tid] = L[r]
            X_gt[tid] = Xw[r]
        todo = todo[~good]
    obs_ptr = np.zeros(n_tracks + 1, np.int64)
    np.cumsum(L_all, out=obs_ptr[1:])
    msk = np.arange(8)[None, :] < L_all[:, None]
    obs_cam = cam_lists[msk].astype(np.int32)
    obs_uv = uv_lists[msk].astype(np.float32)
    n = rng.normal(size=(n_tracks, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = -np.einsum("ni,ni->n", n, X_gt)
    plane_nd = np.concatenate([n, d[:, None]], 1)
    X0 = X_gt + rng.normal(0, 0.05, X_gt.shape)
    return obs_ptr, obs_cam, obs_uv, X0, X_gt, plane_nd


def make_problem(n_poses, n_vox, n_tracks, seed, lidar=True, visual=True, half=15):
    """Returns a dict of numpy arrays in exactly the layout the C-ABI takes
    (include/lvba_b200.h): CSR voxel->pose clusters; CSR track->camera obs."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    R_gt, p_gt = make_trajectory(n_poses, rng)
    dR = so3_exp(rng.normal(0, 0.005, (n_poses, 3)))
    R0 = R_gt @ dR
    p0 = p_gt + rng.normal(0, 0.03, (n_poses, 3))
    out = {"n_poses": n_poses, "seed": seed}
    out["poses_gt"] = np.concatenate([R_gt.reshape(n_poses, 9), p_gt], 1)
    out["poses"] = np.concatenate([R0.reshape(n_poses, 9), p0], 1)
    if lidar and n_vox:
        vp, pi, cl = make_lidar(n_poses, n_vox, R_gt, p_gt, rng, half=half)
        out.update(vox_ptr=vp, pose_idx=pi, clusters=cl, n_vox=n_vox)
    if visual and n_tracks:
        Rc_gt, tc_gt = body_to_cam(R_gt, p_gt)
        Rc0, tc0 = body_to_cam(R0, p0)
        op, oc, ouv, X0, Xgt, pl = make_visual(n_poses, n_tracks, Rc_gt, tc_gt, rng)
        out.update(obs_ptr=op, obs_cam=oc, obs_uv=ouv, X=X0, X_gt=Xgt, plane_nd=pl,
                   q=rot_to_quat_wxyz(Rc0), t=tc0.copy(), q_gt=rot_to_quat_wxyz(Rc_gt), t_gt=tc_gt.copy(),
                   intr=INTR.copy(), sigma_px=SIGMA_PX, sigma_plane=SIGMA_PLANE, n_tracks=n_tracks)
    return out


def make_config(name, **kw):
    idx, n, v, t = CONFIGS[name]
    return make_problem(n, v, t, BASE_SEED + idx, **kw)


def make_window_problem(window_sizes, vox_per_window, seed, k_hi=8):
    """Consecutive independent window-BA problems (LvbaSystem::runWindowBA, src/lvba_system.cpp:232-302) laid out
    the way lvba_lidar_lm_batch takes them: concatenated poses, pose indices into the concatenation, every voxel
    inside one window.  `vox_per_window` may be an int or a per-window list (0 voxels = a window that gets skipped).
    Returns dict(win_ptr, vox_ptr, pose_idx, clusters, poses, windows=[per-window single problems])."""
    sizes = list(window_sizes)
    if np.isscalar(vox_per_window):
        vox_per_window = [int(vox_per_window)] * len(sizes)
    win_ptr = np.zeros(len(sizes) + 1, np.int32)
    win_ptr[1:] = np.cumsum(sizes)
    vp_all, pi_all, cl_all, ps_all, windows = [np.zeros(1, np.int64)], [], [], [], []
    nnz = 0
    for w, (nw, nv) in enumerate(zip(sizes, vox_per_window)):
        p = make_problem(max(nw, 2), max(nv, 1), 0, seed=seed + 17 * w, visual=False)
        poses = p["poses"][:nw]
        if nv > 0 and nw >= 2:
            keep = [a for a in range(len(p["vox_ptr"]) - 1) if p["pose_idx"][p["vox_ptr"][a + 1] - 1] < nw][:nv]
        else:
            keep = []
        vp = [0]; pi = []; cl = []
        for a in keep:
            lo, hi = p["vox_ptr"][a], p["vox_ptr"][a + 1]
            pi.append(p["pose_idx"][lo:hi]); cl.append(p["clusters"][lo:hi]); vp.append(vp[-1] + hi - lo)
        pi = np.concatenate(pi).astype(np.int32) if pi else np.zeros(0, np.int32)
        cl = np.concatenate(cl) if cl else np.zeros((0, 10))
        vp = np.asarray(vp, np.int64)
        windows.append(dict(vox_ptr=vp, pose_idx=pi, clusters=cl, poses=poses.copy()))
        vp_all.append(vp[1:] + nnz); pi_all.append(pi + win_ptr[w]); cl_all.append(cl); ps_all.append(poses)
        nnz += int(vp[-1])
    return dict(win_ptr=win_ptr, vox_ptr=np.concatenate(vp_all), pose_idx=np.concatenate(pi_all).astype(np.int32),
                clusters=np.concatenate(cl_all), poses=np.concatenate(ps_all), windows=windows)


def make_scan_scene(seed, W=5, n_per_scan=2500):
    """A corner of a room (floor + two walls, 1 cm noise) and some clutter, seen from W nearby poses; float32 body frame."""
    rng = np.random.default_rng(seed)
    poses = np.zeros((W, 12))
    scans = []
    for i in range(W):
        R = so3_exp(rng.normal(0, 0.05, (1, 3)))[0]
        p = np.array([0.4 * i - 1.0, 0.3 * np.sin(i), 0.1 * i]) + rng.normal(0, 0.02, 3)
        poses[i, :9] = R.ravel(); poses[i, 9:] = p
        k = n_per_scan
        u = rng.uniform(-3, 3, (k, 2)); kind = rng.integers(0, 4, k)
        w = np.zeros((k, 3))
        w[kind == 0] = np.column_stack([u[kind == 0], -1.2 + rng.normal(0, 0.01, (kind == 0).sum())])               # floor z = -1.2
        w[kind == 1] = np.column_stack([np.full((kind == 1).sum(), 2.6) + rng.normal(0, 0.01, (kind == 1).sum()), u[kind == 1]])   # wall x = 2.6
        w[kind == 2] = np.column_stack([u[kind == 2][:, 0], np.full((kind == 2).sum(), -2.4) + rng.normal(0, 0.01, (kind == 2).sum()), u[kind == 2][:, 1]])
        w[kind == 3] = rng.uniform(-3, 3, ((kind == 3).sum(), 3))                                                  # clutter
        scans.append(((w - p) @ R).astype(np.float32))                                                            # R^T (w - p)
    return scans, poses


def make_depth_scene(seed, F=6, n_per_scan=1500, M=4, width=160, height=120):
    """Scans of make_scan_scene with frame timestamps 0.25 s apart, and M pinhole cameras (mild Brown-Conrady distortion)
    riding on interpolated frame poses, looking at the x = 2.6 wall / the floor.  Returns a dict for oracle.depth_oracle."""
    scans, poses = make_scan_scene(seed, W=F, n_per_scan=n_per_scan)
    rng = np.random.default_rng(seed + 1000)
    frame_ts = 100.0 + 0.25 * np.arange(F)
    image_ts = np.sort(rng.uniform(frame_ts[0] - 0.2, frame_ts[-1] + 0.2, M))
    cams = np.zeros((M, 12))
    for k in range(M):
        f = int(np.clip(np.searchsorted(frame_ts, image_ts[k]), 0, F - 1))
        Rwi = poses[f, :9].reshape(3, 3); pwi = poses[f, 9:]
        # camera axes in the body frame: z_c forward = body +x (towards the wall), x_c = body -y, y_c = body -z, plus a small tilt
        Rci = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]]) @ so3_exp(rng.normal(0, 0.15, (1, 3)))[0]
        tci = rng.normal(0, 0.05, 3)
        Rcw = Rci @ Rwi.T
        cams[k, :9] = Rcw.ravel(); cams[k, 9:] = -Rcw @ pwi + tci                       # lvba_system.cpp:861-862
    intr = np.array([0.6 * width, 0.6 * width, 0.5 * width - 0.3, 0.5 * height + 0.2, -0.08, 0.01, 5e-4, -3e-4])
    return dict(scans=scans, poses=poses, frame_ts=frame_ts, cams=cams, image_ts=image_ts, intr=intr, width=width, height=height)
