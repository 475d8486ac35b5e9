"""CPU restatement of the reference's depth rendering (first half) — TEST INFRASTRUCTURE; nothing in
the product imports it.
PARITY: pinned against the reference's own source (src/lvba_system.cpp and include/utils.hpp compiled where they lie): depth images, projection,
undistortion and the depth-candidate chain BIT FOR BIT (tests/golden/ref_system.npz, ref_balm.npz; tests/test_ref_system_pin.py, test_ref_pin.py).

Follows:
  buildGridMapFromOptimized   src/lvba_system.cpp:1266-1338   world points bucketed into 0.5 m voxels (float key, `-= 1.0f`
                                                              for negatives, int64 truncation); per-frame voxel sets; per
                                                              image the union over the frames within +-0.5 s of its timestamp
  generateDepthWithVoxel      src/lvba_system.cpp:835-919     every point of every listed voxel: pC = Rcw pW + tcw, Z < 1e-3
                                                              skipped, Brown-Conrady projection, (int) truncation of the pixel,
                                                              z-buffer `if (d == 0 || Z < d) d = (float)Z`
  projectCameraToPixel        include/utils.hpp:183-197 ;  distortNormalized  include/utils.hpp:169-181

Two implementations the tests compare: render_literal() (dicts and loops, as the reference) and render() (numpy group-by).
The z-buffer result is float(min Z) per pixel whatever the visiting order (float() is monotone), so the images are
compared EXACTLY.
"""
from __future__ import annotations

import numpy as np

GRID_VOXEL = 0.5       # `const double vox = 0.5;`  :1277
HALF_WINDOW = 0.5      # `const double half_w = 0.5;` :1299


def grid_keys(world, voxel_size=GRID_VOXEL):
    """:1287-1291 — loc = (float)(pw / vox); if (loc < 0) loc -= 1.0f; (int64) truncation."""
    loc = (world / voxel_size).astype(np.float32)
    loc = np.where(loc < 0, loc - np.float32(1.0), loc).astype(np.float32)
    return np.trunc(loc).astype(np.int64)


def world_points(scans, poses):
    """pvec_tran = R * pvec_orig + t per frame (:1282-1286); coefficient order (r0 x + r1 y) + r2 z, then + t."""
    out = []
    for s, pose in zip(scans, poses):
        p = np.asarray(s, np.float32).reshape(-1, 3).astype(np.float64)
        R = pose[:9].reshape(3, 3); t = pose[9:]
        w = np.empty_like(p)
        for k in range(3):
            w[:, k] = ((R[k, 0] * p[:, 0] + R[k, 1] * p[:, 1]) + R[k, 2] * p[:, 2]) + t[k]
        out.append(w)
    return out


def frame_window(frame_ts, t_img, half_window=HALF_WINDOW):
    """:1316-1319 — lower_bound(t_img - half_w) .. upper_bound(t_img + half_w) over the ascending frame timestamps."""
    return int(np.searchsorted(frame_ts, t_img - half_window, side="left")), int(np.searchsorted(frame_ts, t_img + half_window, side="right"))


def project(pc, intr):
    """projectCameraToPixel (utils.hpp:183-197) on rows of camera-frame points; returns (ok, uu, vv)."""
    fx, fy, cx, cy, k1, k2, p1, p2 = intr
    Z = pc[:, 2]
    with np.errstate(all="ignore"):
        ok = np.all(np.isfinite(pc), axis=1) & (Z > 1e-12)
        x = pc[:, 0] / Z; y = pc[:, 1] / Z
        r2 = x * x + y * y
        r4 = r2 * r2
        radial = (1.0 + k1 * r2) + k2 * r4
        x_tan = ((2.0 * p1) * x) * y + p2 * (r2 + (2.0 * x) * x)
        y_tan = p1 * (r2 + (2.0 * y) * y) + ((2.0 * p2) * x) * y
        xd = x * radial + x_tan; yd = y * radial + y_tan
        ok &= np.isfinite(xd) & np.isfinite(yd)
        uu = fx * xd + cx; vv = fy * yd + cy
        ok &= np.isfinite(uu) & np.isfinite(vv)
    return ok, uu, vv


def _splat(depth, pw, cam, intr, width, height):
    """The inner loop of generateDepthWithVoxel (:885-901) for an array of world points."""
    R = cam[:9].reshape(3, 3); t = cam[9:]
    pc = np.empty_like(pw)
    for k in range(3):
        pc[:, k] = ((R[k, 0] * pw[:, 0] + R[k, 1] * pw[:, 1]) + R[k, 2] * pw[:, 2]) + t[k]
    Z = pc[:, 2]
    keep = ~(Z < 1e-3)                                                  # `if (Z < 1e-3) continue;`  (NaN passes, dies below)
    ok, uu, vv = project(pc, intr)
    keep &= ok
    with np.errstate(all="ignore"):
        big = (np.abs(uu) < 2.0 ** 31) & (np.abs(vv) < 2.0 ** 31)      # the (int) cast is only defined in range
    keep &= big
    u = np.trunc(np.where(keep, uu, 0.0)).astype(np.int64); v = np.trunc(np.where(keep, vv, 0.0)).astype(np.int64)
    keep &= (u >= 0) & (u < width) & (v >= 0) & (v < height)
    zf = Z[keep].astype(np.float32)
    flat = v[keep] * width + u[keep]
    buf = np.full(width * height, np.inf, np.float32)
    np.minimum.at(buf, flat, zf)
    cur = depth.reshape(-1)
    upd = np.isfinite(buf) & ((cur == 0) | (buf < cur))
    cur[upd] = buf[upd]


def render(scans, poses, frame_ts, cams, image_ts, intr, width, height, voxel_size=GRID_VOXEL, half_window=HALF_WINDOW):
    """Vectorised restatement.  scans: list of (n_i, 3) float32 body-frame clouds; poses (F, 12); frame_ts (F,) ascending;
    cams (M, 12) = Rcw row-major, tcw; image_ts (M,) (NaN = unparsable name -> empty image).  Returns (M, H, W) float32."""
    worlds = world_points(scans, poses)
    allw = np.concatenate(worlds) if worlds else np.zeros((0, 3))
    keys = grid_keys(allw, voxel_size)
    frame_of = np.concatenate([np.full(len(w), i, np.int64) for i, w in enumerate(worlds)]) if worlds else np.zeros(0, np.int64)
    uniq, inv = np.unique(keys, axis=0, return_inverse=True) if len(keys) else (np.zeros((0, 3), np.int64), np.zeros(0, np.int64))
    inv = inv.reshape(-1)
    out = np.zeros((len(cams), height, width), np.float32)
    for k in range(len(cams)):
        if not np.isfinite(image_ts[k]):
            continue
        fl, fr = frame_window(frame_ts, image_ts[k], half_window)
        touched = np.zeros(len(uniq), bool)
        touched[inv[(frame_of >= fl) & (frame_of < fr)]] = True
        _splat(out[k], allw[touched[inv]], cams[k], intr, width, height)
    return out


def render_literal(scans, poses, frame_ts, cams, image_ts, intr, width, height, voxel_size=GRID_VOXEL, half_window=HALF_WINDOW):
    """The reference's containers, literally (small inputs): grid_map_ dict, per-frame voxel sets, per-image union, point loop."""
    worlds = world_points(scans, poses)
    grid = {}
    per_frame = []
    for w in worlds:
        ks = grid_keys(w, voxel_size)
        s = set()
        for p, k in zip(w, ks):
            kk = (int(k[0]), int(k[1]), int(k[2]))
            grid.setdefault(kk, []).append(p)
            s.add(kk)
        per_frame.append(s)
    fx, fy, cx, cy, k1, k2, p1, p2 = intr
    out = np.zeros((len(cams), height, width), np.float32)
    for k in range(len(cams)):
        if not np.isfinite(image_ts[k]):
            continue
        fl, fr = frame_window(frame_ts, image_ts[k], half_window)
        voxels = set()
        for f in range(fl, fr):
            voxels |= per_frame[f]
        R = cams[k][:9].reshape(3, 3); t = cams[k][9:]
        depth = out[k]
        for vk in sorted(voxels):
            for pw in grid[vk]:
                pc = np.array([((R[i, 0] * pw[0] + R[i, 1] * pw[1]) + R[i, 2] * pw[2]) + t[i] for i in range(3)])
                Z = pc[2]
                if Z < 1e-3:
                    continue
                ok, uu, vv = project(pc[None, :], intr)
                if not ok[0]:
                    continue
                u = int(uu[0]); v = int(vv[0])                          # truncation toward zero: (-1, 0) maps to pixel 0
                if u < 0 or u >= width or v < 0 or v >= height:
                    continue
                d = depth[v, u]
                if d == 0 or Z < d:
                    depth[v, u] = np.float32(Z)
    return out


def fetch_depth_bilinear(depth, u, v):
    """fetchDepthBilinear, include/utils.hpp:246-275, CV_32FC1 branch; everything in float32.  Returns (ok, d)."""
    f = np.float32
    h, w = depth.shape
    u = f(u); v = f(v)
    if u < f(0) or v < f(0) or u >= f(w - 1) or v >= f(h - 1) or not (u == u and v == v):
        return False, f(0)
    x = int(np.floor(u)); y = int(np.floor(v))
    du = f(u - f(x)); dv = f(v - f(y))
    d00, d10, d01, d11 = depth[y, x], depth[y, x + 1], depth[y + 1, x], depth[y + 1, x + 1]
    if d00 <= 0 or d10 <= 0 or d01 <= 0 or d11 <= 0:
        return False, f(0)
    omu = f(f(1) - du); omv = f(f(1) - dv)
    d = f(f(f(f(omu * omv) * d00) + f(f(du * omv) * d10)) + f(f(omu * dv) * d01))
    d = f(d + f(f(du * dv) * d11))
    return bool(d > 0), d


def undistort_pixel(intr, u, v):
    """undistortPixelToNormalized, include/utils.hpp:207-233 (8 fixed-point iterations)."""
    fx, fy, cx, cy, k1, k2, p1, p2 = [float(a) for a in intr]
    if not (np.isfinite(u) and np.isfinite(v)) or abs(fx) < 1e-12 or abs(fy) < 1e-12:
        return False, 0.0, 0.0
    xd = (u - cx) / fx; yd = (v - cy) / fy
    xu, yu = xd, yd
    for _ in range(8):
        r2 = xu * xu + yu * yu
        r4 = r2 * r2
        radial = (1.0 + k1 * r2) + k2 * r4
        if abs(radial) < 1e-12 or not np.isfinite(radial):
            return False, 0.0, 0.0
        x_tan = ((2.0 * p1) * xu) * yu + p2 * (r2 + (2.0 * xu) * xu)
        y_tan = p1 * (r2 + (2.0 * yu) * yu) + ((2.0 * p2) * xu) * yu
        xu = (xd - x_tan) / radial
        yu = (yd - y_tan) / radial
        if not (np.isfinite(xu) and np.isfinite(yu)):
            return False, 0.0, 0.0
    return True, xu, yu


def backproject(depth_images, cams, intr, kp_ptr, kp_uv):
    """The depth-candidate loop of BuildTracksAndFuse3D (src/lvba_system.cpp:1020-1038) for every keypoint of every image:
    fetchDepthBilinear -> backProjectPixelDepthDistorted (utils.hpp:235-243) -> camToWorld (:277-283).
    kp_ptr (M+1,), kp_uv (n, 2) float32.  Returns (Xw (n, 3), valid (n,) uint8)."""
    n = len(kp_uv)
    Xw = np.zeros((n, 3)); valid = np.zeros(n, np.uint8)
    for k in range(len(cams)):
        R = cams[k][:9].reshape(3, 3); t = cams[k][9:]
        for q in range(int(kp_ptr[k]), int(kp_ptr[k + 1])):
            u, v = np.float32(kp_uv[q, 0]), np.float32(kp_uv[q, 1])
            ok, d = fetch_depth_bilinear(depth_images[k], u, v)
            if not ok or d <= 0:
                continue
            dd = float(d)
            ok, x, y = undistort_pixel(intr, float(u), float(v))
            if not ok:
                continue
            Xc = np.array([x * dd, y * dd, dd])
            if not np.all(np.isfinite(Xc)):
                continue
            for i in range(3):
                twc = -((R[0, i] * t[0] + R[1, i] * t[1]) + R[2, i] * t[2])
                Xw[q, i] = ((R[0, i] * Xc[0] + R[1, i] * Xc[1]) + R[2, i] * Xc[2]) + twc
            valid[q] = 1
    return Xw, valid
