"""ctypes binding of oracle/_ref/libbalm_ref.so — the reference's own hot-path sources compiled where they lie
(oracle/ref_driver.cpp, `make -C oracle ref`).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Exists only where /root/reference exists (this container, not the GPU box):
`available()` says whether the library is there; tests that need it skip otherwise, and the fixtures it wrote
(tests/golden/ref_*.npz, by tests/golden/make_golden_ref.py) are what travels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libbalm_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.ref_lidar_hessian.restype = C.c_int64
        _lib.ref_lidar_residual.restype = C.c_double
        _lib.ref_lidar_damping_iter.restype = C.c_int64
        _lib.ref_map_create.restype = C.c_void_p
        _lib.ref_map_num_voxels.restype = C.c_int64
        _lib.ref_map_num_slots.restype = C.c_int64
        _lib.ref_anchor_cloud.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _csr(vox_ptr, pose_idx, clusters, poses):
    return (np.ascontiguousarray(vox_ptr, np.int64), np.ascontiguousarray(pose_idx, np.int32),
            np.ascontiguousarray(clusters, np.float64), np.ascontiguousarray(poses, np.float64))


def lidar_hessian(vox_ptr, pose_idx, clusters, poses, threads=False):
    """VOX_HESS::acc_evaluate2 in one call (threads=False: residual = sum lambda_0) or BALM2::divide_thread
    (threads=True: residual = sum / kept).  Returns (residual, g (W,6), H (6W,6W), kept)."""
    vp, pi, cl, ps = _csr(vox_ptr, pose_idx, clusters, poses)
    W = ps.shape[0]
    H = np.zeros((6 * W, 6 * W)); g = np.zeros(6 * W); res = C.c_double(0)
    kept = load().ref_lidar_hessian(C.c_int(W), C.c_int64(len(vp) - 1), _p(vp), _p(pi), _p(cl), _p(ps), C.c_int(int(threads)),
                                    _p(H), _p(g), C.byref(res))
    return res.value, g.reshape(W, 6), H, int(kept)


def lidar_residual(vox_ptr, pose_idx, clusters, poses):
    vp, pi, cl, ps = _csr(vox_ptr, pose_idx, clusters, poses)
    return float(load().ref_lidar_residual(C.c_int(ps.shape[0]), C.c_int64(len(vp) - 1), _p(vp), _p(pi), _p(cl), _p(ps)))


def lidar_damping_iter(vox_ptr, pose_idx, clusters, poses):
    vp, pi, cl, ps = _csr(vox_ptr, pose_idx, clusters, poses)
    out = ps.copy()
    load().ref_lidar_damping_iter(C.c_int(ps.shape[0]), C.c_int64(len(vp) - 1), _p(vp), _p(pi), _p(cl), _p(out))
    return out


def so3_exp(w):
    w = np.ascontiguousarray(np.atleast_2d(w), np.float64)
    R = np.zeros((len(w), 3, 3))
    load().ref_so3_exp(C.c_int64(len(w)), _p(w), _p(R))
    return R


class Map:
    """cut_voxel -> recut -> tras_opt on a set of scans (runWindowBA / runLidarBA call sequence)."""

    def __init__(self, scans, poses, voxel_size=1.0, eigen_ratio=(0.3, 0.1, 0.06, 0.03)):
        self.lib = load()
        self.W = len(scans)
        ptr = np.zeros(self.W + 1, np.int64)
        ptr[1:] = np.cumsum([len(s) for s in scans])
        xyz = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in scans]), np.float32)
        ps = np.ascontiguousarray(poses, np.float64)
        er = np.asarray(eigen_ratio, np.float32)
        self.voxel_size = float(voxel_size)
        self.h = C.c_void_p(self.lib.ref_map_create(C.c_int(self.W), _p(ptr), _p(xyz), _p(ps), C.c_double(voxel_size), _p(er)))

    def close(self):
        if self.h:
            self.lib.ref_map_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def export(self):
        """Voxels sorted by (key, path) like oracle/voxel_oracle._pack.  Returns (vox_ptr, pose_idx, clusters, meta)."""
        V = int(self.lib.ref_map_num_voxels(self.h)); S = int(self.lib.ref_map_num_slots(self.h))
        key = np.zeros((V, 3), np.int64); path = np.zeros((V, 4), np.int32); layer = np.zeros(V, np.int32)
        vp = np.zeros(V + 1, np.int64); pi = np.zeros(S, np.int32); cl = np.zeros((S, 10))
        centre = np.zeros((V, 3)); direct = np.zeros((V, 3)); ev = np.zeros((V, 3))
        self.lib.ref_map_export(self.h, _p(key), _p(path), _p(layer), _p(vp), _p(pi), _p(cl), _p(centre), _p(direct), _p(ev))
        paths = [tuple(int(x) for x in p if x >= 0) for p in path]
        order = sorted(range(V), key=lambda a: (tuple(key[a]), paths[a]))
        o_vp = [0]; o_pi = []; o_cl = []
        for a in order:
            o_pi.append(pi[vp[a]:vp[a + 1]]); o_cl.append(cl[vp[a]:vp[a + 1]]); o_vp.append(o_vp[-1] + int(vp[a + 1] - vp[a]))
        meta = dict(key=key[order], path=[paths[a] for a in order], layer=layer[order], centre=centre[order], direct=direct[order],
                    eigenvalues=ev[order])
        return (np.asarray(o_vp, np.int64), np.concatenate(o_pi) if o_pi else np.zeros(0, np.int32),
                np.concatenate(o_cl) if o_cl else np.zeros((0, 10)), meta)

    def damping_iter(self):
        out = np.zeros((self.W, 12))
        self.lib.ref_map_damping_iter(self.h, _p(out))
        return out

    def lookup(self, X):
        """Returns (state, direct, centre) of the node findCorrespondPoint ends at; state -1 = no root voxel, 2 = PLANE."""
        X = np.ascontiguousarray(np.asarray(X, np.float64).reshape(-1, 3))
        st = np.zeros(len(X), np.int32); d = np.zeros((len(X), 3)); c = np.zeros((len(X), 3))
        self.lib.ref_map_lookup(self.h, C.c_int64(len(X)), _p(X), C.c_double(self.voxel_size), _p(st), _p(d), _p(c))
        return st, d, c


def anchor_cloud(scans, rel, leaf):
    W = len(scans)
    ptr = np.zeros(W + 1, np.int64)
    ptr[1:] = np.cumsum([len(s) for s in scans])
    xyz = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in scans]), np.float32)
    rl = np.ascontiguousarray(rel, np.float64)
    out = np.zeros((len(xyz), 3), np.float32)
    n = load().ref_anchor_cloud(C.c_int(W), _p(ptr), _p(xyz), _p(rl), C.c_double(leaf), _p(out))
    return out[:int(n)]


def reproj(q, t, X, uv, intr, su=1.0, sv=1.0, jac=True):
    q, t, X, uv = (np.ascontiguousarray(a, np.float64) for a in (q, t, X, uv))
    intr = np.ascontiguousarray(intr, np.float64)
    n = len(q)
    r = np.zeros((n, 2)); J = np.zeros((n, 2, 10)) if jac else None
    load().ref_reproj(C.c_int64(n), _p(q), _p(t), _p(X), _p(uv), _p(intr), C.c_double(su), C.c_double(sv), _p(r),
                      _p(J) if jac else None)
    return r, J


def point_plane(X, nd, sigma, jac=True):
    X = np.ascontiguousarray(X, np.float64); nd = np.ascontiguousarray(nd, np.float64)
    n = len(X)
    r = np.zeros(n); J = np.zeros((n, 3)) if jac else None
    load().ref_point_plane(C.c_int64(n), _p(X), _p(nd), C.c_double(sigma), _p(r), _p(J) if jac else None)
    return r, J


def project_camera_to_pixel(intr, Xc):
    Xc = np.ascontiguousarray(Xc, np.float64); intr = np.ascontiguousarray(intr, np.float64)
    n = len(Xc)
    uv = np.zeros((n, 2)); z = np.zeros(n); ok = np.zeros(n, np.uint8)
    load().ref_project_camera_to_pixel(C.c_int64(n), _p(intr), _p(Xc), _p(uv), _p(z), _p(ok))
    return ok.astype(bool), uv, z


def undistort_pixel(intr, uv):
    uv = np.ascontiguousarray(uv, np.float64); intr = np.ascontiguousarray(intr, np.float64)
    n = len(uv)
    xy = np.zeros((n, 2)); ok = np.zeros(n, np.uint8)
    load().ref_undistort_pixel(C.c_int64(n), _p(intr), _p(uv), _p(xy), _p(ok))
    return ok.astype(bool), xy


def depth_candidate(depth, intr, cam, uv):
    """fetchDepthBilinear -> backProjectPixelDepthDistorted -> camToWorld.  Returns (ok bits, d float32, Xw)."""
    depth = np.ascontiguousarray(depth, np.float32); uv = np.ascontiguousarray(uv, np.float32)
    intr = np.ascontiguousarray(intr, np.float64); cam = np.ascontiguousarray(cam, np.float64)
    n = len(uv)
    d = np.zeros(n, np.float32); Xw = np.zeros((n, 3)); ok = np.zeros(n, np.uint8)
    load().ref_depth_candidate(C.c_int(depth.shape[0]), C.c_int(depth.shape[1]), _p(depth), _p(intr), _p(cam), C.c_int64(n), _p(uv),
                               _p(d), _p(Xw), _p(ok))
    return ok, d, Xw
