"""ctypes binding of oracle/libcpu_ref.so — the threaded C++ restatement of the reference's CPU path.
TEST / BENCH INFRASTRUCTURE (see cpu_ref.cpp header); never imported by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def load():
    global _lib
    if _lib is None:
        so = _HERE / "libcpu_ref.so"
        if not so.exists() or so.stat().st_mtime < (_HERE / "cpu_ref.cpp").stat().st_mtime:
            r = subprocess.run(["make", "-C", str(_HERE)], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stdout + r.stderr)
        _lib = C.CDLL(str(so))
        _lib.ref_lidar_residual.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def hardware_threads():
    return int(load().ref_hardware_threads())


def lidar_lm(vox_ptr, pose_idx, clusters, poses, u0=0.01, v0=2.0, max_iter=10, rel_tol=1e-6, threads=None):
    lib = load()
    vp = np.ascontiguousarray(vox_ptr, np.int64); pi = np.ascontiguousarray(pose_idx, np.int32)
    cl = np.ascontiguousarray(clusters, np.float64); ps = np.array(poses, np.float64, order="C")
    out = np.zeros(16)
    threads = threads or hardware_threads()
    lib.ref_lidar_lm(C.c_int32(ps.shape[0]), C.c_int64(len(vp) - 1), _p(vp, C.c_int64), _p(pi, C.c_int32), _p(cl, C.c_double),
                     _p(ps, C.c_double), C.c_double(u0), C.c_double(v0), C.c_int32(max_iter), C.c_double(rel_tol),
                     C.c_int32(threads), _p(out, C.c_double))
    keys = ["iterations", "accepted", "builds", "cost_first", "cost_last", "u_last", "ms_total", "ms_build", "ms_solve", "ms_residual"]
    info = dict(zip(keys, out)); info["threads"] = threads
    return ps, info


def lidar_step(vox_ptr, pose_idx, clusters, poses, u=0.01, threads=16):
    """dx of one damped step (H + u diag H) dx = -g at `poses`, and sum(lambda_0)."""
    lib = load()
    vp = np.ascontiguousarray(vox_ptr, np.int64); pi = np.ascontiguousarray(pose_idx, np.int32)
    cl = np.ascontiguousarray(clusters, np.float64); ps = np.ascontiguousarray(poses, np.float64)
    W = ps.shape[0]
    dx = np.zeros((W, 6)); r = C.c_double()
    lib.ref_lidar_step(C.c_int32(W), C.c_int64(len(vp) - 1), _p(vp, C.c_int64), _p(pi, C.c_int32), _p(cl, C.c_double), _p(ps, C.c_double),
                       C.c_double(u), C.c_int32(threads), _p(dx, C.c_double), C.byref(r))
    return dx, r.value


def lidar_build(vox_ptr, pose_idx, clusters, poses, threads=4):
    lib = load()
    vp = np.ascontiguousarray(vox_ptr, np.int64); pi = np.ascontiguousarray(pose_idx, np.int32)
    cl = np.ascontiguousarray(clusters, np.float64); ps = np.ascontiguousarray(poses, np.float64)
    W = ps.shape[0]
    nb = C.c_int64()
    lib.ref_lidar_structure(C.c_int32(W), C.c_int64(len(vp) - 1), _p(vp, C.c_int64), _p(pi, C.c_int32), C.byref(nb), None, None)
    br = np.empty(nb.value, np.int32); bc = np.empty(nb.value, np.int32)
    lib.ref_lidar_structure(C.c_int32(W), C.c_int64(len(vp) - 1), _p(vp, C.c_int64), _p(pi, C.c_int32), C.byref(nb), _p(br, C.c_int32), _p(bc, C.c_int32))
    r = C.c_double(); g = np.empty((W, 6)); bl = np.empty((nb.value, 6, 6))
    lib.ref_lidar_build(C.c_int32(W), C.c_int64(len(vp) - 1), _p(vp, C.c_int64), _p(pi, C.c_int32), _p(cl, C.c_double), _p(ps, C.c_double),
                        C.c_int32(threads), C.byref(r), _p(g, C.c_double), _p(bl, C.c_double))
    return r.value, g, br, bc, bl


def lidar_residual(vox_ptr, pose_idx, clusters, poses):
    lib = load()
    vp = np.ascontiguousarray(vox_ptr, np.int64); pi = np.ascontiguousarray(pose_idx, np.int32)
    cl = np.ascontiguousarray(clusters, np.float64); ps = np.ascontiguousarray(poses, np.float64)
    return float(lib.ref_lidar_residual(C.c_int32(ps.shape[0]), C.c_int64(len(vp) - 1), _p(vp, C.c_int64), _p(pi, C.c_int32),
                                        _p(cl, C.c_double), _p(ps, C.c_double)))


def visual_lm(q, t, X, plane_nd, obs_ptr, obs_cam, obs_uv, intr, sigma_px, sigma_plane, fixed_cam=0, max_iter=50,
              threads=None, jacobi_scaling=True, function_tolerance=1e-6):
    lib = load()
    q = np.array(q, np.float64, order="C"); t = np.array(t, np.float64, order="C"); X = np.array(X, np.float64, order="C")
    pl = np.ascontiguousarray(plane_nd, np.float64); op = np.ascontiguousarray(obs_ptr, np.int64)
    oc = np.ascontiguousarray(obs_cam, np.int32); uv = np.ascontiguousarray(obs_uv, np.float32); it = np.ascontiguousarray(intr, np.float64)
    out = np.zeros(16)
    threads = threads or hardware_threads()
    lib.ref_visual_lm(C.c_int32(q.shape[0]), C.c_int64(X.shape[0]), _p(q, C.c_double), _p(t, C.c_double), _p(X, C.c_double),
                      _p(pl, C.c_double), _p(op, C.c_int64), _p(oc, C.c_int32), _p(uv, C.c_float), _p(it, C.c_double),
                      C.c_double(sigma_px), C.c_double(sigma_plane), C.c_int32(fixed_cam), C.c_int32(max_iter), C.c_int32(threads),
                      C.c_int32(int(jacobi_scaling)), C.c_double(function_tolerance), _p(out, C.c_double))
    keys = ["iterations", "accepted", "builds", "cost_first", "cost_last", "radius", "ms_total", "ms_build", "ms_solve", "ms_residual", "termination"]
    info = dict(zip(keys, out)); info["threads"] = threads
    return q, t, X, info
