"""ctypes binding of oracle/_ref/liblvba_system_ref.so — the reference's pipeline source src/lvba_system.cpp compiled where it lies
(oracle/ref_system_driver.cpp, `make -C oracle ref`) and driven through LvbaSystem's own public members.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Exists only where /root/reference exists; `available()` says so and the tests that need it
skip otherwise.  The committed fixture it wrote (tests/golden/ref_system.npz, tests/golden/make_golden_ref_system.py) travels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "liblvba_system_ref.so")
_lib = None
_CB = C.CFUNCTYPE(None, C.c_void_p)


def available() -> bool:
    return os.path.exists(_PATH)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.sys_create.restype = C.c_void_p
        _lib.sys_build_tracks.restype = C.c_int64
        _lib.sys_optimize_camera_poses.restype = C.c_int64
        _lib.sys_run_window_ba.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _scans(scans):
    ptr = np.zeros(len(scans) + 1, np.int64)
    ptr[1:] = np.cumsum([len(s) for s in scans])
    xyz = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in scans]), np.float32)
    return ptr, xyz


def track_helpers(obs_ptr, obs_cam, obs_uv, cams, intr, Xw=None, min_count=3):
    """The reference's file-scope TriangulateTrackDLT (Xw None) / ComputeMeanReproj (Xw given) on CSR tracks.  Returns (Xw, mean, count, ok)."""
    op = np.ascontiguousarray(obs_ptr, np.int64); oc = np.ascontiguousarray(obs_cam, np.int32); uv = np.ascontiguousarray(obs_uv, np.float32)
    cm = np.ascontiguousarray(cams, np.float64).reshape(-1, 12); it = np.ascontiguousarray(intr, np.float64)
    T = len(op) - 1
    out = np.zeros((T, 3)); mean = np.zeros(T); cnt = np.zeros(T, np.int32); ok = np.zeros(T, np.uint8)
    xin = None if Xw is None else np.ascontiguousarray(Xw, np.float64)
    load().sys_track_helpers(C.c_int64(T), _p(op), _p(oc), _p(uv), C.c_int(len(cm)), _p(cm), _p(it), _p(xin), C.c_int(min_count), _p(out), _p(mean), _p(cnt), _p(ok))
    return (out if Xw is None else xin), mean, cnt, ok.astype(bool)


def bucket_count_after_reserve(n):
    lib = load(); lib.sys_bucket_count_after_reserve.restype = C.c_uint64
    return int(lib.sys_bucket_count_after_reserve(C.c_uint64(int(n))))


def unordered_map_order(reserve, keys):
    """Iteration order of std::unordered_map<int,int> after reserve(reserve) and inserting `keys` (distinct) in this order."""
    k = np.ascontiguousarray(keys, np.int32); out = np.zeros(len(k), np.int32)
    load().sys_unordered_map_order(C.c_int64(int(reserve)), C.c_int64(len(k)), _p(k), _p(out))
    return [int(x) for x in out]


def set_eigen_ratio_array(r):
    a = np.asarray(r, np.float32)
    load().sys_set_eigen_ratio_array(_p(a))


class System:
    """One lvba::LvbaSystem.  Parameters of the ROS parameter server the constructor reads can be given as `params`."""

    def __init__(self, params=None):
        """params: {name: number | str | list of numbers} — the ROS parameter server the constructors read (src/dataset_io.cpp:28-58,
        src/lvba_system.cpp:127-133).  With data_config/data_path naming a dataset directory the reference's own loader fills the system."""
        self.lib = load()
        self.lib.sys_clear_params()
        for k, v in (params or {}).items():
            if isinstance(v, str):
                self.lib.sys_set_param_str(k.encode(), v.encode())
            elif isinstance(v, (list, tuple, np.ndarray)):
                a = np.ascontiguousarray(v, np.float64).ravel()
                self.lib.sys_set_param_vec(k.encode(), C.c_int(len(a)), _p(a))
            else:
                self.lib.sys_set_param(k.encode(), C.c_double(float(v)))
        self.h = C.c_void_p(self.lib.sys_create())
        self.W = 0; self.M = 0; self.width = 0; self.height = 0; self.n_points = 0

    def close(self):
        if self.h:
            self.lib.sys_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def dataset(self):
        """What DatasetIO's constructor loaded: dict(frame_ts, frame_poses, scans, intensity, image_ts, image_poses, cam, data_path, db_path)."""
        nf = C.c_int64(); npt = C.c_int64(); ni = C.c_int64()
        self.lib.sys_dataset_sizes(self.h, C.byref(nf), C.byref(npt), C.byref(ni))
        nf, npt, ni = nf.value, npt.value, ni.value
        fts = np.zeros(nf); fp = np.zeros((nf, 12)); sp = np.zeros(nf + 1, np.int64); xyz = np.zeros((npt, 3), np.float32); inten = np.zeros(npt, np.float32)
        its = np.zeros(ni); ip = np.zeros((ni, 12)); cam = np.zeros(12); paths = C.create_string_buffer(4096)
        self.lib.sys_dataset_get(self.h, _p(fts), _p(fp), _p(sp), _p(xyz), _p(inten), _p(its), _p(ip), _p(cam), paths)
        dp, db = paths.value.decode().split("\n")
        return dict(frame_ts=fts, frame_poses=fp, scans=[xyz[sp[i]:sp[i + 1]] for i in range(len(sp) - 1)], intensity=inten, image_ts=its, image_poses=ip,
                    cam=cam, data_path=dp, db_path=db)

    def init_from_dataset(self):
        """initFromDatasetIO (:448-507) on what the reference's own loader read (the constructor ran it on data_config/data_path)."""
        d = self.dataset()
        self.W = len(d["frame_ts"]); self.M = len(d["image_ts"]); self.width = int(d["cam"][0]); self.height = int(d["cam"][1])
        self.n_points = sum(len(x) for x in d["scans"])
        self.lib.sys_init_from_dataset(self.h)
        return d

    # ---- inputs
    def set_lidar(self, scans, poses, ts=None):
        ptr, xyz = _scans(scans)
        ps = np.ascontiguousarray(poses, np.float64)
        t = None if ts is None else np.ascontiguousarray(ts, np.float64)
        self.W = len(scans); self.n_points = len(xyz)
        self.lib.sys_set_lidar(self.h, C.c_int(self.W), _p(ptr), _p(xyz), _p(ps), _p(t))

    def set_lidar_optimised(self, poses):
        self.lib.sys_set_lidar_optimised(self.h, _p(np.ascontiguousarray(poses, np.float64)))

    def set_stages(self, window_enable=True, window_size=10, anchor_leaf=0.1, use_rel=False, stage1_enable=True, s1_voxel=0.5,
                   s1_ratio=(0.3, 0.1, 0.06, 0.03), s2_voxel=0.5, s2_ratio=(0.08, 0.08, 0.08, 0.08)):
        a = np.asarray(s1_ratio, np.float32); b = np.asarray(s2_ratio, np.float32)
        self.lib.sys_set_stages(self.h, C.c_int(int(window_enable)), C.c_int(window_size), C.c_double(anchor_leaf), C.c_int(int(use_rel)),
                                C.c_int(int(stage1_enable)), C.c_double(s1_voxel), _p(a), C.c_double(s2_voxel), _p(b))

    def set_camera(self, width, height, intr, Rcl, tcl, Ril, til, image_ts, image_poses):
        f = lambda a: np.ascontiguousarray(a, np.float64)  # noqa: E731
        self.M = len(image_ts); self.width = width; self.height = height
        self.lib.sys_set_camera(self.h, C.c_int(width), C.c_int(height), _p(f(intr)), _p(f(Rcl)), _p(f(tcl)), _p(f(Ril)), _p(f(til)),
                                C.c_int(self.M), _p(f(image_ts)), _p(f(image_poses)))
        self.lib.sys_init_from_dataset(self.h)

    # ---- LiDAR half
    def run_window_ba(self):
        A = (self.W + 0) or 1
        ap = np.zeros((A, 12)); cp = np.zeros(A + 1, np.int64); cl = np.zeros((max(self.n_points, 1), 3), np.float32)
        rel = np.zeros((self.W, 12)); idx = np.zeros(self.W, np.int32)
        n = self.lib.sys_run_window_ba(self.h, _p(ap), _p(cp), _p(cl), _p(rel), _p(idx))
        return ap[:n], [cl[cp[a]:cp[a + 1]].copy() for a in range(n)], rel, idx

    def run_lidar_ba(self):
        out = np.zeros((self.W, 12))
        self.lib.sys_run_lidar_ba(self.h, _p(out))
        return out

    # ---- camera half
    def build_grid(self):
        self.lib.sys_build_grid(self.h)

    def update_camera_poses(self):
        out = np.zeros((self.M, 12))
        self.lib.sys_update_camera_poses(self.h, _p(out))
        return out

    def generate_depth(self):
        d = np.zeros((self.M, self.height, self.width), np.float32); c0 = np.zeros((self.M, 12)); c1 = np.zeros((self.M, 12))
        self.lib.sys_generate_depth(self.h, _p(d), _p(c0), _p(c1))
        return d, c0, c1

    def set_fusion_inputs(self, cams, depth, intr, kp_ptr, kp_uv, matches, obser_thr=3):
        cams = np.ascontiguousarray(cams, np.float64); depth = np.ascontiguousarray(depth, np.float32)
        self.M, self.height, self.width = depth.shape
        kp_ptr = np.ascontiguousarray(kp_ptr, np.int64); kp_uv = np.ascontiguousarray(kp_uv, np.float32)
        m = np.ascontiguousarray(matches, np.int32).reshape(-1, 4)
        self.lib.sys_set_fusion_inputs(self.h, C.c_int(self.M), C.c_int(self.width), C.c_int(self.height), _p(np.ascontiguousarray(intr, np.float64)),
                                       _p(cams), _p(depth), _p(kp_ptr), _p(kp_uv), C.c_int64(len(m)), _p(m), C.c_int(obser_thr))

    def set_keypoints_and_matches(self, kp_ptr, kp_uv, matches):
        kp_ptr = np.ascontiguousarray(kp_ptr, np.int64); kp_uv = np.ascontiguousarray(kp_uv, np.float32)
        m = np.ascontiguousarray(matches, np.int32).reshape(-1, 4)
        self.lib.sys_set_keypoints_and_matches(self.h, _p(kp_ptr), _p(kp_uv), C.c_int64(len(m)), _p(m))

    def load_colmap_db(self, dataset_path, db_path):
        """loadFromColmapDB; returns (ok, kp_ptr, kp_uv, matches (n, 4))."""
        ok = bool(self.lib.sys_load_colmap_db(self.h, str(dataset_path).encode(), str(db_path).encode()))
        nk = C.c_int64(); nm = C.c_int64()
        self.lib.sys_frontend_sizes(self.h, C.byref(nk), C.byref(nm))
        kp_ptr = np.zeros(self.M + 1, np.int64); kp_uv = np.zeros((nk.value, 2), np.float32); m = np.zeros((nm.value, 4), np.int32)
        self.lib.sys_get_frontend(self.h, _p(kp_ptr), _p(kp_uv), _p(m))
        return ok, kp_ptr, kp_uv, m

    def colmap_export(self, dataset_path, grey=128):
        """VisualizeOptComparison: writes <dataset_path>/Colmap/sparse/images.txt and points3D.txt (every image = one grey value)."""
        self.lib.sys_colmap_export(self.h, str(dataset_path).encode(), C.c_int(self.width), C.c_int(self.height), C.c_int(grey))

    def build_tracks(self):
        no = C.c_int64(); ni = C.c_int64()
        n = int(self.lib.sys_build_tracks(self.h, C.byref(no), C.byref(ni)))
        op = np.zeros(n + 1, np.int64); ob = np.zeros((no.value, 2), np.int32); ip = np.zeros(n + 1, np.int64); il = np.zeros(ni.value, np.int32)
        Xw = np.zeros((n, 3))
        self.lib.sys_get_tracks(self.h, _p(op), _p(ob), _p(ip), _p(il), _p(Xw))
        return dict(obs_ptr=op, obs=ob, inl_ptr=ip, inl=il, Xw=Xw)

    def optimize_camera_poses(self, solver=None):
        """Runs optimizeCameraPoses; `solver(problem dict) -> (q, t, X) or None` stands where ceres::Solve stands.
        Returns (problem dict as recorded, cameras (M, 12) after the reference's write-back)."""
        rec = {}

        def cb(_):
            nc = C.c_int64(); npt = C.c_int64(); no = C.c_int64(); npl = C.c_int64()
            self.lib.sys_problem_sizes(self.h, C.byref(nc), C.byref(npt), C.byref(no), C.byref(npl))
            nc, npt, no, npl = nc.value, npt.value, no.value, npl.value
            P = dict(q=np.zeros((nc, 4)), t=np.zeros((nc, 3)), cam_const=np.zeros(nc, np.uint8), cam_manifold=np.zeros(nc, np.int32), X=np.zeros((npt, 3)),
                     obs_cam=np.zeros(no, np.int32), obs_pt=np.zeros(no, np.int32), obs_uv=np.zeros((no, 2)), obs_intr=np.zeros((no, 8)),
                     obs_sigma=np.zeros((no, 2)), obs_loss=np.zeros(no, np.int32), pl_pt=np.zeros(npl, np.int32), pl_nd=np.zeros((npl, 4)),
                     pl_sigma=np.zeros(npl), pl_loss=np.zeros(npl, np.int32), options=np.zeros(8))
            self.lib.sys_problem_get(self.h, *[_p(P[k]) for k in ("q", "t", "cam_const", "cam_manifold", "X", "obs_cam", "obs_pt", "obs_uv", "obs_intr",
                                                                  "obs_sigma", "obs_loss", "pl_pt", "pl_nd", "pl_sigma", "pl_loss", "options")])
            rec.update(P)
            if solver is not None:
                sol = solver(P)
                if sol is not None:
                    q, t, X = (np.ascontiguousarray(a, np.float64) for a in sol)
                    self.lib.sys_problem_set(self.h, _p(q), _p(t), _p(X))

        keep = _CB(cb)
        self.lib.sys_optimize_camera_poses(self.h, keep, None)
        cams = np.zeros((self.M, 12))
        self.lib.sys_get_cameras_optimized(self.h, _p(cams))
        return rec, cams
