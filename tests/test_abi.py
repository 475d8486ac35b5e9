"""No-GPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/lvba_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    txt = (ROOT / "include" / "lvba_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lvba_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported(pkg):
    lib = pkg.load_library()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(pkg.EXPORTS) == names


def test_library_is_sm100a_only(pkg):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(pkg.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_version_and_strings(pkg):
    lib = pkg.load_library()
    assert lib.lvba_version() == 100
    assert lib.lvba_status_string(-2).decode().startswith("no CUDA device")
    o = pkg.lidar_default_opts()
    assert (o.u0, o.v0, o.max_iter, o.rel_tol) == (0.01, 2.0, 10, 1e-6)       # bavoxel.hpp:664,686,760
    v = pkg.visual_default_opts()
    assert (v.max_iter, v.initial_radius, v.min_relative_decrease) == (50, 1e4, 1e-3)


def test_no_cpu_fallback(pkg, problem_small):
    """On a box without a GPU every compute entry point must fail loudly, never compute on the CPU."""
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = problem_small
    with pytest.raises(pkg.LvbaError) as e:
        pkg.lidar_lm(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
    assert e.value.status == -2
    with pytest.raises(pkg.LvbaError) as e:
        pkg.visual_lm(p["q"], p["t"], p["X"], p["plane_nd"], p["obs_ptr"], p["obs_cam"], p["obs_uv"], p["intr"], 0.5, 0.01)
    assert e.value.status == -2


def test_argument_validation_happens_before_device_use(pkg, problem_small):
    p = problem_small
    bad = p["pose_idx"].copy(); bad[0] = -1
    with pytest.raises(pkg.LvbaError) as e:
        pkg.lidar_lm(p["vox_ptr"], bad, p["clusters"], p["poses"])
    assert e.value.status == -1
    poses_before = p["poses"].copy()
    assert np.array_equal(p["poses"], poses_before)          # in/out buffer untouched on error


def test_shard_owner_rule(pkg):
    """Contiguous pose-block rows: rows [p*n/P, (p+1)*n/P) -> rank p."""
    for n, P in ((10, 3), (5000, 8), (7, 7), (3, 8)):
        owners = [pkg.shard_owner(i, n, P) for i in range(n)]
        assert owners == sorted(owners) and (n < P or owners[0] == 0) and max(owners) <= P - 1
        for r in range(P):
            rows = [i for i in range(n) if owners[i] == r]
            assert rows == list(range(r * n // P, (r + 1) * n // P))


def test_window_batch_argument_checks_need_no_gpu(pkg):
    """lvba_lidar_lm_batch validates its window table on the host before touching a device."""
    p_ = np.zeros((4, 12)); wp = np.array([0, 4], np.int32)
    with pytest.raises(pkg.LvbaError) as e:
        pkg.lidar_lm_batch(np.array([1, 4], np.int32), np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros((0, 10)), p_)
    assert e.value.status == -1                                              # win_ptr[0] != 0
    with pytest.raises(pkg.LvbaError) as e:
        pkg.lidar_lm_batch(np.array([0, 4, 2], np.int32), np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros((0, 10)), p_)
    assert e.value.status == -1                                              # decreasing
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.lidar_lm_batch(wp, np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros((0, 10)), p_)
        assert e.value.status == -2
    o = pkg.FuseOpts(); pkg.load_library().lvba_fuse_default_opts(__import__("ctypes").byref(o))
    assert o.map_order == pkg.FUSE_ORDER_LIBSTDCXX == 1 and pkg.FUSE_ORDER_ASCENDING == 0      # the default is the order of a g++ build of the reference                                          # no CPU fallback


def test_voxel_map_argument_checks_need_no_gpu(pkg):
    """lvba_voxel_map_create validates scans / options on the host; without a device it then refuses (no CPU path)."""
    o = pkg.voxel_default_opts()
    assert (o.voxel_size, o.layer_limit, o.min_points) == (1.0, 2, 15)                    # bavoxel.hpp:13,24
    assert [round(float(x), 6) for x in o.eigen_ratio] == [0.3, 0.1, 0.06, 0.03]          # bavoxel.hpp:17
    scans = [np.zeros((4, 3), np.float32), np.ones((5, 3), np.float32)]
    poses = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (2, 1))
    for kw in (dict(voxel_size=0.0), dict(voxel_size=float("nan")), dict(min_points=-1), dict(eigen_ratio=(0.3, -1, 0.06, 0.03))):
        with pytest.raises(pkg.LvbaError) as e:
            pkg.VoxelMap(scans, poses, **kw)
        assert e.value.status == -1
    with pytest.raises(pkg.LvbaError) as e:
        pkg.VoxelMap(scans, poses, layer_limit=3)
    assert e.value.status == -4
    with pytest.raises(pkg.LvbaError) as e:                                               # non-monotone scan table
        pkg.VoxelMap(np.zeros((9, 3), np.float32), poses, scan_ptr=[0, 6, 4])
    assert e.value.status == -1
    bad = poses.copy(); bad[1, 4] = np.inf
    with pytest.raises(pkg.LvbaError) as e:
        pkg.VoxelMap(scans, bad)
    assert e.value.status == -1
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.VoxelMap(scans, poses)
        assert e.value.status == -2


def test_depth_grid_argument_checks_need_no_gpu(pkg):
    """lvba_depth_grid_create validates on the host; without a device it then refuses (no CPU path)."""
    scans = [np.zeros((4, 3), np.float32), np.ones((5, 3), np.float32)]
    poses = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (2, 1))
    for kw, ts in ((dict(voxel_size=0.0), [0.0, 0.1]), (dict(), [0.2, 0.1]), (dict(), [0.0, float("nan")])):
        with pytest.raises(pkg.LvbaError) as e:
            pkg.DepthGrid(scans, poses, ts, **kw)
        assert e.value.status == -1
    with pytest.raises(pkg.LvbaError) as e:
        pkg.DepthGrid(np.zeros((9, 3), np.float32), poses, [0.0, 0.1], scan_ptr=[0, 6, 4])
    assert e.value.status == -1
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.DepthGrid(scans, poses, [0.0, 0.1])
        assert e.value.status == -2


def test_windowed_voxel_map_argument_checks_need_no_gpu(pkg):
    scans = [np.zeros((4, 3), np.float32), np.ones((5, 3), np.float32), np.ones((3, 3), np.float32)]
    poses = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (3, 1))
    for wp in ([1, 3], [0, 2, 1, 3]):                                   # win_ptr[0] != 0 ; decreasing
        with pytest.raises(pkg.LvbaError) as e:
            pkg.VoxelMap(scans, poses, win_ptr=wp)
        assert e.value.status == -1
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.VoxelMap(scans, poses, win_ptr=[0, 2, 3])
        assert e.value.status == -2


def test_track_numerics_argument_checks_need_no_gpu(pkg):
    cams = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (2, 1)); intr = np.array([100.0, 100, 50, 50, 0, 0, 0, 0])
    with pytest.raises(pkg.LvbaError) as e:
        pkg.tracks_triangulate([0, 3, 2], np.zeros(3, np.int32), np.zeros((3, 2), np.float32), cams, intr)      # non-monotone CSR
    assert e.value.status == -1
    bad = intr.copy(); bad[4] = np.nan
    with pytest.raises(pkg.LvbaError) as e:
        pkg.tracks_mean_reproj([0, 2], np.zeros(2, np.int32), np.zeros((2, 2), np.float32), cams, bad, np.zeros((1, 3)), 2)
    assert e.value.status == -1
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.tracks_triangulate([0, 2], np.zeros(2, np.int32), np.zeros((2, 2), np.float32), cams, intr)
        assert e.value.status == -2


def test_anchor_clouds_argument_checks_need_no_gpu(pkg):
    scans = [np.zeros((4, 3), np.float32), np.ones((5, 3), np.float32)]
    rel = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (2, 1))
    for wp, leaf in (([1, 2], 0.1), ([0, 2, 1], 0.1), ([0, 2], -1.0), ([0, 2], float("nan"))):
        with pytest.raises(pkg.LvbaError) as e:
            pkg.anchor_clouds(scans, rel, wp, leaf)
        assert e.value.status == -1
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.anchor_clouds(scans, rel, [0, 2], 0.1)
        assert e.value.status == -2


def test_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: include/lvba_b200.h must compile as C99 without extensions."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "lvba_b200.h"\nint main(void) { lvba_voxel_opts o; lvba_voxel_default_opts(&o); return lvba_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "hdr.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_null_handles_and_pointers_are_refused_not_dereferenced(pkg):
    """Every handle-taking entry point of the set-up boundaries with NULL where a pointer is required: a status, never a crash."""
    import ctypes as C
    lib = pkg.load_library()
    null = C.c_void_p()
    i64 = C.c_int64(); vs = pkg.VoxelSummary(); ds = pkg.DepthSummary(); s = pkg.Summary()
    poses = (C.c_double * 12)(*([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]))
    intr = (C.c_double * 8)(100, 100, 50, 50, 0, 0, 0, 0)
    sp = (C.c_int64 * 2)(0, 0)
    wp = (C.c_int32 * 2)(0, 1)
    out = C.c_void_p()
    calls = [
        lambda: lib.lvba_voxel_map_create(C.c_int32(1), None, None, C.c_int32(3), poses, None, C.byref(out), None),
        lambda: lib.lvba_voxel_map_create(C.c_int32(1), sp, None, C.c_int32(3), poses, None, None, None),
        lambda: lib.lvba_voxel_map_create(C.c_int32(-1), sp, None, C.c_int32(3), poses, None, C.byref(out), None),
        lambda: lib.lvba_voxel_map_create(C.c_int32(1), sp, None, C.c_int32(2), poses, None, C.byref(out), None),
        lambda: lib.lvba_voxel_map_create_windows(C.c_int32(0), wp, sp, None, C.c_int32(3), poses, None, C.byref(out), None),
        lambda: lib.lvba_voxel_map_create_windows(C.c_int32(1), None, sp, None, C.c_int32(3), poses, None, C.byref(out), None),
        lambda: lib.lvba_voxel_map_summary(null, C.byref(vs)),
        lambda: lib.lvba_voxel_map_export(null, None, None, None, None, None, None, None, None),
        lambda: lib.lvba_voxel_map_lookup(null, C.c_int64(1), poses, poses),
        lambda: lib.lvba_voxel_map_windows(null, None, None),
        lambda: lib.lvba_voxel_map_lidar_create(null, poses, C.byref(out)),
        lambda: lib.lvba_voxel_map_lidar_lm(null, poses, C.c_int32(0), None, C.byref(s)),
        lambda: lib.lvba_voxel_map_lidar_lm_batch(null, poses, C.c_int32(3), None, None, None),
        lambda: lib.lvba_depth_grid_create(C.c_int32(1), None, None, C.c_int32(3), poses, poses, C.c_double(0.5), C.c_int32(-1), C.byref(out), None),
        lambda: lib.lvba_depth_grid_create(C.c_int32(1), sp, None, C.c_int32(3), poses, poses, C.c_double(0.5), C.c_int32(-1), None, None),
        lambda: lib.lvba_depth_render(null, C.c_int32(1), poses, poses, C.c_double(0.5), intr, C.c_int32(4), C.c_int32(4), None, C.byref(ds)),
        lambda: lib.lvba_depth_backproject(null, C.c_int32(1), poses, poses, C.c_double(0.5), intr, C.c_int32(4), C.c_int32(4), sp, None, None, None, None),
        lambda: lib.lvba_anchor_clouds_create(C.c_int32(1), wp, None, None, C.c_int32(3), poses, C.c_double(0.1), C.c_int32(-1), C.byref(out), C.byref(i64)),
        lambda: lib.lvba_anchor_clouds_create(C.c_int32(1), wp, sp, None, C.c_int32(3), poses, C.c_double(0.1), C.c_int32(-1), None, C.byref(i64)),
        lambda: lib.lvba_anchor_clouds_export(null, None, None, None),
        lambda: lib.lvba_tracks_triangulate(C.c_int64(1), None, None, None, C.c_int32(1), poses, intr, C.c_int32(-1), None, None, None, None),
        lambda: lib.lvba_tracks_mean_reproj(C.c_int64(-1), sp, None, None, C.c_int32(1), poses, intr, C.c_int32(-1), None, C.c_int32(1), None, None, None),
    ]
    for k, call in enumerate(calls):
        assert call() == -1, k                                       # LVBA_ERR_INVALID_ARG
    for destroy in (lib.lvba_voxel_map_destroy, lib.lvba_depth_grid_destroy, lib.lvba_anchor_clouds_destroy):
        assert destroy(null) == 0                                    # destroying nothing is fine


def test_track_fusion_argument_checks_need_no_gpu(pkg):
    """lvba_tracks_fuse_create (boundary B7): every malformed input is refused before a device is looked for."""
    cams = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (2, 1)); intr = np.array([100.0, 100, 50, 50, 0, 0, 0, 0])
    kp_ptr = np.array([0, 2, 4], np.int64); uv = np.zeros((4, 2), np.float32); m = np.array([[0, 0, 1, 0], [0, 1, 1, 1]], np.int32)
    X = np.zeros((4, 3)); valid = np.ones(4, np.uint8)
    bad_cams = cams.copy(); bad_cams[1, 10] = np.inf
    bad_intr = intr.copy(); bad_intr[0] = np.nan
    bad_X = X.copy(); bad_X[2, 1] = np.nan
    cases = [dict(kp_ptr=np.array([1, 2, 4], np.int64)), dict(kp_ptr=np.array([0, 3, 2], np.int64)), dict(cams=bad_cams), dict(intr=bad_intr),
             dict(kp_Xw=bad_X), dict(obser_thr=0), dict(reproj_thr=-1.0), dict(depth_gate=float("nan")), dict(min_view_angle_deg=float("inf")),
             dict(map_order=7), dict(map_order=-1)]
    for kw in cases:
        a = dict(kp_ptr=kp_ptr, kp_uv=uv, matches=m, cams=cams, intr=intr, kp_Xw=X, kp_valid=valid)
        opt = {k: kw.pop(k) for k in list(kw) if k in ("obser_thr", "reproj_thr", "depth_gate", "min_view_angle_deg", "map_order")}
        a.update(kw)
        with pytest.raises(pkg.LvbaError) as e:
            pkg.tracks_fuse(a["kp_ptr"], a["kp_uv"], a["matches"], a["cams"], a["intr"], a["kp_Xw"], a["kp_valid"], **opt)
        assert e.value.status == -1, (kw, opt)
    # a NaN depth candidate that is flagged invalid is not an error; without a device the call then stops at the device check
    ok_X = bad_X.copy(); v2 = valid.copy(); v2[2] = 0
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.tracks_fuse(kp_ptr, uv, m, cams, intr, ok_X, v2)
        assert e.value.status == -2


def test_solver_entry_argument_checks_need_no_gpu(pkg):
    """lvba_env_solve: the envelope description is validated on the host (first[r] in [0, r], non-decreasing, known path)."""
    n = 4
    blocks = np.zeros((10, 36)); dadd = np.ones(6 * n); rhs = np.ones(6 * n)
    for first, path in (([0, 0, 3, 2], pkg.SOLVE_AUTO), ([0, 2, 1, 1], pkg.SOLVE_AUTO), ([0, -1, 0, 0], pkg.SOLVE_AUTO), ([0, 0, 0, 0], 99), ([0, 0, 0, 0], -3)):
        with pytest.raises(pkg.LvbaError) as e:
            pkg.env_solve(first, blocks, dadd, rhs, path=path)
        assert e.value.status == -1, (first, path)
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LvbaError) as e:
            pkg.env_solve([0, 0, 0, 0], blocks, dadd, rhs)
        assert e.value.status == -2


def test_communicator_argument_checks_need_no_gpu(pkg):
    lib = pkg.load_library()
    import ctypes as C
    assert lib.lvba_comm_init(C.c_int32(0), C.c_int32(0), None, C.c_int32(0)) == -1          # n_ranks < 1
    assert lib.lvba_comm_init(C.c_int32(2), C.c_int32(2), None, C.c_int32(0)) == -1          # rank >= n_ranks
    assert lib.lvba_comm_unique_id(None) == -1
    n, r = C.c_int32(-5), C.c_int32(-5)
    assert lib.lvba_comm_info(C.byref(n), C.byref(r)) == 0 and (n.value, r.value) == (1, 0)   # no communicator: one rank, rank 0
    assert lib.lvba_comm_destroy() == 0                                                       # idempotent
