"""The depth-rendering pipeline (global-lvba_b200/csrc/depth_pipeline.h — boundary B4) checked WITHOUT
a GPU: the pass functors the CUDA kernels run, instantiated with the sequential host policy (tests/emu/, test
infrastructure only), against oracle/depth_oracle.py.  The z-buffer is order independent, so images are compared
EXACTLY.  tests/test_zz_depth_gpu.py runs the same comparisons through the C ABI."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import depth_oracle as dep
from oracle import synth

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "libdepth_emu.so"
    cmd = ["g++", "-std=c++17", "-O2", *(["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []), "-ffp-contract=off", "-Wall", "-fPIC", "-shared",
           str(ROOT / "tests" / "emu" / "depth_emu.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(str(so))


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def emu_render(lib, s, voxel_size=0.5, half_window=0.5):
    scans = s["scans"]; F = len(scans)
    sp = np.zeros(F + 1, np.int64); sp[1:] = np.cumsum([len(x) for x in scans])
    xyz = np.ascontiguousarray(np.concatenate([np.asarray(x, np.float32).reshape(-1, 3) for x in scans]) if F else np.zeros((0, 3), np.float32))
    poses = np.ascontiguousarray(s["poses"], np.float64); ts = np.ascontiguousarray(s["frame_ts"], np.float64)
    h = ctypes.c_void_p(); nv = ctypes.c_int64(); npairs = ctypes.c_int64()
    rc = lib.emu_depth_grid_create(ctypes.c_int32(F), _ptr(sp, ctypes.c_int64), _ptr(xyz, ctypes.c_float), _ptr(poses, ctypes.c_double),
                                   _ptr(ts, ctypes.c_double), ctypes.c_double(voxel_size), ctypes.byref(h), ctypes.byref(nv), ctypes.byref(npairs))
    if rc != 0:
        return rc, None, None
    cams = np.ascontiguousarray(s["cams"], np.float64); its = np.ascontiguousarray(s["image_ts"], np.float64)
    intr = np.ascontiguousarray(s["intr"], np.float64)
    M, W, H = len(cams), s["width"], s["height"]
    out = np.full((M, H, W), -7.0, np.float32)
    work = np.zeros(2, np.int64)
    rc = lib.emu_depth_render(h, ctypes.c_int64(M), _ptr(cams, ctypes.c_double), _ptr(its, ctypes.c_double), ctypes.c_double(half_window),
                              _ptr(intr, ctypes.c_double), ctypes.c_int32(W), ctypes.c_int32(H), _ptr(out, ctypes.c_float), _ptr(work, ctypes.c_int64))
    lib.emu_depth_grid_destroy(h)
    return rc, out, dict(n_voxels=nv.value, n_pairs=npairs.value, pairs=int(work[0]), chunks=int(work[1]))


def oracle_render(s, **kw):
    return dep.render(s["scans"], s["poses"], s["frame_ts"], s["cams"], s["image_ts"], s["intr"], s["width"], s["height"], **kw)


@pytest.mark.parametrize("seed,voxel_size,half_window", [(1, 0.5, 0.5), (2, 0.5, 0.3), (3, 1.0, 0.5), (4, 0.25, 0.12)])
def test_images_equal_oracle_exactly(emu, seed, voxel_size, half_window):
    s = synth.make_depth_scene(seed, F=7, n_per_scan=2500, M=5)
    rc, got, info = emu_render(emu, s, voxel_size, half_window)
    assert rc == 0
    ref = oracle_render(s, voxel_size=voxel_size, half_window=half_window)
    assert np.array_equal(got, ref)
    assert np.count_nonzero(got) > 500 and info["n_pairs"] >= info["n_voxels"] > 50


def test_literal_oracle_and_window_dedup(emu):
    """Against the literal restatement, and: a voxel seen by several frames of a window is rendered once."""
    s = synth.make_depth_scene(5, F=6, n_per_scan=1200, M=3)
    rc, got, info = emu_render(emu, s)
    lit = dep.render_literal(s["scans"], s["poses"], s["frame_ts"], s["cams"], s["image_ts"], s["intr"], s["width"], s["height"])
    assert rc == 0 and np.array_equal(got, lit)
    # the work list holds every (frame, voxel) pair of the windows, the chunk list only the first pair of each voxel
    pts = sum(len(x) for x in s["scans"])
    assert info["chunks"] * 64 < 64 * info["pairs"] + 1 and info["chunks"] <= info["pairs"] + pts // 64 + 1


def test_edge_cases(emu):
    s = synth.make_depth_scene(6, F=5, n_per_scan=800, M=4)
    s2 = dict(s); s2["image_ts"] = s["image_ts"].copy(); s2["image_ts"][2] = np.nan           # unparsable image name -> empty image
    rc, got, _ = emu_render(emu, s2)
    assert rc == 0 and not got[2].any() and np.array_equal(got, oracle_render(s2))
    far = dict(s); far["image_ts"] = s["image_ts"] + 1e4                                      # no frame in any window
    rc, got, _ = emu_render(emu, far)
    assert rc == 0 and not got.any()
    ragged = dict(s); ragged["scans"] = [s["scans"][0], np.zeros((0, 3), np.float32), s["scans"][2], s["scans"][3][:1], np.zeros((0, 3), np.float32)]
    rc, got, _ = emu_render(emu, ragged)
    assert rc == 0 and np.array_equal(got, oracle_render(ragged))
    empty = dict(s); empty["scans"] = [np.zeros((0, 3), np.float32)] * 5
    rc, got, _ = emu_render(emu, empty)
    assert rc == 0 and not got.any()
    behind = dict(s); behind["cams"] = s["cams"].copy(); behind["cams"][:, 6:9] *= -1; behind["cams"][:, 11] *= -1   # cameras looking away
    rc, got, _ = emu_render(emu, behind)
    assert rc == 0 and np.array_equal(got, oracle_render(behind))
    bad = dict(s); bad["scans"] = [x.copy() for x in s["scans"]]; bad["scans"][1][3, 0] = np.inf
    assert emu_render(emu, bad)[0] == -1
    shifted = dict(s); shifted["poses"] = s["poses"].copy(); shifted["poses"][:, 9:] += [-91.3, 40.2, -7.7]       # negative grid keys
    shifted["cams"] = s["cams"].copy()
    for k in range(len(shifted["cams"])):
        R = shifted["cams"][k, :9].reshape(3, 3); shifted["cams"][k, 9:] -= R @ np.array([-91.3, 40.2, -7.7])
    rc, got, _ = emu_render(emu, shifted)
    assert rc == 0 and np.array_equal(got, oracle_render(shifted)) and got.any()


def test_backprojected_keypoints_equal_oracle_exactly(emu):
    """The depth-fused 3-D candidates of the track fusion (lvba_system.cpp:1020-1038): bilinear fetch in float, fixed-point
    undistortion, camera -> world.  Compared bit for bit; the back-projected points must reproject onto their pixels."""
    s = synth.make_depth_scene(7, F=8, n_per_scan=6000, M=4)
    img = oracle_render(s)
    rng = np.random.default_rng(2)
    W, H = s["width"], s["height"]
    counts = [300, 0, 250, 200]
    uv = np.concatenate([np.column_stack([rng.uniform(-2, W + 1, c), rng.uniform(-2, H + 1, c)]) for c in counts]).astype(np.float32)
    uv[5] = [W - 1, 3.0]; uv[6] = [3.0, H - 1]; uv[7] = [0.0, 0.0]; uv[8] = [np.nan, 4.0]           # borders of the valid range
    kp_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    Xw_ref, valid_ref = dep.backproject(img, s["cams"], s["intr"], kp_ptr, uv)
    assert 20 < valid_ref.sum() < len(uv)
    # through the pipeline: render on the (emulated) device, then sample there
    scans = s["scans"]; F = len(scans)
    sp = np.zeros(F + 1, np.int64); sp[1:] = np.cumsum([len(x) for x in scans])
    xyz = np.ascontiguousarray(np.concatenate(scans), np.float32)
    poses = np.ascontiguousarray(s["poses"]); ts = np.ascontiguousarray(s["frame_ts"])
    h = ctypes.c_void_p(); nv = ctypes.c_int64(); npairs = ctypes.c_int64()
    assert emu.emu_depth_grid_create(ctypes.c_int32(F), _ptr(sp, ctypes.c_int64), _ptr(xyz, ctypes.c_float), _ptr(poses, ctypes.c_double),
                                     _ptr(ts, ctypes.c_double), ctypes.c_double(0.5), ctypes.byref(h), ctypes.byref(nv), ctypes.byref(npairs)) == 0
    cams = np.ascontiguousarray(s["cams"]); its = np.ascontiguousarray(s["image_ts"]); intr = np.ascontiguousarray(s["intr"])
    depth = np.zeros((4, H, W), np.float32)
    assert emu.emu_depth_render(h, ctypes.c_int64(4), _ptr(cams, ctypes.c_double), _ptr(its, ctypes.c_double), ctypes.c_double(0.5),
                                _ptr(intr, ctypes.c_double), ctypes.c_int32(W), ctypes.c_int32(H), _ptr(depth, ctypes.c_float), None) == 0
    Xw = np.full((len(uv), 3), 9.0); valid = np.full(len(uv), 9, np.uint8)
    assert emu.emu_depth_backproject(h, ctypes.c_int64(4), _ptr(depth, ctypes.c_float), _ptr(cams, ctypes.c_double), _ptr(intr, ctypes.c_double),
                                     ctypes.c_int32(W), ctypes.c_int32(H), _ptr(kp_ptr, ctypes.c_int64), _ptr(uv, ctypes.c_float),
                                     _ptr(Xw, ctypes.c_double), valid.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))) == 0
    emu.emu_depth_grid_destroy(h)
    assert np.array_equal(valid, valid_ref) and np.array_equal(Xw, Xw_ref)
    # size-independent property: a valid candidate projects back onto its keypoint (distortion inverse converged)
    for q in np.nonzero(valid)[0][:50]:
        k = int(np.searchsorted(kp_ptr, q, side="right") - 1)
        R = s["cams"][k][:9].reshape(3, 3); t = s["cams"][k][9:]
        ok, uu, vv = dep.project((R @ Xw[q] + t)[None, :], s["intr"])
        assert ok[0] and abs(uu[0] - uv[q, 0]) < 1e-4 and abs(vv[0] - uv[q, 1]) < 1e-4
