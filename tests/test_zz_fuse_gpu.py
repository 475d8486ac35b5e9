"""Boundary B7 on a B200: lvba_tracks_fuse_create / _export (global-lvba_b200/csrc/fuse_api.cuh — BuildTracksAndFuse3D, reference
src/lvba_system.cpp:921-1263) through the C ABI against the literal restatement oracle/fuse_oracle.py: same tracks in the same
order, same observations, same inlier sets, same candidate choice, points to 1e-9; then the chain the reference runs —
fused tracks -> observation CSR -> lvba_visual_lm.  CPU twin of the functor: tests/test_fuse_emu.py."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import fuse_scene  # noqa: E402
from oracle import fuse_oracle as fo  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as graft
    p = graft.load_package(); p.load_library()
    if p.device_count() < 1:
        pytest.fail("no CUDA device: the LVBA hot path has no CPU fallback")
    return p


def compare(got, ref):
    assert len(got["source"]) == len(ref)
    for i, t in enumerate(ref):
        a, b = got["obs_ptr"][i], got["obs_ptr"][i + 1]
        assert np.array_equal(got["img"][a:b], t["obs"][:, 0]) and np.array_equal(got["kp"][a:b], t["obs"][:, 1])
        assert got["source"][i] == t["source"]
        assert np.array_equal(got["inlier"][a:b].astype(bool), t["inlier"])
        assert np.abs(got["Xw"][i] - t["Xw"]).max() <= 1e-9 * max(1.0, np.abs(t["Xw"]).max())
        assert abs(got["mean"][i] - t["mean"]) <= 1e-9 * max(1.0, t["mean"])


@pytest.mark.parametrize("seed,kw", [(0, {}), (1, dict(wrong=0.15)), (3, dict(bad_depth=0.3, px_noise=1.2)), (4, dict(n_images=30, n_points=300)),
                                     (5, dict(no_depth=1.0))])
def test_fusion_matches_the_oracle(pkg, seed, kw):
    s = fuse_scene.make(seed=seed, **kw)
    ref = fo.fuse(s["kp_ptr"], s["kp_uv"], s["matches"], s["cams"], s["intr"], s["kp_Xw"], s["kp_valid"], map_order=fo.ascending_order)
    got = pkg.tracks_fuse(s["kp_ptr"], s["kp_uv"], s["matches"], s["cams"], s["intr"], s["kp_Xw"], s["kp_valid"], map_order=pkg.FUSE_ORDER_ASCENDING)
    compare(got, ref)
    sm = got["summary"]
    assert sm["n_tracks"] == len(ref) and sm["n_attempts"] >= sm["n_candidates"] and sm["kernel_launches"] == sm["n_rounds"]


def test_fused_tracks_feed_the_visual_lm(pkg):
    """the reference's chain: BuildTracksAndFuse3D -> optimizeCameraPoses (inlier observations only, :1610-1617)"""
    s = fuse_scene.make(seed=7, n_images=24, n_points=400, wrong=0.02)
    got = pkg.tracks_fuse(s["kp_ptr"], s["kp_uv"], s["matches"], s["cams"], s["intr"], s["kp_Xw"], s["kp_valid"], map_order=pkg.FUSE_ORDER_ASCENDING)
    n = len(got["source"])
    assert n > 50
    keep = got["inlier"].astype(bool)
    cnt = np.add.reduceat(keep.astype(np.int64), got["obs_ptr"][:-1]) if n else np.zeros(0, np.int64)
    obs_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    g = s["kp_ptr"][got["img"][keep]] + got["kp"][keep]
    obs_cam = got["img"][keep].astype(np.int32); obs_uv = s["kp_uv"][g]
    M = len(s["kp_ptr"]) - 1
    q = np.zeros((M, 4)); t = np.zeros((M, 3))
    for i in range(M):                                           # Rcw -> quaternion (w, x, y, z)
        R = s["cams"][i, :9].reshape(3, 3)
        w = np.sqrt(max(0.0, 1 + np.trace(R))) / 2
        q[i] = [w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)]
        t[i] = s["cams"][i, 9:]
    nrm = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    plane = np.column_stack([nrm, -(nrm * got["Xw"]).sum(1)])
    _, _, _, sv = pkg.visual_lm(q, t, got["Xw"].copy(), plane, obs_ptr, obs_cam, obs_uv, s["intr"], 0.5, 0.01)
    assert sv["iterations"] >= 1 and sv["cost_last"] <= sv["cost_first"]


def test_shim_mirror_of_BuildTracksAndFuse3D(pkg, tmp_path):
    """host/lvba_shim.hpp::build_tracks_and_fuse_3d with mock keypoint / match-table types through the C ABI"""
    import subprocess
    exe = tmp_path / "test_shim"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "shim" / "test_shim.cpp"),
           "-o", str(exe), str(pkg.LIB_PATH), f"-Wl,-rpath,{pkg.LIB_PATH.parent}", "-L/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), "fuse"], capture_output=True, text=True)
    assert r.returncode == 0 and "fuse ok: 12 tracks" in r.stdout, r.stdout + r.stderr
