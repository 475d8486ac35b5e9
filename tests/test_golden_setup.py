"""Committed known-answer vectors of the set-up stages B4 (depth images), B6 (anchor clouds) and B7 (fused tracks):
tests/golden/setup_stages.npz, written by tests/golden/make_golden_setup.py from the LITERAL restatements of the reference lines.
Without a GPU: (1) the vectorised oracles still reproduce the file (the oracles are pinned over time), (2) the device passes, run
through the sequential host policy (tests/emu/), reproduce it — images and clouds bit for bit, tracks observation for observation.
The device itself is held against the same file in tests/test_zz_golden_gpu.py."""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))

from oracle import anchor_oracle as ao, depth_oracle as dep, fuse_oracle as fo  # noqa: E402
import test_anchor_emu, test_depth_emu, test_fuse_emu  # noqa: E402,E401

GOLD = np.load(ROOT / "tests" / "golden" / "setup_stages.npz")


def split(flat, ptr):
    return [flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]


def depth_scene():
    return dict(scans=split(GOLD["d_xyz"], GOLD["d_scan_ptr"]), poses=GOLD["d_poses"], frame_ts=GOLD["d_frame_ts"], cams=GOLD["d_cams"],
                image_ts=GOLD["d_image_ts"], intr=GOLD["d_intr"], width=int(GOLD["d_size"][0]), height=int(GOLD["d_size"][1]))


def fuse_scene_in():
    return {k: GOLD["f_" + k] for k in ("kp_ptr", "kp_uv", "matches", "cams", "intr", "kp_Xw", "kp_valid")}


def golden_tracks():
    p = GOLD["f_obs_ptr"]
    return [dict(obs=np.column_stack([GOLD["f_obs_img"][p[i]:p[i + 1]], GOLD["f_obs_kp"][p[i]:p[i + 1]]]), inlier=GOLD["f_inlier"][p[i]:p[i + 1]].astype(bool),
                 Xw=GOLD["f_Xw"][i], source=int(GOLD["f_source"][i]), mean=float(GOLD["f_mean"][i]), seed=int(GOLD["f_seed"][i])) for i in range(len(p) - 1)]


def _build(tmp, name):
    so = tmp / f"lib{name}.so"
    san = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []
    r = subprocess.run(["g++", "-std=c++17", "-O2", *san, "-ffp-contract=off", "-Wall", "-fPIC", "-shared", str(ROOT / "tests" / "emu" / f"{name}.cpp"), "-o", str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(str(so))


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("emu_golden")
    fuse = _build(tmp, "fuse_emu"); fuse.fuse_emu_run.restype = ctypes.c_longlong
    return dict(depth=_build(tmp, "depth_emu"), anchor=_build(tmp, "anchor_emu"), fuse=fuse)


def test_oracles_reproduce_the_fixture():
    s = depth_scene()
    img = dep.render(s["scans"], s["poses"], s["frame_ts"], s["cams"], s["image_ts"], s["intr"], s["width"], s["height"],
                     voxel_size=float(GOLD["d_voxel_size"]), half_window=float(GOLD["d_half_window"]))
    assert np.array_equal(img, GOLD["d_images"]) and np.count_nonzero(img) > 2000
    clouds = ao.anchor_clouds(split(GOLD["a_xyz"], GOLD["a_scan_ptr"]), GOLD["a_rel"], GOLD["a_win_ptr"], float(GOLD["a_leaf"]))
    assert all(np.array_equal(c, g) for c, g in zip(clouds, split(GOLD["a_cloud_xyz"], GOLD["a_cloud_ptr"])))
    f = fuse_scene_in()
    tracks = fo.fuse(f["kp_ptr"], f["kp_uv"], f["matches"], f["cams"], f["intr"], f["kp_Xw"], f["kp_valid"], map_order=fo.ascending_order)   # the file was written under LVBA_FUSE_ORDER_ASCENDING
    gold = golden_tracks()
    assert len(tracks) == len(gold) == 24
    for t, g in zip(tracks, gold):
        assert t["seed"] == g["seed"] and np.array_equal(t["obs"], g["obs"]) and np.array_equal(t["inlier"], g["inlier"]) and t["source"] == g["source"]
        assert np.abs(t["Xw"] - g["Xw"]).max() <= 1e-12 * max(1.0, np.abs(g["Xw"]).max()) and abs(t["mean"] - g["mean"]) <= 1e-12 * max(1.0, g["mean"])


def test_host_policy_runs_reproduce_the_fixture(libs):
    rc, got, info = test_depth_emu.emu_render(libs["depth"], depth_scene(), float(GOLD["d_voxel_size"]), float(GOLD["d_half_window"]))
    assert rc == 0 and np.array_equal(got, GOLD["d_images"])
    rc, clouds = test_anchor_emu.run(libs["anchor"], split(GOLD["a_xyz"], GOLD["a_scan_ptr"]), GOLD["a_rel"], GOLD["a_win_ptr"], float(GOLD["a_leaf"]))
    assert rc == 0 and all(np.array_equal(c, g) for c, g in zip(clouds, split(GOLD["a_cloud_xyz"], GOLD["a_cloud_ptr"])))
    test_fuse_emu.compare(test_fuse_emu.run_emu(libs["fuse"], fuse_scene_in()), golden_tracks())
