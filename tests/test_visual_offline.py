"""The visual side of the offline tool without a GPU: the reference's YAML, the image list with its sampling
step, the COLMAP database reader (loadFromColmapDB, src/lvba_system.cpp:510-685: names -> ids, float32 keypoint blobs, inlier
matches stored by ascending database id) and updateCameraPosesFromLidar (:412-446), through `lvba_offline --check --visual`."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import dataset_writer as dw
import sys
sys.path.insert(0, str(Path(__file__).resolve().parent))
import visual_scene  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def tool(pkg, tmp_path_factory):
    exe = tmp_path_factory.mktemp("offline_visual") / "lvba_offline"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tools" / "lvba_offline.cpp"), "-o", str(exe),
           str(pkg.LIB_PATH), f"-Wl,-rpath,{pkg.LIB_PATH.parent}", "-L/usr/local/cuda/lib64", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def expected_cam_sum(sc, lidar_opt):
    """update_camera_poses_from_lidar + world_to_camera in numpy"""
    w = np.concatenate([np.arange(1, 10), np.arange(10, 13)]).astype(np.float64)
    ts = np.array(sc["ts"])
    tot = 0.0
    for i, t_img in enumerate(sc["image_ts"]):
        idx = int(np.searchsorted(ts, t_img, side="left"))
        if idx == len(ts):
            idx = len(ts) - 1
        elif idx > 0 and abs(ts[idx - 1] - t_img) < abs(ts[idx] - t_img):
            idx -= 1
        Ro, po = lidar_opt[idx, :9].reshape(3, 3), lidar_opt[idx, 9:]
        Rb, pb = sc["poses"][idx, :9].reshape(3, 3), sc["poses"][idx, 9:]
        Rd = Ro @ Rb.T
        pd = po - Rd @ pb
        q = dw.R_to_quat(sc["image_poses"][i, :9].reshape(3, 3))
        Rc = dw.quat_to_R(q)                                               # the pose file round trip
        Rwi = Rd @ Rc
        pwi = Rd @ sc["image_poses"][i, 9:] + pd
        Rcw = visual_scene.RCL @ Rwi.T
        tcw = -Rcw @ pwi + visual_scene.PCL
        tot += float(w[:9] @ Rcw.ravel() + w[9:] @ tcw)
    return tot


def test_check_mode_reads_images_database_and_config(tool, tmp_path):
    sc = visual_scene.make(tmp_path, seed=4, W=6, n_per_scan=800, n_landmarks=60)
    # a LiDAR result that differs from the odometry: the camera poses must follow it
    rng = np.random.default_rng(1)
    opt = sc["poses"].copy()
    opt[:, 9:] += rng.normal(0, 0.05, (len(opt), 3))
    with open(tmp_path / "opt.txt", "w") as f:
        for i in range(len(opt)):
            q = dw.R_to_quat(opt[i, :9].reshape(3, 3))
            f.write(f"{sc['ts'][i]:.6f} {opt[i, 9]:.12f} {opt[i, 10]:.12f} {opt[i, 11]:.12f} {q[1]:.15f} {q[2]:.15f} {q[3]:.15f} {q[0]:.15f}\n")
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--config", str(tmp_path / "config.yaml"), "--lidar-opt", str(tmp_path / "opt.txt"),
                        "--check", "--visual"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["images"] == 6 and abs(info["first_image"] - sc["image_ts"][0]) < 1e-9          # every 2nd of the 12 files
    assert info["width"] == 80 and info["height"] == 64 and abs(info["fx"] - 0.5 * visual_scene.INTR_FULL[0]) < 1e-9
    assert info["keypoints"] == sum(len(k) for k in sc["keypoints"])
    want = sum(float(k[:, 0].astype(np.float64).sum() + 2.0 * k[:, 1].astype(np.float64).sum()) for k in sc["keypoints"])
    assert abs(info["kp_sum"] - want) <= 1e-6 * abs(want)
    # matches come back per dataset pair (i < j) with columns (keypoint of i, keypoint of j), whatever the database ids were
    N = 6
    n_match = msum = 0
    for (i, j), m in sc["pair"].items():
        k = i * N - i * (i + 1) // 2 + (j - i - 1)
        for a, b in m:
            if a < len(sc["keypoints"][i]) and b < len(sc["keypoints"][j]):
                n_match += 1; msum += (k + 1) * (a + 3 * b)
    assert info["matches"] == n_match and info["match_sum"] == msum
    assert any(sc["db_ids"][i] > sc["db_ids"][j] for (i, j) in sc["pair"])                        # the swapped branch was exercised
    quat_opt = opt.copy()
    for i in range(len(opt)):
        quat_opt[i, :9] = dw.quat_to_R(dw.R_to_quat(opt[i, :9].reshape(3, 3))).ravel()
    assert abs(info["cam_sum"] - expected_cam_sum(sc, quat_opt)) <= 1e-8


def test_database_with_other_image_count_is_refused(tool, tmp_path):
    sc = visual_scene.make(tmp_path, seed=5, W=5, n_per_scan=500, n_landmarks=30)
    dw.write_colmap_db(tmp_path / "Colmap" / "colmap.db", sc["image_ts"][:-1], sc["keypoints"][:-1], {})
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--config", str(tmp_path / "config.yaml"), "--check", "--visual"], capture_output=True, text=True)
    assert r.returncode == 1 and "images count (4) != dataset images count (5)" in r.stderr


def test_visual_needs_a_config_and_bad_configs_are_refused(tool, tmp_path):
    sc = visual_scene.make(tmp_path, seed=6, W=4, n_per_scan=300, n_landmarks=20)
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--visual"], capture_output=True, text=True)
    assert r.returncode == 64 and "--config" in r.stderr
    bad = tmp_path / "bad.yaml"
    bad.write_text((tmp_path / "config.yaml").read_text().replace("  Pcl:", "  Pcl_missing:"))
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--config", str(bad), "--visual"], capture_output=True, text=True)
    assert r.returncode == 64 and "extrin_calib" in r.stderr
    (tmp_path / "all_image" / "image_poses.txt").write_text("# nothing\n")
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--config", str(tmp_path / "config.yaml"), "--check", "--visual"], capture_output=True, text=True)
    assert r.returncode == 1
