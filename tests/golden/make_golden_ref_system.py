"""Writes tests/golden/ref_system.npz — seeded inputs and what THE REFERENCE'S PIPELINE SOURCE computed from them.

The outputs come from oracle/_ref/liblvba_system_ref.so: /root/reference/src/lvba_system.cpp (all of it, unmodified) compiled where it
lies and driven through LvbaSystem's own public members (oracle/ref_system_driver.cpp, `make -C oracle ref`), on the stand-in library
headers of oracle/ref_shim/ (what they are and what they leave open: ref_shim/mini_eigen.h).  /root/reference does not
exist on the GPU box, so the vectors are committed; tests/test_ref_system_pin.py holds the oracles and the host-policy runs of the device
passes against them and regenerates the file bit for bit where the library can be built; tests/test_zzz_ref_gpu.py holds the device.

Sections:  L  runWindowBA + runLidarBA (window stage with a skipped window, anchors, the two global stages, poses of every frame)
           D  buildGridMapFromOptimized -> updateCameraPosesFromLidar -> generateDepthWithVoxel (after a LiDAR correction)
           F  BuildTracksAndFuse3D on depth images + keypoints + matches (track order, observation order, inlier order, points)
           P  the Ceres problem optimizeCameraPoses builds (recorded, not solved)
           V  the whole camera half end to end (reference reader + fusion + problem, the restated Ceres loop in the solver hook): results only

Run from the repo root (only where /root/reference exists):   python tests/golden/make_golden_ref_system.py
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import lvba_system_ref as sr, synth  # noqa: E402
import visual_scene as vs  # noqa: E402

OUT = Path(__file__).with_name("ref_system.npz")


def lex(a):
    return a[np.lexsort(a.T[::-1])]


def flat(scans):
    ptr = np.zeros(len(scans) + 1, np.int64); ptr[1:] = np.cumsum([len(s) for s in scans])
    return ptr, np.concatenate(scans).astype(np.float32)


def lidar_scene():
    W, win = 9, 3
    scans, poses = synth.make_scan_scene(23, W=W, n_per_scan=2000)
    rng = np.random.default_rng(8)
    noisy = poses.copy()
    for i in range(W):
        noisy[i, :9] = (noisy[i, :9].reshape(3, 3) @ synth.so3_exp(rng.normal(0, 0.004, (1, 3)))[0]).ravel()
        noisy[i, 9:] += rng.normal(0, 0.01, 3)
    for i in range(3, 6):                        # the middle window sees too little: fewer than 3 plane voxels per pose -> skipped (:259-263)
        scans[i] = scans[i][:40]
    return scans, noisy, win


def generate():
    o = {}
    # ---------------------------------------------------------------- L
    scans, noisy, win = lidar_scene()
    o["L_scan_ptr"], o["L_xyz"] = flat(scans)
    o["L_poses"] = noisy; o["L_window"] = np.int64(win); o["L_anchor_leaf"] = np.float64(0.1)
    o["L_s1_voxel"] = np.float64(1.0); o["L_s1_ratio"] = np.array([0.3, 0.1, 0.06, 0.03], np.float32)
    o["L_s2_voxel"] = np.float64(1.0); o["L_s2_ratio"] = np.array([0.08, 0.08, 0.08, 0.08], np.float32)
    DEFAULT_RATIO = np.array([0.3, 0.1, 0.06, 0.03], np.float32)   # bavoxel.hpp:17 — what a fresh process has when the window stage runs (runLidarBA
    sr.set_eigen_ratio_array(DEFAULT_RATIO)                        # sets the configured arrays only before the global stages, :358)
    S = sr.System()
    S.set_lidar(scans, noisy)
    S.set_stages(True, win, 0.1, False, True, 1.0, o["L_s1_ratio"], 1.0, o["L_s2_ratio"])
    ap, ac, rel, idx = S.run_window_ba()
    o["L_anchor_poses"] = ap; o["L_anchor_index"] = idx; o["L_rel_poses"] = rel
    o["L_anchor_cloud_ptr"] = np.concatenate([[0], np.cumsum([len(c) for c in ac])]).astype(np.int64)
    o["L_anchor_clouds_sorted"] = np.concatenate([lex(c) for c in ac])
    sr.set_eigen_ratio_array(DEFAULT_RATIO)
    o["L_final_poses"] = S.run_lidar_ba()
    S.close()
    sr.set_eigen_ratio_array(DEFAULT_RATIO)                        # a fresh process again (the run above left stage 2's array behind)
    # the same with the window stage's own relative poses (use_window_ba_rel, :267-279)
    S = sr.System()
    S.set_lidar(scans, noisy)
    S.set_stages(True, win, 0.1, True, True, 1.0, o["L_s1_ratio"], 1.0, o["L_s2_ratio"])
    _, ac_rel, o["L_rel_poses_rel"], _ = S.run_window_ba()          # rel = anchor^-1 * (window-LM pose aligned to the anchor): the window solves, frame by frame
    o["L_anchor_clouds_rel_sorted"] = np.concatenate([lex(c) for c in ac_rel])
    sr.set_eigen_ratio_array(DEFAULT_RATIO)
    o["L_final_poses_rel"] = S.run_lidar_ba()
    S.close()

    # ---------------------------------------------------------------- D
    sc = synth.make_depth_scene(7, F=6, n_per_scan=2500, M=4, width=96, height=72)
    img_ts = np.round(sc["image_ts"], 6)                     # the reference goes through std::to_string (6 decimals, :1310-1311)
    rng = np.random.default_rng(3)
    Rcl = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]]) @ synth.so3_exp(np.array([[0.05, -0.03, 0.02]]))[0]
    tcl = np.array([0.03, -0.02, 0.05])
    Ril = synth.so3_exp(np.array([[0.01, 0.02, -0.015]]))[0]; til = np.array([0.01, -0.005, 0.02])
    M = len(img_ts)
    img_pose = np.zeros((M, 12))
    for k in range(M):
        f = int(np.clip(np.searchsorted(sc["frame_ts"], img_ts[k]), 0, len(sc["frame_ts"]) - 1))
        R = sc["poses"][f, :9].reshape(3, 3) @ synth.so3_exp(rng.normal(0, 0.01, (1, 3)))[0]
        img_pose[k, :9] = R.ravel(); img_pose[k, 9:] = sc["poses"][f, 9:] + rng.normal(0, 0.02, 3)
    before = sc["poses"].copy()
    for i in range(len(before)):                             # the odometry the LiDAR half started from: the optimised poses moved away from it
        before[i, :9] = (before[i, :9].reshape(3, 3) @ synth.so3_exp(rng.normal(0, 0.006, (1, 3)))[0]).ravel()
        before[i, 9:] += rng.normal(0, 0.015, 3)
    o["D_scan_ptr"], o["D_xyz"] = flat(sc["scans"])
    o["D_poses"] = sc["poses"]; o["D_poses_before"] = before; o["D_frame_ts"] = sc["frame_ts"]; o["D_image_ts"] = img_ts; o["D_image_poses"] = img_pose
    o["D_intr"] = sc["intr"]; o["D_size"] = np.array([sc["width"], sc["height"]], np.int64)
    o["D_Rcl"] = Rcl; o["D_tcl"] = tcl; o["D_Ril"] = Ril; o["D_til"] = til
    S = sr.System()
    S.set_lidar(sc["scans"], before, sc["frame_ts"])          # x_buf_before_ = x_buf_ = the odometry ...
    S.set_lidar_optimised(sc["poses"])                        # ... then x_buf_ = the LiDAR result, as runLidarBA leaves it (:405)
    S.set_stages(False)
    S.set_camera(sc["width"], sc["height"], sc["intr"], Rcl.ravel(), tcl, Ril.ravel(), til, img_ts, img_pose)
    S.build_grid()
    o["D_camera_body_poses"] = S.update_camera_poses()
    d, c0, c1 = S.generate_depth()
    o["D_depth"] = d; o["D_cams_odometry"] = c0; o["D_cams"] = c1
    S.close()

    # ---------------------------------------------------------------- F + P (one scene: the fused tracks feed the problem)
    root = Path(tempfile.mkdtemp())
    g = vs.make(root, seed=3, W=8)
    S = sr.System()
    S.set_lidar(g["scans"], g["poses"], g["ts"])
    S.set_stages(False, window_size=4, anchor_leaf=0.1, s2_voxel=0.5, s2_ratio=(0.08,) * 4)
    S.set_camera(g["width"], g["height"], g["intr"], vs.RCL.ravel(), vs.PCL, np.eye(3).ravel(), np.zeros(3), np.array(g["image_ts"]), g["image_poses"])
    S.build_grid(); S.update_camera_poses()
    d, c0, c1 = S.generate_depth()
    kp_ptr = np.concatenate([[0], np.cumsum([len(k) for k in g["keypoints"]])]).astype(np.int64)
    kp_uv = np.concatenate(g["keypoints"]).astype(np.float32)
    M = len(g["keypoints"])
    matches = np.array([(a, ka, b, kb) for a in range(M) for b in range(a + 1, M) for (ka, kb) in g["pair"].get((a, b), [])], np.int32)
    o["F_cams"] = c1; o["F_depth"] = d; o["F_intr"] = g["intr"]; o["F_kp_ptr"] = kp_ptr; o["F_kp_uv"] = kp_uv; o["F_matches"] = matches
    S.set_keypoints_and_matches(kp_ptr, kp_uv, matches)
    T = S.build_tracks()
    for k in ("obs_ptr", "obs", "inl_ptr", "inl", "Xw"):
        o["F_" + k] = T[k]
    sr.set_eigen_ratio_array(o["L_s2_ratio"])                 # what the LiDAR half's last stage leaves in the process-wide array (:358)
    rec, _ = S.optimize_camera_poses(None)                    # record only: the parameter blocks stay as they were
    for k, v in rec.items():
        o["P_" + k] = v
    o["P_options"] = rec["options"][:5]                       # [5] is std::thread::hardware_concurrency() of the machine
    o["P_window"] = np.int64(4); o["P_anchor_leaf"] = np.float64(0.1); o["P_voxel"] = np.float64(0.5); o["P_ratio"] = np.array([0.08] * 4, np.float32)
    # ---------------------------------------------------------------- V: the whole camera half, end to end, on the scene of tests/test_zz_offline_gpu.py
    # runVisualBAWithLidarAssist (:144-154) with LiDAR BA disabled: the reference's own dataset loader -> grid -> camera poses -> depth -> its own COLMAP reader -> track
    # fusion -> optimizeCameraPoses; where ceres::Solve stands, the restated Ceres loop (oracle/visual_oracle.py) solves the recorded problem and the
    # reference's own code writes the result back.  Only results are stored: the test rebuilds the (seeded) scene.
    from oracle import visual_oracle as vis
    rootv = Path(tempfile.mkdtemp())
    gv = vs.make(rootv, seed=3, W=8, n_landmarks=700)
    F = vs.INTR_FULL                                          # the reference's own loader reads the directory (src/dataset_io.cpp): what config.yaml says, as ROS parameters
    S2 = sr.System({"data_config/data_path": str(rootv) + "/", "data_config/colmap_db_path": "Colmap/colmap.db", "data_config/image_sample_step": 2,
                    "cam_model/cam_width": vs.WIDTH_FULL, "cam_model/cam_height": vs.HEIGHT_FULL, "cam_model/scale": vs.SCALE, "cam_model/cam_fx": F[0],
                    "cam_model/cam_fy": F[1], "cam_model/cam_cx": F[2], "cam_model/cam_cy": F[3], "cam_model/cam_d0": F[4], "cam_model/cam_d1": F[5],
                    "cam_model/cam_d2": F[6], "cam_model/cam_d3": F[7], "extrin_calib/Rcl": vs.RCL.ravel(), "extrin_calib/Pcl": vs.PCL,
                    "extrin_calib/extrinsic_R": np.eye(3).ravel(), "extrin_calib/extrinsic_T": np.zeros(3),
                    "window_ba/enable": 0, "window_ba/size": 10, "window_ba/anchor_leaf_size": 0.05, "BALM_stage2/root_voxel_size": 0.5})
    dsv = S2.init_from_dataset()
    assert len(dsv["frame_ts"]) == 8 and len(dsv["image_ts"]) == 8
    S2.build_grid(); S2.update_camera_poses()
    _, c0v, c1v = S2.generate_depth()
    ok, _, _, _ = S2.load_colmap_db(str(rootv) + "/", rootv / "Colmap" / "colmap.db")
    assert ok
    Tv = S2.build_tracks()
    sr.set_eigen_ratio_array(DEFAULT_RATIO)                    # enable_lidar_ba = false: set_eigen_ratio_array never ran (bavoxel.hpp:17 in force)
    info = {}

    def solver(P):
        obs_ptr = np.concatenate([[0], np.cumsum(np.bincount(P["obs_pt"], minlength=len(P["X"])))])
        nd = np.zeros((len(P["X"]), 4)); nd[P["pl_pt"]] = P["pl_nd"]
        pr = vis.VisualProblem(P["q"], P["t"], P["X"], nd, obs_ptr, P["obs_cam"], P["obs_uv"].astype(np.float32), P["obs_intr"][0],
                               float(P["obs_sigma"][0, 0]), float(P["pl_sigma"][0]))
        info["cost_first"] = pr.cost()
        pr, lm = vis.ceres_lm(pr, max_iter=int(P["options"][0]))
        info.update(cost_last=lm["cost"], iters=lm["iters"], points=len(P["X"]))
        return pr.q, pr.t, pr.X

    _, cams_after = S2.optimize_camera_poses(solver)
    o["V_cams_before"] = c1v; o["V_cams_after"] = cams_after
    o["V_counts"] = np.array([len(Tv["Xw"]), info["points"], info["iters"]], np.int64)
    o["V_costs"] = np.array([info["cost_first"], info["cost_last"]])
    S2.close()
    # the anchor clouds the surf map of optimizeCameraPoses is cut from are not stored (2.9 MB of scans behind them): the live test rebuilds the scene
    S.close()
    return o


if __name__ == "__main__":
    assert sr.available(), "oracle/_ref/liblvba_system_ref.so is missing: run `make -C oracle ref` where /root/reference exists"
    out = generate()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(out), "arrays")
    print("L anchors", out["L_anchor_index"].tolist(), "F tracks", len(out["F_Xw"]), "P pts", len(out["P_X"]), "obs", len(out["P_obs_cam"]))
