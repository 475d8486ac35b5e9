"""Writes tests/golden/ref_balm.npz — seeded inputs and the outputs of THE REFERENCE'S OWN SOURCE for the hot path.

The outputs come from oracle/_ref/libbalm_ref.so, i.e. /root/reference/include/BALM/{tools,bavoxel}.hpp and
/root/reference/include/utils.hpp compiled where they lie (oracle/ref_driver.cpp, `make -C oracle ref`) on top of
stand-ins for the libraries this image lacks (oracle/ref_shim/: own 3x3 eigen-solver, sparse LDL^T, dual numbers — see
ref_shim/mini_eigen.h for what that does and does not pin).  /root/reference does not exist on the GPU box, so the
vectors are committed; tests/test_ref_pin.py holds the oracles (numpy and C++) against them and, where the library can
be built, regenerates them and demands the identical bits; tests/test_zzz_ref_gpu.py holds the CUDA path against them.

Run from the repo root (only where /root/reference exists):   python tests/golden/make_golden_ref.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import anchor_oracle as ao, balm_ref as br, synth, visual_oracle as vis  # noqa: E402

OUT = Path(__file__).with_name("ref_balm.npz")


def lex(a):
    return a[np.lexsort(a.T[::-1])]


def generate():
    o = {}
    # ---------------------------------------------------------------- path A: two CSR problems (W = 14 and W = 24)
    for tag, (W, V, seed) in (("L1", (14, 120, 777)), ("L2", (24, 260, 778))):
        p = synth.make_problem(W, V, 0, seed=seed, visual=False)
        a = (p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
        for k in ("vox_ptr", "pose_idx", "clusters", "poses", "poses_gt"):
            o[f"{tag}_{k}"] = p[k]
        res, g, H, kept = br.lidar_hessian(*a)                       # ONE acc_evaluate2 call: sum of lambda_0
        res_t, g_t, H_t, _ = br.lidar_hessian(*a, threads=True)      # divide_thread: 16 threads, sum / kept
        o[f"{tag}_kept"] = np.int64(kept)
        o[f"{tag}_residual_sum"] = np.float64(res); o[f"{tag}_g"] = g; o[f"{tag}_H"] = H
        o[f"{tag}_residual_avg_threads"] = np.float64(res_t); o[f"{tag}_g_threads"] = g_t
        o[f"{tag}_residual_gt"] = np.float64(br.lidar_residual(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses_gt"]))
        lm = br.lidar_damping_iter(*a)
        o[f"{tag}_lm_poses"] = lm
        o[f"{tag}_lm_residual_sum"] = np.float64(br.lidar_residual(p["vox_ptr"], p["pose_idx"], p["clusters"], lm))
    # a voxel seen from one pose only and an empty voxel: push_voxel drops both (bavoxel.hpp:45-54)
    p = synth.make_problem(8, 30, 0, seed=779, visual=False)
    vp, pi, cl = p["vox_ptr"], p["pose_idx"], p["clusters"]
    keep1 = np.r_[np.arange(vp[0], vp[0] + 1), np.arange(vp[1], vp[-1])]            # voxel 0 keeps one slot
    vp_e = np.r_[0, 1, 1 + (vp[2:] - vp[1]), 1 + (vp[-1] - vp[1])].astype(np.int64)  # ... and an empty voxel at the end
    o["L3_vox_ptr"] = vp_e; o["L3_pose_idx"] = pi[keep1]; o["L3_clusters"] = cl[keep1]; o["L3_poses"] = p["poses"]
    res, g, H, kept = br.lidar_hessian(vp_e, pi[keep1], cl[keep1], p["poses"])
    o["L3_kept"] = np.int64(kept); o["L3_residual_sum"] = np.float64(res); o["L3_g"] = g; o["L3_H"] = H
    w = np.array([[0, 0, 0], [1e-12, 0, 0], [0.3, -0.2, 0.1], [3.0, 0.1, -0.4], [1e-9, 2e-9, -1e-9]])
    o["E_w"] = w; o["E_R"] = br.so3_exp(w)

    # ---------------------------------------------------------------- set-up: cut_voxel -> recut -> tras_opt, lookup, window LM
    scans, poses = synth.make_scan_scene(21, W=5, n_per_scan=1500)
    ptr = np.zeros(len(scans) + 1, np.int64); ptr[1:] = np.cumsum([len(s) for s in scans])
    o["M_scan_ptr"] = ptr; o["M_xyz"] = np.concatenate(scans).astype(np.float32); o["M_poses"] = poses
    o["M_voxel_size"] = np.float64(1.0); o["M_eigen_ratio"] = np.array([0.3, 0.1, 0.06, 0.03], np.float32)
    m = br.Map(scans, poses, 1.0, o["M_eigen_ratio"])
    vp, pi, cl, meta = m.export()
    o["M_vox_ptr"] = vp; o["M_pose_idx"] = pi; o["M_clusters"] = cl; o["M_key"] = meta["key"]; o["M_layer"] = meta["layer"]
    o["M_path"] = np.array([list(q) + [-1] * (2 - len(q)) for q in meta["path"]], np.int32).reshape(-1, 2)
    o["M_centre"] = meta["centre"]; o["M_direct"] = meta["direct"]; o["M_eigenvalues"] = meta["eigenvalues"]
    rng = np.random.default_rng(5)
    X = rng.uniform(-3.5, 3.5, (600, 3)); X[7] = np.nan; X[8, 1] = np.inf; X[9] = [50.0, 50.0, 50.0]
    st, d, c = m.lookup(X)
    o["M_lookup_X"] = X; o["M_lookup_state"] = st; o["M_lookup_direct"] = d; o["M_lookup_centre"] = c
    o["M_lm_poses"] = m.damping_iter()
    m.close()
    # anchor cloud of the same scans (tail of the window loop)
    win_ptr = np.array([0, len(scans)], np.int64)
    rel = ao.rel_poses(poses, win_ptr)
    o["A_rel"] = rel; o["A_leaf"] = np.float64(0.1)
    o["A_cloud_sorted"] = lex(br.anchor_cloud(scans, rel, 0.1))

    # ---------------------------------------------------------------- path B: the two cost functors on a whole problem
    p = synth.make_problem(10, 0, 120, seed=780, lidar=False)
    for k in ("q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr"):
        o[f"V_{k}"] = p[k]
    o["V_sigma_px"] = np.float64(p["sigma_px"]); o["V_sigma_plane"] = np.float64(p["sigma_plane"])
    trk = np.repeat(np.arange(len(p["obs_ptr"]) - 1), np.diff(p["obs_ptr"]))
    r, J = br.reproj(p["q"][p["obs_cam"]], p["t"][p["obs_cam"]], p["X"][trk], p["obs_uv"], p["intr"], p["sigma_px"], p["sigma_px"])
    o["V_reproj_r"] = r; o["V_reproj_J"] = J
    r, J = br.point_plane(p["X"], p["plane_nd"], p["sigma_plane"])
    o["V_plane_r"] = r; o["V_plane_J"] = J
    # observations behind / at the camera plane (utils.hpp:78) and a non-unit quaternion (QuaternionRotatePoint normalises)
    q = p["q"][p["obs_cam"]][:12].copy(); t = p["t"][p["obs_cam"]][:12].copy(); Xo = p["X"][trk][:12].copy()
    Xo[:4] = -Xo[:4]; q[4:8] *= 1.7
    for row, zc in ((8, 0.5e-8), (9, 0.9e-8), (10, 2.0e-8)):       # z_c below, below, above the 1e-8 threshold of utils.hpp:78
        Rm = vis.quat_to_rot(q[row]); Xo[row] = Rm.T @ (np.array([1e-9, -2e-9, zc]) - t[row])
    o["V_edge_q"] = q; o["V_edge_t"] = t; o["V_edge_X"] = Xo; o["V_edge_uv"] = p["obs_uv"][:12]
    r, J = br.reproj(q, t, Xo, p["obs_uv"][:12], p["intr"], p["sigma_px"], p["sigma_px"])
    o["V_edge_r"] = r; o["V_edge_J"] = J

    # ---------------------------------------------------------------- camera helpers (B4 / B7)
    sc = synth.make_depth_scene(4)
    rng = np.random.default_rng(6)
    Xc = rng.normal(0, 2, (800, 3)); Xc[:, 2] = np.abs(Xc[:, 2]) + 0.05
    Xc[:10, 2] *= -1; Xc[10] = np.nan; Xc[11, 2] = 1e-13; Xc[12, 2] = 0.0
    ok, uv, z = br.project_camera_to_pixel(sc["intr"], Xc)
    o["C_intr"] = sc["intr"]; o["C_Xc"] = Xc; o["C_proj_ok"] = ok; o["C_proj_uv"] = uv; o["C_proj_z"] = z
    uvp = np.column_stack([rng.uniform(-40, sc["width"] + 40, 600), rng.uniform(-40, sc["height"] + 40, 600)])
    uvp[3] = np.nan
    ok, xy = br.undistort_pixel(sc["intr"], uvp)
    o["C_und_uv"] = uvp; o["C_und_ok"] = ok; o["C_und_xy"] = xy
    h, w_ = 30, 40
    depth = rng.uniform(0.5, 6.0, (h, w_)).astype(np.float32)
    depth[rng.uniform(size=(h, w_)) < 0.15] = 0.0                                   # holes: any zero corner rejects the keypoint
    kp = np.column_stack([rng.uniform(-1.5, w_ + 0.5, 700), rng.uniform(-1.5, h + 0.5, 700)]).astype(np.float32)
    kp[0] = [0.0, 0.0]; kp[1] = [w_ - 1, 3.0]; kp[2] = [3.0, h - 1]      # (a NaN keypoint is undefined behaviour in utils.hpp:249-251: not a vector)
    ok, d, Xw = br.depth_candidate(depth, sc["intr"], sc["cams"][1], kp)
    o["C_depth"] = depth; o["C_cam"] = sc["cams"][1]; o["C_kp"] = kp; o["C_cand_ok"] = ok; o["C_cand_d"] = d; o["C_cand_Xw"] = Xw
    return o


if __name__ == "__main__":
    assert br.available(), "oracle/_ref/libbalm_ref.so is missing: run `make -C oracle ref` where /root/reference exists"
    out = generate()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(out), "arrays")
    print({k: (int(out[k]) if out[k].ndim == 0 and out[k].dtype.kind == "i" else out[k].shape) for k in out if k.endswith(("kept", "_H", "vox_ptr", "cand_ok"))})
