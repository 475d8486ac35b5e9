"""Regenerates tests/golden/small_problem.npz — inputs AND oracle outputs of a tiny LiDAR-visual problem.

The reference has no golden vectors, so these fixtures pin the *oracle* against itself
over time (regression) and give the CUDA path committed known-answer vectors that do not depend on
regenerating inputs on the GPU box.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import lidar_oracle as lo, synth, visual_oracle as vo  # noqa: E402

p = synth.make_problem(12, 80, 60, seed=4242)
W = 12
out = {k: v for k, v in p.items() if isinstance(v, np.ndarray)}
out["sigma_px"] = np.float64(p["sigma_px"]); out["sigma_plane"] = np.float64(p["sigma_plane"])
r, g, blocks = lo.acc_evaluate2(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], W)
out["A_residual_sum"] = np.float64(r); out["A_g"] = g; out["A_H"] = lo.assemble_dense(blocks, W)
out["A_residual_gt"] = np.float64(lo.only_residual(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses_gt"]))
poses, info = lo.damping_iter(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
out["A_lm_poses"] = poses; out["A_lm_cost_first"] = np.float64(info["r_first"]); out["A_lm_cost_last"] = np.float64(info["r_last"])
out["A_lm_iters"] = np.int64(info["iters"]); out["A_lm_accepted"] = np.int64(info["accepted"])
K = ("q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr", "sigma_px", "sigma_plane")
pr = vo.VisualProblem(*[p[k] for k in K])
st = vo.single_step(pr, 1e4, True)
out["B_cost0"] = np.float64(st["cost"]); out["B_model"] = np.float64(st["model"]); out["B_cam_step"] = st["cam_step"]; out["B_pt_step"] = st["pt_step"]
pr, info = vo.ceres_lm(pr)
out["B_lm_q"] = pr.q; out["B_lm_t"] = pr.t; out["B_lm_X"] = pr.X; out["B_lm_cost"] = np.float64(info["cost"])
out["B_lm_iters"] = np.int64(info["iters"]); out["B_lm_accepted"] = np.int64(info["accepted"])
np.savez_compressed(Path(__file__).with_name("small_problem.npz"), **out)

# ---- window BA (lvba_lidar_lm_batch): five small windows, one below the 3-voxels-per-pose rule, one empty
wp = synth.make_window_problem([8, 6, 5, 7, 4], [60, 40, 9, 0, 30], seed=99)
wposes, winfos = lo.window_ba(wp["win_ptr"], wp["vox_ptr"], wp["pose_idx"], wp["clusters"], wp["poses"])
wout = {k: wp[k] for k in ("win_ptr", "vox_ptr", "pose_idx", "clusters", "poses")}
wout["W_poses"] = wposes
wout["W_skipped"] = np.array([i is None for i in winfos])
wout["W_iters"] = np.array([0 if i is None else i["iters"] for i in winfos], np.int64)
wout["W_accepted"] = np.array([0 if i is None else i["accepted"] for i in winfos], np.int64)
wout["W_cost_first"] = np.array([0.0 if i is None else i["r_first"] for i in winfos])
wout["W_cost_last"] = np.array([0.0 if i is None else i["r_last"] for i in winfos])
np.savez_compressed(Path(__file__).with_name("window_problem.npz"), **wout)
print("wrote window_problem.npz", wout["W_iters"], wout["W_accepted"], wout["W_skipped"])
print("wrote", Path(__file__).with_name("small_problem.npz"), {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items() if k.startswith(('A_', 'B_'))})
