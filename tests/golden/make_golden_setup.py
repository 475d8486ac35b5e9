"""Regenerates tests/golden/setup_stages.npz — committed known-answer vectors for the set-up stages around the solves:
  B4  depth images            oracle/depth_oracle.py render_literal   (src/lvba_system.cpp:1266-1338, 835-919)
  B6  anchor clouds           oracle/anchor_oracle.py anchor_clouds_literal   (src/lvba_system.cpp:284-301, tools.hpp:301-359, 385-395)
  B7  fused tracks            oracle/fuse_oracle.py fuse              (src/lvba_system.cpp:921-1263)
each produced by the LITERAL restatement of the reference lines.  The reference ships no fixtures:
the file pins the oracles over time and gives the host-policy and device runs inputs + outputs that nothing regenerates.
Run from the repo root:  python tests/golden/make_golden_setup.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fuse_scene  # noqa: E402
from oracle import anchor_oracle as ao, depth_oracle as dep, fuse_oracle as fo, synth  # noqa: E402

out = {}
# ---- B4
s = synth.make_depth_scene(21, F=6, n_per_scan=1800, M=4)
sp = np.concatenate([[0], np.cumsum([len(x) for x in s["scans"]])]).astype(np.int64)
img = dep.render_literal(s["scans"], s["poses"], s["frame_ts"], s["cams"], s["image_ts"], s["intr"], s["width"], s["height"])
out.update(d_xyz=np.concatenate(s["scans"]).astype(np.float32), d_scan_ptr=sp, d_poses=s["poses"], d_frame_ts=s["frame_ts"], d_cams=s["cams"],
           d_image_ts=s["image_ts"], d_intr=s["intr"], d_size=np.array([s["width"], s["height"]]), d_voxel_size=np.array(0.5), d_half_window=np.array(0.5),
           d_images=img.astype(np.float32))
print("B4:", img.shape, "filled", float(np.mean(img > 0)))
# ---- B6
sizes = [3, 1, 4]
scans, poses = synth.make_scan_scene(22, W=sum(sizes), n_per_scan=1200)
scans[4] = scans[4][:0]
win_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
rel = ao.rel_poses(poses, win_ptr)
clouds = ao.anchor_clouds_literal(scans, rel, win_ptr, 0.2)
out.update(a_xyz=np.concatenate(scans).astype(np.float32), a_scan_ptr=np.concatenate([[0], np.cumsum([len(x) for x in scans])]).astype(np.int64),
           a_rel=rel, a_win_ptr=win_ptr, a_leaf=np.array(0.2), a_cloud_ptr=np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64),
           a_cloud_xyz=np.concatenate(clouds).astype(np.float32))
print("B6:", [len(c) for c in clouds], "points of", [sum(len(scans[j]) for j in range(win_ptr[w], win_ptr[w + 1])) for w in range(len(sizes))])
# ---- B7
f = fuse_scene.make(seed=23, n_images=12, n_points=90, wrong=0.1, bad_depth=0.15)
tracks = fo.fuse(f["kp_ptr"], f["kp_uv"], f["matches"], f["cams"], f["intr"], f["kp_Xw"], f["kp_valid"], map_order=fo.ascending_order)   # LVBA_FUSE_ORDER_ASCENDING (the container order of the
# reference is pinned by tests/golden/ref_system.npz, written by the reference's own source)
obs = np.concatenate([t["obs"] for t in tracks]).astype(np.int32)
out.update(f_kp_ptr=f["kp_ptr"], f_kp_uv=f["kp_uv"], f_matches=f["matches"], f_cams=f["cams"], f_intr=f["intr"], f_kp_Xw=f["kp_Xw"], f_kp_valid=f["kp_valid"],
           f_obs_ptr=np.concatenate([[0], np.cumsum([len(t["obs"]) for t in tracks])]).astype(np.int64), f_obs_img=obs[:, 0], f_obs_kp=obs[:, 1],
           f_inlier=np.concatenate([t["inlier"] for t in tracks]).astype(np.uint8), f_Xw=np.array([t["Xw"] for t in tracks]),
           f_source=np.array([t["source"] for t in tracks], np.uint8), f_mean=np.array([t["mean"] for t in tracks]),
           f_seed=np.array([t["seed"] for t in tracks], np.int64))
print("B7:", len(tracks), "tracks, sources", np.bincount(out["f_source"], minlength=3))
np.savez_compressed(Path(__file__).with_name("setup_stages.npz"), **out)
