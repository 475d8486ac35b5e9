"""The CUDA path against what THE REFERENCE'S OWN SOURCE computed (tests/golden/ref_balm.npz; how it was made and what the
stand-in libraries under it do and do not pin: tests/golden/make_golden_ref.py, tests/test_ref_pin.py, oracle/ref_shim/mini_eigen.h).
Nothing is regenerated on the GPU box and /root/reference is not read: the committed vectors are the reference here.

Through the C ABI: B1 Hessian build / residual pass / full damping_iter, B3 voxel map + plane lookup + the window solve on the map
without the clusters leaving HBM, B6 anchor clouds, B2 cost (the two Ceres functors summed over the problem).
Tolerances as in tests/test_lidar_gpu.py (lambda_0 is a difference of O(1e4) terms): residual sums 1e-8 relative,
g / H 1e-7 of their largest entry, LM end poses 1e-6, final cost 1e-6 (north star); integer data, keys, point counts and the
float32 anchor points exact.  Runs in a child process under a timeout like the other test_zz_* files and sorts after all of them.  All tests
below have passed on a B200 (profiles/r02_ref_pin_gpu.txt, profiles/r02_ref_pin_gpu_system.txt)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

CODE = """
import sys
import numpy as np
sys.path.insert(0, %r)
import __graft_entry__ as graft
pkg = graft.load_package(); pkg.load_library()
assert pkg.device_count() >= 1
G = np.load(%r)
split = lambda flat, ptr: [flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
lex = lambda a: a[np.lexsort(a.T[::-1])]
launches = 0

# ---- B1: acc_evaluate2 / evaluate_only_residual / damping_iter
for tag in ("L1", "L2", "L3"):
    vp, pi, cl, ps = G[tag + "_vox_ptr"], G[tag + "_pose_idx"], G[tag + "_clusters"], G[tag + "_poses"]
    W = len(ps)
    keep = np.diff(vp) >= 2                                   # push_voxel (bavoxel.hpp:45-54) is the caller's side of B1
    assert int(keep.sum()) == int(G[tag + "_kept"])
    if not keep.all():
        sel = np.concatenate([np.arange(vp[a], vp[a + 1]) for a in np.nonzero(keep)[0]])
        vp = np.concatenate([[0], np.cumsum(np.diff(vp)[keep])]).astype(np.int64); pi = pi[sel]; cl = cl[sel]
    P = pkg.LidarProblem(vp, pi, cl, ps)
    r = P.build()
    g, br, bc, bl = P.get_system()
    H = pkg.env_blocks_to_dense(br, bc, bl, W)
    r_ref, g_ref, H_ref = float(G[tag + "_residual_sum"]), G[tag + "_g"], G[tag + "_H"]
    assert abs(r - r_ref) <= 1e-8 * abs(r_ref), (tag, r, r_ref)
    assert np.abs(g - g_ref).max() <= 1e-7 * np.abs(g_ref).max(), tag
    assert np.abs(H - H_ref).max() <= 1e-7 * np.abs(H_ref).max(), tag
    if tag != "L3":
        r_gt = P.residual(G[tag + "_poses_gt"])
        assert abs(r_gt - float(G[tag + "_residual_gt"])) <= 1e-8 * abs(float(G[tag + "_residual_gt"])), tag
    P.close()
    if tag != "L3":
        poses, s = pkg.lidar_lm(vp, pi, cl, ps)
        assert np.abs(poses - G[tag + "_lm_poses"]).max() <= 1e-6, (tag, np.abs(poses - G[tag + "_lm_poses"]).max())
        V = len(vp) - 1
        assert abs(s["cost_last"] * V - float(G[tag + "_lm_residual_sum"])) <= 1e-6 * float(G[tag + "_lm_residual_sum"]), tag
        assert abs(s["cost_first"] - float(G[tag + "_residual_avg_threads"])) <= 1e-8 * float(G[tag + "_residual_avg_threads"]), tag
        launches += s["kernel_launches"]

# ---- B3: cut_voxel -> recut -> tras_opt, findCorrespondPoint, damping_iter on the map
scans = split(G["M_xyz"], G["M_scan_ptr"])
m = pkg.VoxelMap(scans, G["M_poses"], float(G["M_voxel_size"]), G["M_eigen_ratio"])
got = m.export()
assert np.array_equal(got["vox_ptr"], G["M_vox_ptr"]) and np.array_equal(got["pose_idx"], G["M_pose_idx"])
assert np.array_equal(got["key"], G["M_key"]) and np.array_equal(got["path"][:, 0], G["M_layer"])
assert np.array_equal(got["path"][:, 1:].astype(np.int32), G["M_path"])
assert np.array_equal(got["clusters"][:, 9], G["M_clusters"][:, 9])
assert np.abs(got["clusters"] - G["M_clusters"]).max() <= 1e-12 * np.abs(G["M_clusters"]).max()
assert np.abs(got["centre"] - G["M_centre"]).max() <= 1e-9 and np.abs(got["eigenvalues"] - G["M_eigenvalues"]).max() <= 1e-9
assert np.all(np.abs(np.einsum("ij,ij->i", got["normal"], G["M_direct"])) >= 1 - 1e-6)         # eigenvector sign is the library's
st, d, c = G["M_lookup_state"], G["M_lookup_direct"], G["M_lookup_centre"]
nd_ref = np.zeros((len(st), 4))
for i in np.nonzero(st == 2)[0]:                              # the (n, d) step of recompute_local_planes (lvba_system.cpp:1552-1563)
    if np.all(np.isfinite(d[i])) and np.linalg.norm(d[i]) >= 1e-6 and np.all(np.isfinite(c[i])):
        n = d[i] / np.linalg.norm(d[i]); nd_ref[i, :3] = n; nd_ref[i, 3] = -n @ c[i]
nd = m.lookup(G["M_lookup_X"])
assert np.array_equal(np.all(nd == 0, axis=1), np.all(nd_ref == 0, axis=1))
sgn = np.sign(np.einsum("ij,ij->i", nd[:, :3], nd_ref[:, :3])); sgn[sgn == 0] = 1
assert np.abs(nd * sgn[:, None] - nd_ref).max() <= 1e-6
poses, s = m.lidar_lm(G["M_poses"])
assert np.abs(poses - G["M_lm_poses"]).max() <= 1e-6, np.abs(poses - G["M_lm_poses"]).max()
launches += s["kernel_launches"]
m.close()

# ---- B6: pl_transform + down_sampling_voxel2 (the reference's order is its unordered_map's: compared as sets)
cloud = pkg.anchor_clouds(scans, G["A_rel"], np.array([0, len(scans)], np.int32), float(G["A_leaf"]))[0]
assert np.array_equal(lex(cloud), G["A_cloud_sorted"])

# ---- B2: the cost Ceres evaluates = 1/2 (sum of ReprojErrorWhitenedDistorted^2 + sum of PointPlaneErrorWhitened^2)
n = G["V_plane_nd"][:, :3]
tv = np.isfinite(G["V_plane_nd"]).all(1) & (np.abs(n) > 1e-6).any(1)      # has_valid_plane, lvba_system.cpp:1598
trk = np.repeat(np.arange(len(G["V_obs_ptr"]) - 1), np.diff(G["V_obs_ptr"]))
cost_ref = 0.5 * (float((G["V_reproj_r"][tv[trk]] ** 2).sum()) + float((G["V_plane_r"][tv] ** 2).sum()))
P = pkg.VisualProblem(G["V_q"], G["V_t"], G["V_X"], G["V_plane_nd"], G["V_obs_ptr"], G["V_obs_cam"], G["V_obs_uv"], G["V_intr"],
                      float(G["V_sigma_px"]), float(G["V_sigma_plane"]))
cost = P.cost()
P.close()
assert abs(cost - cost_ref) <= 1e-10 * cost_ref, (cost, cost_ref)
assert launches > 0
print('CHILD-OK')
""" % (str(ROOT), str(ROOT / "tests" / "golden" / "ref_balm.npz"))


@pytest.mark.gpu
def test_device_reproduces_what_the_reference_source_computed():
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


CODE_SYSTEM = """
import sys
import numpy as np
sys.path.insert(0, %r)
import __graft_entry__ as graft
pkg = graft.load_package(); pkg.load_library()
assert pkg.device_count() >= 1
G = np.load(%r)
split = lambda flat, ptr: [flat[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
lex = lambda a: a[np.lexsort(a.T[::-1])]
# ---- B4: buildGridMapFromOptimized + generateDepthWithVoxel of src/lvba_system.cpp, bit for bit (grid voxel 0.5 m, window +-0.5 s: :1279, :1300)
g = pkg.DepthGrid(split(G["D_xyz"], G["D_scan_ptr"]), G["D_poses"], G["D_frame_ts"], 0.5)
img, info = g.render(G["D_cams"], G["D_image_ts"], G["D_intr"], int(G["D_size"][0]), int(G["D_size"][1]), 0.5)
g.close()
assert np.array_equal(img, G["D_depth"]) and info["kernel_launches"] > 0
# ---- B6 inside runWindowBA: the anchor clouds of the two windows that were solved (use_window_ba_rel = false: relative poses from the odometry)
scans = split(G["L_xyz"], G["L_scan_ptr"])
keep = [0, 1, 2, 6, 7, 8]
clouds = pkg.anchor_clouds([scans[i] for i in keep], G["L_rel_poses"][keep], np.array([0, 3, 6], np.int32), float(G["L_anchor_leaf"]))
for c, r in zip(clouds, split(G["L_anchor_clouds_sorted"], G["L_anchor_cloud_ptr"])):
    assert np.array_equal(lex(c), r)
# ---- N1: the window stage of runWindowBA (:232-279) — one voxel map and one damping_iter per window, batched on the device, the
# 3-voxels-per-pose skip rule — against the window solves the reference's own code produced (use_window_ba_rel: rel = anchor^-1 * aligned pose)
W, win = len(scans), int(G["L_window"])
win_ptr = np.array(list(range(0, W, win)) + [W], np.int32)
m = pkg.VoxelMap(scans, G["L_poses"], float(G["L_s1_voxel"]), (0.3, 0.1, 0.06, 0.03), win_ptr=win_ptr)     # bavoxel.hpp:17's ratios are in force here
solved, sums, tot = m.lidar_lm_batch(G["L_poses"], min_voxels_per_pose=3)
m.close()
rel = np.zeros((W, 12)); rel[:, [0, 4, 8]] = 1.0
for w in range(len(win_ptr) - 1):
    a, b = int(win_ptr[w]), int(win_ptr[w + 1])
    if np.array_equal(solved[a:b], G["L_poses"][a:b]):
        continue                                              # a skipped window keeps its poses and gets no anchor (:259-263)
    Ro, po = G["L_poses"][a, :9].reshape(3, 3), G["L_poses"][a, 9:]
    R_align = Ro @ solved[a, :9].reshape(3, 3).T; p_align = po - R_align @ solved[a, 9:]              # :267-279
    for j in range(a, b):
        Rj = R_align @ solved[j, :9].reshape(3, 3); pj = R_align @ solved[j, 9:] + p_align
        rel[j, :9] = (Ro.T @ Rj).ravel(); rel[j, 9:] = Ro.T @ (pj - po)                                # :286-289
assert np.array_equal(rel[3:6], G["L_rel_poses_rel"][3:6])     # the middle window was skipped on the device too
assert np.abs(rel - G["L_rel_poses_rel"]).max() <= 1e-6, np.abs(rel - G["L_rel_poses_rel"]).max()
assert tot["kernel_launches"] > 0
print('CHILD-OK')
""" % (str(ROOT), str(ROOT / "tests" / "golden" / "ref_system.npz"))


@pytest.mark.gpu
def test_device_reproduces_what_the_reference_pipeline_source_computed():
    """tests/golden/ref_system.npz (src/lvba_system.cpp compiled where it lies; CPU twin: tests/test_ref_system_pin.py): depth images and
    anchor clouds bit for bit.  The track fusion is compared on the CPU only: the reference's container order differs from the ABI's
    documented ascending order (tests/test_ref_system_pin.py::test_ascending_order_is_a_different_but_documented_choice)."""
    r = subprocess.run([sys.executable, "-c", CODE_SYSTEM], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


CODE_FUSE = """
import sys
import numpy as np
sys.path.insert(0, %r)
import __graft_entry__ as graft
from oracle import depth_oracle as dep
pkg = graft.load_package(); pkg.load_library()
assert pkg.device_count() >= 1
G = np.load(%r)
kp_ptr, kp_uv = G["F_kp_ptr"], G["F_kp_uv"]
Xw, valid = dep.backproject(list(G["F_depth"]), G["F_cams"], G["F_intr"], kp_ptr, kp_uv)      # equals the reference's depth candidates bit for bit (CPU test)
t = pkg.tracks_fuse(kp_ptr, kp_uv, G["F_matches"], G["F_cams"], G["F_intr"], Xw, valid, map_order=pkg.FUSE_ORDER_LIBSTDCXX)
assert t["summary"]["n_tracks"] == len(G["F_Xw"]) == 70 and t["summary"]["kernel_launches"] > 0
assert np.array_equal(t["obs_ptr"], G["F_obs_ptr"])
assert np.array_equal(np.column_stack([t["img"], t["kp"]]), G["F_obs"])                        # same tracks, same order, same member order
inl = np.zeros(len(t["img"]), bool)
for k in range(70):
    inl[G["F_obs_ptr"][k] + G["F_inl"][G["F_inl_ptr"][k]:G["F_inl_ptr"][k + 1]]] = True
assert np.array_equal(t["inlier"].astype(bool), inl)
assert np.abs(t["Xw"] - G["F_Xw"]).max() <= 1e-9
a = pkg.tracks_fuse(kp_ptr, kp_uv, G["F_matches"], G["F_cams"], G["F_intr"], Xw, valid, map_order=pkg.FUSE_ORDER_ASCENDING)   # the library-independent order: another answer here
assert a["summary"]["n_tracks"] == 63
print('CHILD-OK')
""" % (str(ROOT), str(ROOT / "tests" / "golden" / "ref_system.npz"))


@pytest.mark.gpu
def test_device_track_fusion_reproduces_the_reference_source_under_its_container_order():
    """B7 with lvba_fuse_opts::map_order = LVBA_FUSE_ORDER_LIBSTDCXX against the reference's own BuildTracksAndFuse3D (tests/golden/ref_system.npz):
    track order, member order, inlier sets, fused points.  CPU twin through the host policy:
    tests/test_ref_system_pin.py::test_device_fusion_pass_equals_reference_source_under_its_container_order."""
    r = subprocess.run([sys.executable, "-c", CODE_FUSE], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.gpu
def test_offline_camera_half_reproduces_the_reference_pipeline_source(tmp_path):
    """`lvba_offline --visual --fuse-order libstdcxx` (the whole camera half on the device: depth candidates -> track fusion -> anchors -> surf map ->
    planes -> visual LM) against the reference's own runVisualBAWithLidarAssist chain on the same seeded dataset (tests/golden/ref_system.npz, section V:
    src/lvba_system.cpp compiled where it lies, its own COLMAP reader, the restated Ceres loop in the place of ceres::Solve): the same number of fused
    tracks, the same points entering the problem, the same initial cost, cameras equal to what images.txt's six decimals resolve."""
    import json
    import numpy as np
    import __graft_entry__ as graft
    sys.path.insert(0, str(ROOT / "tests"))
    import visual_scene
    from oracle import dataset_writer as dw
    G = np.load(ROOT / "tests" / "golden" / "ref_system.npz")
    pkg = graft.load_package()
    exe = tmp_path / "lvba_offline"
    cmd = ["g++", "-std=c++17", "-O2", "-I", str(ROOT / "include"), str(ROOT / "tools" / "lvba_offline.cpp"), "-o", str(exe),
           str(pkg.LIB_PATH), f"-Wl,-rpath,{pkg.LIB_PATH.parent}", "-L/usr/local/cuda/lib64", "-lcudart", "-ldl"]
    assert subprocess.run(cmd, capture_output=True, text=True).returncode == 0
    data = tmp_path / "data"
    visual_scene.make(data, seed=3, W=8, n_landmarks=700)
    r = subprocess.run([str(exe), "--data", str(data), "--config", str(data / "config.yaml"), "--visual", "--fuse-order", "libstdcxx"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    vis = [x for x in (json.loads(ln) for ln in r.stdout.strip().splitlines()) if x.get("stage") == "visual"][0]
    assert vis["tracks"] == int(G["V_counts"][0]) and vis["points_kept"] == int(G["V_counts"][1])
    assert abs(vis["cost_first"] - G["V_costs"][0]) <= 1e-6 * G["V_costs"][0]
    assert abs(vis["cost_last"] - G["V_costs"][1]) <= 1e-3 * G["V_costs"][1]
    rows = [ln.split() for ln in (data / "Colmap" / "sparse" / "images.txt").read_text().splitlines()]
    after = np.zeros((len(rows) // 2, 12))
    for k in range(len(rows) // 2):
        q = np.array([float(v) for v in rows[2 * k][1:5]]); t = np.array([float(v) for v in rows[2 * k][5:8]])
        after[k, :9] = dw.quat_to_R(q).ravel(); after[k, 9:] = t
    assert after.shape == G["V_cams_after"].shape
    assert np.abs(after - G["V_cams_after"]).max() <= 1e-4, np.abs(after - G["V_cams_after"]).max()      # the cameras moved by 8e-2 from their start
    assert np.abs(G["V_cams_after"] - G["V_cams_before"]).max() > 1e-2
