"""The device against the committed known-answer vectors of the set-up stages (tests/golden/setup_stages.npz; CPU twin:
tests/test_golden_setup.py): B4 depth images and B6 anchor clouds bit for bit, B7 fused tracks observation for observation
(points / mean reprojection errors to 1e-9), all through the C ABI.  Nothing is regenerated on the GPU box.  Runs in a child
process under a timeout like the other test_zz_* files."""
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
# ---- B4
g = pkg.DepthGrid(split(G["d_xyz"], G["d_scan_ptr"]), G["d_poses"], G["d_frame_ts"], float(G["d_voxel_size"]))
img, info = g.render(G["d_cams"], G["d_image_ts"], G["d_intr"], int(G["d_size"][0]), int(G["d_size"][1]), float(G["d_half_window"]))
g.close()
assert np.array_equal(img, G["d_images"]) and info["kernel_launches"] > 0
# ---- B6
clouds = pkg.anchor_clouds(split(G["a_xyz"], G["a_scan_ptr"]), G["a_rel"], G["a_win_ptr"], float(G["a_leaf"]))
assert all(np.array_equal(c, r) for c, r in zip(clouds, split(G["a_cloud_xyz"], G["a_cloud_ptr"])))
# ---- B7
t = pkg.tracks_fuse(G["f_kp_ptr"], G["f_kp_uv"], G["f_matches"], G["f_cams"], G["f_intr"], G["f_kp_Xw"], G["f_kp_valid"],
                    map_order=pkg.FUSE_ORDER_ASCENDING)       # the file holds the ascending-order answer (the container order: tests/test_zzz_ref_gpu.py)
assert np.array_equal(t["obs_ptr"], G["f_obs_ptr"]) and np.array_equal(t["img"], G["f_obs_img"]) and np.array_equal(t["kp"], G["f_obs_kp"])
assert np.array_equal(t["inlier"].astype(np.uint8), G["f_inlier"]) and np.array_equal(np.asarray(t["source"], np.uint8), G["f_source"])
assert np.abs(t["Xw"] - G["f_Xw"]).max() <= 1e-9 * max(1.0, np.abs(G["f_Xw"]).max())
assert np.abs(t["mean"] - G["f_mean"]).max() <= 1e-9 * max(1.0, np.abs(G["f_mean"]).max())
assert t["summary"]["n_tracks"] == len(G["f_source"]) == 24 and t["summary"]["kernel_launches"] > 0
print('CHILD-OK')
""" % (str(ROOT), str(ROOT / "tests" / "golden" / "setup_stages.npz"))


@pytest.mark.gpu
def test_device_reproduces_the_setup_stage_fixture():
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
