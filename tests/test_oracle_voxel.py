"""Pins the set-up-stage oracle (oracle/voxel_oracle.py — cut_voxel / recut / tras_opt of the reference,
include/BALM/bavoxel.hpp:335-474, 799-836) without a GPU: the vectorised restatement against the literal recursive
tree, structural invariants of the result, and the float / negative-coordinate quirks of the root key."""
import numpy as np
import pytest

from oracle import lidar_oracle as lo
from oracle import synth
from oracle import voxel_oracle as vox


_scene = synth.make_scan_scene


@pytest.mark.parametrize("seed,voxel_size", [(1, 1.0), (2, 0.5), (3, 2.0)])
def test_vectorised_equals_literal(seed, voxel_size):
    scans, poses = _scene(seed)
    a = vox.voxelize(scans, poses, voxel_size)
    b = vox.voxelize_literal(scans, poses, voxel_size)
    assert len(a[0]) > 3                                              # the scene does produce plane voxels
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[3]["key"], b[3]["key"]) and a[3]["path"] == b[3]["path"]
    assert np.array_equal(a[3]["layer"], b[3]["layer"])
    assert np.abs(a[2] - b[2]).max() <= 1e-9 * max(1.0, np.abs(b[2]).max())
    if voxel_size >= 1.0:
        assert len(set(a[3]["layer"])) > 1                            # more than one octree layer is exercised


def test_invariants_and_feed_into_lidar_oracle():
    scans, poses = _scene(7, W=6)
    vp, pi, cl, meta = vox.voxelize(scans, poses, 1.0)
    V = len(vp) - 1
    for a in range(V):
        sl = slice(vp[a], vp[a + 1])
        assert vp[a + 1] - vp[a] >= 2                                 # push_voxel: seen from >= 2 poses
        assert np.all(np.diff(pi[sl]) > 0)                            # ascending pose index (ABI of lvba_lidar_lm)
        assert cl[sl, 9].sum() >= vox.MIN_PS                          # min_ps
        lam = meta["eigenvalues"][a]
        assert lam[0] / lam[2] <= vox.EIGEN_RATIO_DEFAULT[meta["layer"][a]] + 1e-12
        assert len(meta["path"][a]) == meta["layer"][a]
    # every point ends in at most one voxel: cluster counts never exceed the number of points of that pose
    for i in range(len(scans)):
        assert cl[pi == i, 9].sum() <= len(scans[i])
    # the result is a valid input of path A
    r, g, _ = lo.acc_evaluate2(vp, pi, cl, poses, len(scans))
    assert np.isfinite(r) and np.all(np.isfinite(g)) and r > 0


def test_root_key_quirks():
    """:812-816 — the division is rounded to float before the sign test, negatives are shifted by one BEFORE truncation,
    so an exactly negative integer coordinate lands one voxel lower than floor() would put it."""
    w = np.array([[0.3, -0.3, -2.0], [1.99999999, -1.0000001, 2.0], [-0.0, 5.5, -7.25]])
    k = vox.root_keys(w, 1.0)
    assert k.tolist() == [[0, -1, -3], [2, -2, 2], [0, 5, -8]]        # 1.99999999 rounds to 2.0f ; -2.0 -> -3 ; -0.0 is not < 0
    assert vox.root_keys(np.array([[0.74, -0.74, 0.76]]), 0.5).tolist() == [[1, -2, 1]]


def test_threshold_and_layer_limit():
    scans, poses = _scene(11)
    strict = vox.voxelize(scans, poses, 1.0, eigen_ratio=(1e-9, 1e-9, 1e-9, 1e-9))
    assert len(strict[0]) - 1 == 0                                    # nothing passes an impossible plane test
    coarse = vox.voxelize(scans, poses, 1.0, layer_limit=0)
    full = vox.voxelize(scans, poses, 1.0, layer_limit=2)
    assert np.all(coarse[3]["layer"] == 0) and len(full[0]) >= len(coarse[0])


def test_golden_fixture_is_reproduced():
    """tests/golden/voxel_scene.npz (literal restatement, committed) against the vectorised restatement run now."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "voxel_scene.npz")
    sp = g["scan_ptr"]
    scans = [g["xyz"][sp[j]:sp[j + 1]] for j in range(len(sp) - 1)]
    layers = set()
    for tag, vs in zip("ab", g["voxel_sizes"]):
        vp, pi, cl, meta = vox.voxelize(scans, g["poses"], float(vs))
        assert np.array_equal(vp, g[f"vox_ptr_{tag}"]) and np.array_equal(pi, g[f"pose_idx_{tag}"])
        assert np.array_equal(meta["key"], g[f"key_{tag}"]) and np.array_equal(meta["layer"], g[f"path_{tag}"][:, 0])
        assert np.abs(cl - g[f"clusters_{tag}"]).max() <= 1e-9 * np.abs(cl).max()
        roots = vox.build_tree_literal(scans, g["poses"], float(vs))
        assert np.array_equal(vox.plane_lookup_literal(roots, g["X"], float(vs)), g[f"plane_nd_{tag}"])
        layers |= set(meta["layer"].tolist())
    assert layers == {0, 1, 2}
