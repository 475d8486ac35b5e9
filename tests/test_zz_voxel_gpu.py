"""Boundary B3 on a B200: the device-built adaptive voxel map (lvba_voxel_map_*, global-lvba_b200/csrc/voxel_api.cuh)
against oracle/voxel_oracle.py through the C ABI — the comparisons of tests/test_voxel_emu.py, plus the chain
scans -> voxel map -> lvba_lidar_lm against the oracle's own chain.

Every case runs in a CHILD process under a timeout: this path had its first hardware run after the rest of the suite,
so a device fault here cannot poison the CUDA context of the other GPU tests (the file also sorts last)."""
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

# First hardware run (B200, profiles/r01_voxel_gpu_first_run.txt): every case below passed except the windowed one, whose
# test body shadowed the `lo` oracle module at the time (fixed since, not re-run: the round's GPU budget was spent) while the
# C++ mirror of the same stage (run_window_stage, last case) passed.  That one case stays non-gating until it is re-run.

PRELUDE = """
import sys
import numpy as np
sys.path.insert(0, %r)
import __graft_entry__ as graft
from oracle import synth, voxel_oracle as vox, lidar_oracle as lo
pkg = graft.load_package(); pkg.load_library()
assert pkg.device_count() >= 1

def compare_with_oracle(got, ref):
    vp, pi, cl, meta = ref
    assert np.array_equal(got["vox_ptr"], vp), (len(got["vox_ptr"]), len(vp))
    assert np.array_equal(got["pose_idx"], pi)
    assert np.array_equal(got["key"], meta["key"])
    assert np.array_equal(got["path"][:, 0], meta["layer"])
    for v, p in enumerate(meta["path"]):
        assert got["path"][v, 1:].tolist() == list(p) + [-1] * (2 - len(p))
    assert np.array_equal(got["clusters"][:, 9], cl[:, 9])
    scale = max(1.0, np.abs(cl).max()) if len(cl) else 1.0
    assert np.abs(got["clusters"] - cl).max(initial=0.0) <= 1e-12 * scale
    assert np.abs(got["centre"] - meta["centre"]).max(initial=0.0) <= 1e-9
    assert np.abs(got["eigenvalues"] - meta["eigenvalues"]).max(initial=0.0) <= 1e-9
    lam = meta["eigenvalues"]
    distinct = (lam[:, 1] - lam[:, 0]) > 1e-6 * np.maximum(lam[:, 2], 1e-300)
    dots = np.abs(np.einsum("ij,ij->i", got["normal"], meta["direct"]))
    assert np.all(dots[distinct] >= 1 - 1e-6)

def compare_lookup(got, ref):
    assert np.array_equal(np.all(got == 0, axis=1), np.all(ref == 0, axis=1))
    sgn = np.sign(np.einsum("ij,ij->i", got[:, :3], ref[:, :3])); sgn[sgn == 0] = 1
    assert np.abs(got * sgn[:, None] - ref).max(initial=0.0) <= 1e-6
""" % str(ROOT)


def _run(body, timeout=300):
    code = PRELUDE + textwrap.dedent(body) + "\nprint('CHILD-OK')\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))
    assert r.returncode == 0 and "CHILD-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    return r.stdout


@pytest.mark.gpu
def test_voxel_map_equals_oracle():
    _run("""
    for seed, vs, ll in [(1, 1.0, 2), (2, 0.5, 2), (3, 2.0, 2), (4, 1.0, 1), (5, 1.0, 0), (6, 4.0, 2)]:
        scans, poses = synth.make_scan_scene(seed)
        m = pkg.VoxelMap(scans, poses, vs, layer_limit=ll)
        got = m.export()
        compare_with_oracle(got, vox.voxelize(scans, poses, vs, layer_limit=ll))
        assert m.summary["n_voxels"] == len(got["vox_ptr"]) - 1 > 0 and m.summary["kernel_launches"] > 0
        m.close()
    """)


@pytest.mark.gpu
def test_voxel_map_equals_golden_fixture():
    """Committed known-answer vectors (tests/golden/voxel_scene.npz): planes at every octree layer + plane lookups."""
    _run("""
    g = np.load("tests/golden/voxel_scene.npz")
    sp = g["scan_ptr"]
    for tag, vs in zip("ab", g["voxel_sizes"]):
        path = g[f"path_{tag}"]
        meta = dict(key=g[f"key_{tag}"], layer=path[:, 0].astype(np.int32), path=[tuple(int(x) for x in r[1:1 + r[0]]) for r in path],
                    centre=g[f"centre_{tag}"], direct=g[f"direct_{tag}"], eigenvalues=g[f"eigenvalues_{tag}"])
        m = pkg.VoxelMap(g["xyz"], g["poses"], float(vs), scan_ptr=sp)
        compare_with_oracle(m.export(), (g[f"vox_ptr_{tag}"], g[f"pose_idx_{tag}"], g[f"clusters_{tag}"], meta))
        compare_lookup(m.lookup(g["X"]), g[f"plane_nd_{tag}"])
        m.close()
    """)


@pytest.mark.gpu
def test_voxel_map_edge_cases():
    _run("""
    scans, poses = synth.make_scan_scene(9, W=5, n_per_scan=1500)
    ragged = [scans[0], np.zeros((0, 3), np.float32), scans[2], scans[3][:1], np.zeros((0, 3), np.float32)]
    m = pkg.VoxelMap(ragged, poses)
    compare_with_oracle(m.export(), vox.voxelize(ragged, poses)); m.close()
    m = pkg.VoxelMap([np.zeros((0, 3), np.float32)] * 3, poses[:3])
    assert m.export()["vox_ptr"].tolist() == [0] and m.lookup(np.zeros((2, 3))).tolist() == [[0] * 4] * 2; m.close()
    m = pkg.VoxelMap(scans, poses, eigen_ratio=(1e-9,) * 4)
    assert m.summary["n_voxels"] == 0; m.close()
    for min_ps in (1, 40, 10 ** 6):
        m = pkg.VoxelMap(scans, poses, min_points=min_ps)
        compare_with_oracle(m.export(), vox.voxelize(scans, poses, min_ps=min_ps)); m.close()
    shift = poses.copy(); shift[:, 9:] += np.array([-37.25, 12.5, -3.0])
    m = pkg.VoxelMap(scans, shift, 0.5)
    compare_with_oracle(m.export(), vox.voxelize(scans, shift, 0.5)); m.close()
    bad = [s.copy() for s in scans]; bad[2][7, 1] = np.nan
    try:
        pkg.VoxelMap(bad, poses); raise SystemExit("NaN point accepted")
    except pkg.LvbaError as e:
        assert e.status == -1
    # PCL-style records: 12 floats per point, xyz in front
    N = sum(len(s) for s in scans)
    rec = np.full((N, 12), 7.0, np.float32); rec[:, :3] = np.concatenate(scans)
    sp = np.concatenate([[0], np.cumsum([len(s) for s in scans])])
    m = pkg.VoxelMap(rec, poses, scan_ptr=sp)
    compare_with_oracle(m.export(), vox.voxelize(scans, poses)); m.close()
    """)


@pytest.mark.gpu
def test_plane_lookup_equals_oracle():
    _run("""
    scans, poses = synth.make_scan_scene(8, W=4, n_per_scan=2000)
    rng = np.random.default_rng(0)
    X = np.concatenate([rng.uniform(-3.5, 3.5, (3000, 3)),
                        np.column_stack([rng.uniform(-3, 3, (1500, 2)), np.full(1500, -1.2)]),
                        np.array([[np.nan, 0, 0], [0, np.inf, 0], [1e6, 1e6, 1e6], [-1e6, 0, 0]])])
    for vs, ll in [(1.0, 2), (2.0, 2), (2.0, 1), (1.0, 0)]:
        m = pkg.VoxelMap(scans, poses, vs, layer_limit=ll)
        roots = vox.build_tree_literal(scans, poses, vs, layer_limit=ll)
        compare_lookup(m.lookup(X), vox.plane_lookup_literal(roots, X, vs, ll)); m.close()
    """)


@pytest.mark.gpu
def test_scans_to_lm_chain():
    """raw scans -> device voxel map -> lvba_lidar_lm, against the same chain through the oracles."""
    _run("""
    scans, poses = synth.make_scan_scene(21, W=6, n_per_scan=4000)
    rng = np.random.default_rng(3)
    noisy = poses.copy()
    for i in range(1, len(noisy)):
        noisy[i, :9] = (noisy[i, :9].reshape(3, 3) @ synth.so3_exp(rng.normal(0, 0.01, (1, 3)))[0]).ravel()
        noisy[i, 9:] += rng.normal(0, 0.02, 3)
    m = pkg.VoxelMap(scans, noisy)
    g = m.export(); m.close()
    vp, pi, cl, _ = vox.voxelize(scans, noisy)
    assert np.array_equal(g["vox_ptr"], vp) and np.array_equal(g["pose_idx"], pi)
    out_gpu, s = pkg.lidar_lm(g["vox_ptr"], g["pose_idx"], g["clusters"], noisy)
    out_ref, info = lo.damping_iter(vp, pi, cl, noisy)
    assert s["cost_last"] < s["cost_first"]
    assert abs(s["cost_last"] - info["r_last"]) <= 1e-6 * info["r_last"]
    assert np.abs(out_gpu - out_ref).max() <= 1e-6
    # the same solve with the clusters never leaving the device (lvba_voxel_map_lidar_lm): identical input, identical result
    m = pkg.VoxelMap(scans, noisy)
    out_dev, s2 = m.lidar_lm(noisy)
    assert np.abs(out_dev - out_gpu).max() <= 1e-9 and s2["iterations"] == s["iterations"]   # (atomic sums: not bitwise)
    assert s2["h2d_bytes"] < s["h2d_bytes"] - 70 * len(g["pose_idx"])          # the 80 B records were not uploaded
    P = m.lidar_problem(noisy)
    assert abs(P.build() - s["cost_first"] * (len(vp) - 1)) <= 1e-9 * abs(P.build())
    P.close()
    out_skip, s3 = m.lidar_lm(noisy, min_voxels_per_pose=10 ** 6)
    assert s3["termination"] == 6 and np.array_equal(out_skip, noisy)          # LVBA_TERM_SKIPPED
    m.close()
    """)


@pytest.mark.gpu
def test_windowed_map_and_batched_lm():
    """runWindowBA's window stage: one map per window built together, then every window solved in one batched LM —
    against one oracle map + one oracle damping_iter per window."""
    _run("""
    sizes = [5, 4, 1, 6]
    scans, poses = synth.make_scan_scene(13, W=sum(sizes), n_per_scan=3000)
    rng = np.random.default_rng(4)
    noisy = poses.copy()
    for i in range(len(noisy)):
        noisy[i, :9] = (noisy[i, :9].reshape(3, 3) @ synth.so3_exp(rng.normal(0, 0.004, (1, 3)))[0]).ravel()
        noisy[i, 9:] += rng.normal(0, 0.01, 3)
    win_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    m = pkg.VoxelMap(scans, noisy, 1.0, win_ptr=win_ptr)
    g = m.export(); wins = m.windows()
    a = 0; off = 0; refs = []
    for w, n in enumerate(sizes):
        vp, pi, cl, meta = vox.voxelize(scans[a:a + n], noisy[a:a + n], 1.0)
        sel = wins == w
        assert sel.sum() == len(vp) - 1 and np.array_equal(g["key"][sel], meta["key"])
        assert np.array_equal(g["pose_idx"][off:off + len(pi)], pi + a)
        assert np.abs(g["clusters"][off:off + len(pi)] - cl).max(initial=0.0) <= 1e-12 * max(1.0, np.abs(cl).max(initial=0.0))
        refs.append((vp, pi, cl)); a += n; off += len(pi)
    out, sums, tot = m.lidar_lm_batch(noisy, min_voxels_per_pose=3)
    m.close()
    a = 0
    for w, n in enumerate(sizes):
        vp, pi, cl = refs[w]
        if len(vp) - 1 < 3 * n:
            assert sums[w]["termination"] == 6 and np.array_equal(out[a:a + n], noisy[a:a + n])      # skipped (:262-266)
        else:
            ref_poses, info = lo.damping_iter(vp, pi, cl, noisy[a:a + n])
            assert abs(sums[w]["cost_last"] - info["r_last"]) <= 1e-6 * info["r_last"]
            assert np.abs(out[a:a + n] - ref_poses).max() <= 1e-6
        a += n
    assert any(s_["termination"] == 6 for s_ in sums) and any(s_["termination"] != 6 for s_ in sums)
    # the same stage through the host-array entry point gives the same poses
    out2, _, _ = pkg.lidar_lm_batch(win_ptr, g["vox_ptr"], g["pose_idx"], g["clusters"], noisy, 3)
    assert np.abs(out - out2).max() <= 1e-9
    """)


@pytest.mark.gpu
def test_large_map_invariants():
    """A map far beyond what the oracle can follow (2 M points): size-independent properties."""
    _run("""
    scans, poses = synth.make_scan_scene(31, W=40, n_per_scan=50000)
    m = pkg.VoxelMap(scans, poses, 0.5)
    g = m.export()
    V = m.summary["n_voxels"]
    assert V > 100 and m.summary["n_points"] == 40 * 50000
    vp, pi, cl = g["vox_ptr"], g["pose_idx"], g["clusters"]
    assert np.all(np.diff(vp) >= 2)                                     # push_voxel: seen from >= 2 poses
    starts = np.zeros(len(pi), bool); starts[vp[:-1]] = True
    assert np.all((np.diff(pi) > 0) | starts[1:])                       # ascending pose index inside a voxel
    npts = np.add.reduceat(cl[:, 9], vp[:-1])
    assert np.all(npts >= 15)                                           # min_ps
    lam = g["eigenvalues"]; lim = np.array([0.3, 0.1, 0.06, 0.03], np.float32).astype(np.float64)
    assert np.all(lam[:, 0] / lam[:, 2] <= lim[g["path"][:, 0]])        # plane test of the voxel's own layer
    k = g["key"]; order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    assert np.array_equal(order, np.arange(V)) or np.all(np.diff(k[:, 0]) >= 0)
    for i in range(40):                                                 # a point is in at most one voxel
        assert cl[pi == i, 9].sum() <= 50000
    # idempotence: the same input gives the same map, bit for bit
    m2 = pkg.VoxelMap(scans, poses, 0.5); g2 = m2.export(); m2.close()
    assert all(np.array_equal(g[x], g2[x]) for x in g)
    # every voxel's own centre looks itself up
    nd = m.lookup(g["centre"])
    assert np.count_nonzero(np.any(nd != 0, axis=1)) >= 0.9 * V
    m.close()
    """, timeout=600)


@pytest.mark.gpu
def test_shim_surf_map(tmp_path):
    """The C++ mirror of the reference call sites (lvba_b200::SurfMap, host/lvba_shim.hpp): scans -> map -> LM -> planes."""
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as graft
    pkg = graft.load_package()
    exe = tmp_path / "test_shim"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "shim" / "test_shim.cpp"),
           "-o", str(exe), str(pkg.LIB_PATH), f"-Wl,-rpath,{pkg.LIB_PATH.parent}", "-L/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), "surfmap"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "surf map ok" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([str(exe), "windows"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "window stage ok" in r.stdout, r.stdout + r.stderr
