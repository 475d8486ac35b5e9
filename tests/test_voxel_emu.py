"""The voxel-map pipeline (global-lvba_b200/csrc/voxel_pipeline.h — boundary B3) checked WITHOUT a GPU:
the same pass functors the CUDA kernels run are instantiated with a sequential host policy (tests/emu/voxel_emu.cpp,
test infrastructure only) and compared with oracle/voxel_oracle.py — structure (voxels, poses, keys, paths) exactly,
cluster sums / centres / eigenvalues to rounding, normals up to sign.  The GPU tests (test_zz_voxel_gpu.py) run the
same comparisons through the C ABI."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import synth
from oracle import voxel_oracle as vox

_scene = synth.make_scan_scene

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "libvoxel_emu.so"
    cmd = ["g++", "-std=c++17", "-O2", *(["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []), "-ffp-contract=off", "-Wall", "-fPIC", "-shared",
           str(ROOT / "tests" / "emu" / "voxel_emu.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(str(so))


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class EmuMap:
    def __init__(self, lib, scans, poses, voxel_size=1.0, eigen_ratio=vox.EIGEN_RATIO_DEFAULT, layer_limit=2, min_ps=15):
        self.lib = lib
        W = len(scans)
        scan_ptr = np.zeros(W + 1, np.int64)
        scan_ptr[1:] = np.cumsum([len(s) for s in scans])
        xyz = (np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in scans]) if W else np.zeros((0, 3), np.float32))
        xyz = np.ascontiguousarray(xyz, np.float32)
        poses = np.ascontiguousarray(poses, np.float64)
        er = np.asarray(eigen_ratio, np.float32)
        self.h = ctypes.c_void_p()
        self.rc = lib.emu_voxel_map_create(ctypes.c_int32(W), _ptr(scan_ptr, ctypes.c_int64), _ptr(xyz, ctypes.c_float),
                                           _ptr(poses, ctypes.c_double), ctypes.c_double(voxel_size), _ptr(er, ctypes.c_float),
                                           ctypes.c_int32(layer_limit), ctypes.c_int32(min_ps), ctypes.byref(self.h))

    def export(self):
        V = ctypes.c_int64(); nnz = ctypes.c_int64(); nodes = np.zeros(3, np.int64)
        self.lib.emu_voxel_map_sizes(self.h, ctypes.byref(V), ctypes.byref(nnz), _ptr(nodes, ctypes.c_int64))
        V, nnz = V.value, nnz.value
        o = dict(vox_ptr=np.zeros(V + 1, np.int64), pose_idx=np.zeros(nnz, np.int32), clusters=np.zeros((nnz, 10)),
                 key=np.zeros((V, 3), np.int64), path=np.zeros((V, 3), np.int8), centre=np.zeros((V, 3)),
                 normal=np.zeros((V, 3)), eigenvalues=np.zeros((V, 3)), n_nodes=nodes)
        self.lib.emu_voxel_map_export(self.h, _ptr(o["vox_ptr"], ctypes.c_int64), _ptr(o["pose_idx"], ctypes.c_int32),
                                      _ptr(o["clusters"], ctypes.c_double), _ptr(o["key"], ctypes.c_int64),
                                      _ptr(o["path"], ctypes.c_int8), _ptr(o["centre"], ctypes.c_double),
                                      _ptr(o["normal"], ctypes.c_double), _ptr(o["eigenvalues"], ctypes.c_double))
        return o

    def lookup(self, X):
        X = np.ascontiguousarray(X, np.float64).reshape(-1, 3)
        out = np.zeros((len(X), 4))
        rc = self.lib.emu_voxel_map_lookup(self.h, ctypes.c_int64(len(X)), _ptr(X, ctypes.c_double), _ptr(out, ctypes.c_double))
        assert rc == 0
        return out

    def close(self):
        if self.h:
            self.lib.emu_voxel_map_destroy(self.h)
            self.h = ctypes.c_void_p()


def compare_with_oracle(got, ref):
    """got: dict from the pipeline; ref: (vox_ptr, pose_idx, clusters, meta) from oracle.voxel_oracle.voxelize*."""
    vp, pi, cl, meta = ref
    assert np.array_equal(got["vox_ptr"], vp)
    assert np.array_equal(got["pose_idx"], pi)
    assert np.array_equal(got["key"], meta["key"])
    assert np.array_equal(got["path"][:, 0], meta["layer"])
    for v, p in enumerate(meta["path"]):
        want = list(p) + [-1] * (2 - len(p))
        assert got["path"][v, 1:].tolist() == want
    assert np.array_equal(got["clusters"][:, 9], cl[:, 9])                                # point counts: exact
    scale = max(1.0, np.abs(cl).max()) if len(cl) else 1.0
    assert np.abs(got["clusters"] - cl).max(initial=0.0) <= 1e-12 * scale                 # sums: order of additions only
    assert np.abs(got["centre"] - meta["centre"]).max(initial=0.0) <= 1e-9
    assert np.abs(got["eigenvalues"] - meta["eigenvalues"]).max(initial=0.0) <= 1e-9
    lam = meta["eigenvalues"]
    distinct = (lam[:, 1] - lam[:, 0]) > 1e-6 * np.maximum(lam[:, 2], 1e-300)             # a repeated lambda0 has no unique vector
    dots = np.abs(np.einsum("ij,ij->i", got["normal"], meta["direct"]))
    assert np.all(dots[distinct] >= 1 - 1e-6)                                             # eigenvector sign is free


def compare_lookup(got, ref):
    """(n, d) rows up to a common sign; zero rows must agree exactly."""
    assert np.array_equal(np.all(got == 0, axis=1), np.all(ref == 0, axis=1))
    sgn = np.sign(np.einsum("ij,ij->i", got[:, :3], ref[:, :3]))
    sgn[sgn == 0] = 1
    assert np.abs(got * sgn[:, None] - ref).max(initial=0.0) <= 1e-6


@pytest.mark.parametrize("seed,voxel_size,layer_limit", [(1, 1.0, 2), (2, 0.5, 2), (3, 2.0, 2), (4, 1.0, 1), (5, 1.0, 0), (6, 4.0, 2)])
def test_pipeline_equals_oracle(emu, seed, voxel_size, layer_limit):
    scans, poses = _scene(seed)
    m = EmuMap(emu, scans, poses, voxel_size, layer_limit=layer_limit)
    assert m.rc == 0
    got = m.export()
    compare_with_oracle(got, vox.voxelize(scans, poses, voxel_size, layer_limit=layer_limit))
    compare_with_oracle(got, vox.voxelize_literal(scans, poses, voxel_size, layer_limit=layer_limit))
    assert len(got["vox_ptr"]) > 1
    m.close()


def test_plane_lookup_equals_oracle(emu):
    scans, poses = _scene(8, W=4, n_per_scan=2000)
    rng = np.random.default_rng(0)
    X = np.concatenate([rng.uniform(-3.5, 3.5, (3000, 3)),                                   # anywhere in and around the scene
                        np.column_stack([rng.uniform(-3, 3, (1500, 2)), np.full(1500, -1.2)]),   # on the floor
                        np.array([[np.nan, 0, 0], [0, np.inf, 0], [1e6, 1e6, 1e6], [-1e6, 0, 0]])])
    for voxel_size, layer_limit in [(1.0, 2), (2.0, 2), (2.0, 1), (1.0, 0)]:
        m = EmuMap(emu, scans, poses, voxel_size, layer_limit=layer_limit)
        roots = vox.build_tree_literal(scans, poses, voxel_size, layer_limit=layer_limit)
        ref = vox.plane_lookup_literal(roots, X, voxel_size, layer_limit)
        got = m.lookup(X)
        compare_lookup(got, ref)
        assert 0 < np.count_nonzero(np.any(got != 0, axis=1)) < len(X)
        m.close()


def test_empty_and_ragged_inputs(emu):
    scans, poses = _scene(9, W=5, n_per_scan=1500)
    # an empty scan in the middle, one at the end, and a one-point scan
    ragged = [scans[0], np.zeros((0, 3), np.float32), scans[2], scans[3][:1], np.zeros((0, 3), np.float32)]
    m = EmuMap(emu, ragged, poses)
    assert m.rc == 0
    compare_with_oracle(m.export(), vox.voxelize(ragged, poses))
    m.close()
    # no points at all
    none = [np.zeros((0, 3), np.float32)] * 3
    m = EmuMap(emu, none, poses[:3])
    assert m.rc == 0
    got = m.export()
    assert got["vox_ptr"].tolist() == [0] and m.lookup(np.zeros((2, 3))).tolist() == [[0] * 4] * 2
    m.close()
    # a single pose never yields a BA voxel (push_voxel needs two), but planes are still found by the lookup
    m = EmuMap(emu, scans[:1], poses[:1])
    assert m.export()["vox_ptr"].tolist() == [0]
    roots = vox.build_tree_literal(scans[:1], poses[:1])
    X = np.column_stack([np.linspace(-2, 2, 50), np.linspace(-2, 2, 50), np.full(50, -1.2)])
    got = m.lookup(X)
    compare_lookup(got, vox.plane_lookup_literal(roots, X))
    m.close()


def test_thresholds_min_points_and_bad_points(emu):
    scans, poses = _scene(10)
    m = EmuMap(emu, scans, poses, eigen_ratio=(1e-9,) * 4)
    assert m.export()["vox_ptr"].tolist() == [0]
    m.close()
    for min_ps in (1, 40, 10 ** 6):
        m = EmuMap(emu, scans, poses, min_ps=min_ps)
        compare_with_oracle(m.export(), vox.voxelize(scans, poses, min_ps=min_ps))
        m.close()
    bad = [s.copy() for s in scans]
    bad[2][7, 1] = np.nan
    assert EmuMap(emu, bad, poses).rc == -1                                # LVBA_ERR_INVALID_ARG
    far = [s.copy() for s in scans]
    far[0][0, 0] = 3e9
    assert EmuMap(emu, far, poses, voxel_size=1.0).rc == -1


def test_negative_coordinates_and_key_order(emu):
    """The root-key quirks (float rounding, `-= 1` before truncation) and the (x, y, z)-major voxel order."""
    scans, poses = _scene(12)
    shift = poses.copy()
    shift[:, 9:] += np.array([-37.25, 12.5, -3.0])                         # most keys negative on x and z
    m = EmuMap(emu, scans, shift, 0.5)
    got = m.export()
    compare_with_oracle(got, vox.voxelize(scans, shift, 0.5))
    k = got["key"]
    assert np.any(k[:, 0] < 0) and np.all(np.diff(k[:, 0]) >= 0)
    m.close()


def golden_case(tag):
    g = np.load(ROOT / "tests" / "golden" / "voxel_scene.npz")
    sp = g["scan_ptr"]
    scans = [g["xyz"][sp[j]:sp[j + 1]] for j in range(len(sp) - 1)]
    path = g[f"path_{tag}"]
    meta = dict(key=g[f"key_{tag}"], layer=path[:, 0].astype(np.int32), path=[tuple(int(x) for x in r[1:1 + r[0]]) for r in path],
                centre=g[f"centre_{tag}"], direct=g[f"direct_{tag}"], eigenvalues=g[f"eigenvalues_{tag}"])
    ref = (g[f"vox_ptr_{tag}"], g[f"pose_idx_{tag}"], g[f"clusters_{tag}"], meta)
    return scans, g["poses"], float(g["voxel_sizes"]["ab".index(tag)]), ref, g["X"], g[f"plane_nd_{tag}"]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pipeline_equals_golden_fixture(emu, tag):
    scans, poses, voxel_size, ref, X, plane_nd = golden_case(tag)
    m = EmuMap(emu, scans, poses, voxel_size)
    compare_with_oracle(m.export(), ref)
    compare_lookup(m.lookup(X), plane_nd)
    m.close()


def oracle_windows(scans, poses, sizes, voxel_size):
    """One oracle map per window (runWindowBA builds a fresh surf_map per window, lvba_system.cpp:247-258), concatenated with
    global pose indices — the layout lvba_lidar_lm_batch takes."""
    vps, pis, cls, wins, keys, layers = [np.zeros(1, np.int64)], [], [], [], [], []
    a = 0; off = 0
    for w, n in enumerate(sizes):
        vp, pi, cl, meta = vox.voxelize(scans[a:a + n], poses[a:a + n], voxel_size)
        vps.append(vp[1:] + off); off += int(vp[-1]); pis.append(pi + a); cls.append(cl)
        wins.append(np.full(len(vp) - 1, w, np.int32)); keys.append(meta["key"]); layers.append(meta["layer"])
        a += n
    return (np.concatenate(vps), np.concatenate(pis).astype(np.int32), np.concatenate(cls), np.concatenate(wins),
            np.concatenate(keys), np.concatenate(layers))


def test_windowed_map_equals_one_oracle_map_per_window(emu):
    sizes = [5, 4, 1, 6]
    scans, poses = _scene(13, W=sum(sizes), n_per_scan=1800)
    win_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    sp = np.zeros(len(scans) + 1, np.int64); sp[1:] = np.cumsum([len(s) for s in scans])
    xyz = np.ascontiguousarray(np.concatenate(scans), np.float32)
    er = np.asarray(vox.EIGEN_RATIO_DEFAULT, np.float32)
    for voxel_size in (1.0, 2.5):
        m = EmuMap.__new__(EmuMap); m.lib = emu; m.h = ctypes.c_void_p()
        rc = emu.emu_voxel_map_create_windows(ctypes.c_int32(len(sizes)), _ptr(win_ptr, ctypes.c_int32), _ptr(sp, ctypes.c_int64),
                                              _ptr(xyz, ctypes.c_float), _ptr(np.ascontiguousarray(poses), ctypes.c_double),
                                              ctypes.c_double(voxel_size), _ptr(er, ctypes.c_float), ctypes.c_int32(2), ctypes.c_int32(15),
                                              ctypes.byref(m.h))
        assert rc == 0
        got = m.export()
        win = np.zeros(len(got["vox_ptr"]) - 1, np.int32)
        emu.emu_voxel_map_windows(m.h, _ptr(win, ctypes.c_int32))
        vp, pi, cl, wins, keys, layers = oracle_windows(scans, poses, sizes, voxel_size)
        assert np.array_equal(got["vox_ptr"], vp) and np.array_equal(got["pose_idx"], pi) and np.array_equal(win, wins)
        assert np.array_equal(got["key"], keys) and np.array_equal(got["path"][:, 0], layers)
        assert np.abs(got["clusters"] - cl).max() <= 1e-12 * max(1.0, np.abs(cl).max())
        assert len(set(wins.tolist())) >= 3 and 2 not in wins          # the one-scan window has no voxel seen from two poses
        # every voxel lies inside ONE window (what lvba_lidar_lm_batch requires)
        for v in range(len(vp) - 1):
            p = pi[vp[v]:vp[v + 1]]
            assert win_ptr[wins[v]] <= p.min() and p.max() < win_ptr[wins[v] + 1]
        assert emu.emu_voxel_map_lookup(m.h, ctypes.c_int64(1), _ptr(np.zeros(3), ctypes.c_double), _ptr(np.zeros(4), ctypes.c_double)) == -4
        m.close()


def test_randomised_scenes_and_options(emu):
    """Forty random combinations of scan count, density, root size, layer limit, min_ps and a random offset of the scene."""
    rng = np.random.default_rng(123)
    for seed in range(100, 140):
        W = int(rng.integers(2, 8)); npts = int(rng.integers(300, 3000))
        vs = float(rng.choice([0.3, 0.5, 0.8, 1.0, 1.7, 2.5, 4.0])); ll = int(rng.integers(0, 3)); mp = int(rng.choice([5, 15, 30]))
        scans, poses = _scene(seed, W=W, n_per_scan=npts)
        poses = poses.copy(); poses[:, 9:] += rng.uniform(-50, 50, 3)
        m = EmuMap(emu, scans, poses, vs, layer_limit=ll, min_ps=mp)
        assert m.rc == 0
        compare_with_oracle(m.export(), vox.voxelize(scans, poses, vs, layer_limit=ll, min_ps=mp))
        m.close()
