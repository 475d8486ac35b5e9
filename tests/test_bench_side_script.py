"""tools/bench_voxel_map.py (the set-up-stage side measurement bench.py attaches as `voxel_map`) needs a GPU for its numbers,
but not for its own plumbing: with the two device classes replaced by stand-ins of the same shape the script must run to the
end and print well-formed JSON — so that a Python-level slip cannot cost the round-end bench line its side measurement."""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


class _FakeMap:
    def __init__(self, xyz, poses, vs, scan_ptr=None):
        self.summary = dict(n_voxels=3, nnz=7, n_nodes=[5, 2, 0], ms_device=1.5, ms_total=3.0, ms_upload=1.0, kernel_launches=40, h2d_bytes=123, n_points=len(xyz))

    def close(self):
        pass

    def export(self):
        return dict(vox_ptr=np.array([0, 2, 4, 7]), pose_idx=np.array([0, 1, 0, 2, 1, 2, 3], np.int32),
                    clusters=np.concatenate([np.zeros((7, 9)), np.full((7, 1), 9.0)], 1), key=np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0]]),
                    path=np.zeros((3, 3), np.int8), centre=np.zeros((3, 3)), normal=np.zeros((3, 3)), eigenvalues=np.ones((3, 3)))

    def lookup(self, X):
        return np.zeros((len(X), 4))


class _FakeGrid:
    def __init__(self, xyz, poses, ts, vs, scan_ptr=None):
        self.summary = dict(ms_device=2.0, n_voxels=10, n_pairs=20)

    def render(self, cams, ts, intr, w, h, half_window=0.5):
        return np.zeros((len(cams), h, w), np.float32), dict(ms_device=1.0, ms_total=2.0, work_chunks=5, kernel_launches=9)

    def backproject(self, cams, ts, intr, w, h, kp, uv, half_window=0.5):
        return np.zeros((len(uv), 3)), np.zeros(len(uv), np.uint8), dict(ms_total=1.0, ms_device=0.5, d2h_bytes=25 * len(uv))

    def close(self):
        pass


def test_side_script_runs_to_the_end_with_stand_in_device_classes(pkg, monkeypatch, capsys):
    monkeypatch.setattr(pkg, "VoxelMap", _FakeMap)
    monkeypatch.setattr(pkg, "DepthGrid", _FakeGrid)
    monkeypatch.setattr(pkg, "device_count", lambda: 1)
    spec = importlib.util.spec_from_file_location("bench_voxel_map", ROOT / "tools" / "bench_voxel_map.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["bench_voxel_map.py", "--scans", "20", "--points", "1500", "--repeats", "1", "--queries", "500", "--cpu-sample-scans", "2"])
    assert mod.main() == 0
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 2                                               # B3 alone first, then with the B4 part
    first, last = json.loads(lines[0]), json.loads(lines[1])
    assert first["depth"] is None and last["depth"]["render"]["images"] == 16 and last["n_points"] == 30000
    assert all(last["checks"].values()) and last["cpu"]["points_per_s"] > 0
    xyz, sp, poses = mod.street_scans(3, 1000)
    assert xyz.shape == (3000, 3) and xyz.dtype == np.float32 and sp.tolist() == [0, 1000, 2000, 3000] and poses.shape == (3, 12)
