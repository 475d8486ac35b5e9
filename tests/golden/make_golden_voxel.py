"""Regenerates tests/golden/voxel_scene.npz — a small scan set with the LITERAL restatement's voxel map and plane
lookups (oracle/voxel_oracle.py: build_tree_literal / voxelize_literal / plane_lookup_literal, which follow
include/BALM/bavoxel.hpp:320-474, 799-836 and src/lvba_system.cpp:1529-1566 statement by statement).
The reference ships no fixtures for this stage either; the file pins the oracle over time and gives the
device path committed known-answer vectors.  Run from the repo root:  python tests/golden/make_golden_voxel.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import synth, voxel_oracle as vox  # noqa: E402

scans, poses = synth.make_scan_scene(1, W=4, n_per_scan=2000)
rng = np.random.default_rng(5)
X = np.concatenate([rng.uniform(-3.5, 3.5, (400, 3)), np.column_stack([rng.uniform(-3, 3, (200, 2)), np.full(200, -1.2)])])
out = dict(xyz=np.concatenate(scans), scan_ptr=np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.int64),
           poses=poses, X=X, voxel_sizes=np.array([2.5, 3.5]))
for tag, voxel_size in zip("ab", out["voxel_sizes"]):         # 2.5 m: planes at layers 0 and 2; 3.5 m: layers 1 and 2
    vp, pi, cl, meta = vox.voxelize_literal(scans, poses, voxel_size)
    path = np.full((len(vp) - 1, 3), -1, np.int8)
    path[:, 0] = meta["layer"]
    for v, p in enumerate(meta["path"]):
        path[v, 1:1 + len(p)] = p
    roots = vox.build_tree_literal(scans, poses, voxel_size)
    out.update({f"vox_ptr_{tag}": vp, f"pose_idx_{tag}": pi, f"clusters_{tag}": cl, f"key_{tag}": meta["key"], f"path_{tag}": path,
                f"centre_{tag}": meta["centre"], f"direct_{tag}": meta["direct"], f"eigenvalues_{tag}": meta["eigenvalues"],
                f"plane_nd_{tag}": vox.plane_lookup_literal(roots, X, voxel_size)})
    print(tag, voxel_size, len(vp) - 1, "voxels, layers", np.bincount(meta["layer"], minlength=3), "lookup hits",
          int(np.any(out[f"plane_nd_{tag}"] != 0, axis=1).sum()))
np.savez_compressed(Path(__file__).with_name("voxel_scene.npz"), **out)
