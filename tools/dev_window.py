"""Development timing of the batched window BA against per-window calls (GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import __graft_entry__ as graft
from oracle import synth

pkg = graft.load_package()
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 100
p = synth.make_window_problem([20] * nw, 300, seed=9)
for rep in range(2):
    t0 = time.perf_counter()
    poses, sums, tot = pkg.lidar_lm_batch(p["win_ptr"], p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
    t1 = time.perf_counter()
print(f"batch: {nw} windows x 20 poses x 300 voxels: {1e3 * (t1 - t0):.2f} ms wall; device build {tot['ms_build']:.2f} solve {tot['ms_solve']:.2f} "
      f"resid {tot['ms_residual']:.2f} ms; passes {tot['iterations']}; launches {tot['kernel_launches']}")
t0 = time.perf_counter()
for w, win in enumerate(p["windows"]):
    pkg.lidar_lm(win["vox_ptr"], win["pose_idx"], win["clusters"], win["poses"])
t1 = time.perf_counter()
print(f"separate lvba_lidar_lm calls: {1e3 * (t1 - t0):.2f} ms wall ({1e3 * (t1 - t0) / nw:.3f} ms per window)")
