import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); pkg.load_library()
p = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C")
def pin(a): return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hp = {k: pin(p[k]) for k in ("vox_ptr", "pose_idx", "clusters", "poses", "q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr")}
for i in range(3):
    t0 = time.perf_counter(); _, sa = pkg.lidar_lm(hp["vox_ptr"], hp["pose_idx"], hp["clusters"], hp["poses"]); t1 = time.perf_counter()
    _, _, _, sb = pkg.visual_lm(hp["q"], hp["t"], hp["X"], hp["plane_nd"], hp["obs_ptr"], hp["obs_cam"], hp["obs_uv"], hp["intr"], p["sigma_px"], p["sigma_plane"]); t2 = time.perf_counter()
    print(f"call {i}: lidar {1e3*(t1-t0):.1f} ms (setup {sa['ms_setup']:.1f}, iters {sa['iterations']}, dev {sa['ms_build']+sa['ms_solve']+sa['ms_residual']:.1f})  visual {1e3*(t2-t1):.1f} ms (setup {sb['ms_setup']:.1f}, iters {sb['iterations']}, dev {sb['ms_build']+sb['ms_solve']+sb['ms_residual']:.1f})", flush=True)
