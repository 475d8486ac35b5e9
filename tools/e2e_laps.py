"""One process, one generation of the config: the one-shot ABI calls (lvba_lidar_lm / lvba_visual_lm from pinned host buffers) timed
N times, then once with LVBA_SETUP_TIMING=1 (host laps of the set-up on stderr).  GPU box.
    python tools/e2e_laps.py [C] [calls]"""
import os, sys, time, statistics
sys.path.insert(0, '.')
import numpy as np, torch
import __graft_entry__ as g
from oracle import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.cuda.set_device(0)
pkg = g.load_package(); pkg.load_library()
t0 = time.perf_counter(); p = synth.make_config(cfg); print(f"config {cfg} generated in {time.perf_counter() - t0:.1f} s", flush=True)
def pin(a): return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hp = {k: pin(p[k]) for k in ("vox_ptr", "pose_idx", "clusters", "poses", "q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr")}
def call():
    torch.cuda.synchronize()
    t0 = time.perf_counter(); _, sa = pkg.lidar_lm(hp["vox_ptr"], hp["pose_idx"], hp["clusters"], hp["poses"]); t1 = time.perf_counter()
    _, _, _, sb = pkg.visual_lm(hp["q"], hp["t"], hp["X"], hp["plane_nd"], hp["obs_ptr"], hp["obs_cam"], hp["obs_uv"], hp["intr"], p["sigma_px"], p["sigma_plane"]); t2 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1), sa, sb
call()
ta, tb = [], []
for i in range(n_calls):
    a, b, sa, sb = call(); ta.append(a); tb.append(b)
    print(f"call {i}: lidar {a:.2f} ms (setup {sa['ms_setup']:.2f}, passes {sa['iterations']}, builds {sa['hessian_builds']}, dev {sa['ms_build']+sa['ms_solve']+sa['ms_residual']:.2f})  "
          f"visual {b:.2f} ms (setup {sb['ms_setup']:.2f}, passes {sb['iterations']}, dev {sb['ms_build']+sb['ms_solve']+sb['ms_residual']:.2f})", flush=True)
ma, mb = statistics.median(ta), statistics.median(tb)
pa, pb = max(sa["iterations"], 1), max(sb["iterations"], 1)
print(f"median: lidar {ma:.2f} ms / {pa} passes, visual {mb:.2f} ms / {pb} passes -> e2e {1e3 / (ma / pa + mb / pb):.1f} LM it/s; costs A {sa['cost_last']:.13e} B {sb['cost_last']:.13e}", flush=True)
os.environ["LVBA_SETUP_TIMING"] = "1"
call()
