"""dev helper: time path B phases on a named config (GPU box)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); pkg.load_library()
name = sys.argv[1] if len(sys.argv) > 1 else "C"
t = time.time(); p = synth.make_config(name, lidar=False); print("gen", name, time.time() - t, flush=True)
K = ("q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr", "sigma_px", "sigma_plane")
t = time.time(); P = pkg.VisualProblem(*[p[k] for k in K]); print("create", time.time() - t)
print(P.counts())
P.reset_lm()
s = P.iterate(50)
print(json.dumps(s, indent=1))
