"""dev helper: time path A phases on a named config (GPU box)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); pkg.load_library()
name = sys.argv[1] if len(sys.argv) > 1 else "C"
t = time.time(); p = synth.make_config(name, visual=False); print("gen", name, time.time() - t, flush=True)
t = time.time(); P = pkg.LidarProblem(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"]); print("create", time.time() - t)
print(P.counts(nonzero=False))
for rep in range(3):
    t = time.time(); r = P.build(); t1 = time.time() - t
    t = time.time(); dx = P.solve(0.01); t2 = time.time() - t
    t = time.time(); r2 = P.residual(); t3 = time.time() - t
    print(f"build {t1*1e3:.3f} ms  solve {t2*1e3:.3f} ms  resid {t3*1e3:.3f} ms   r={r/P.V:.6e} |dx|={np.abs(dx).max():.3e}")
P.reset_lm()
s = P.iterate(10)
print(json.dumps(s, indent=1))
