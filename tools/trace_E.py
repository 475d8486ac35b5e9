"""Per-iteration LM trace of path A at a config (default E): GPU twice (run-to-run reproducibility: the Hessian scatter uses
RED.ADD.F64, whose order is not fixed) and the CPU restatement."""
import sys
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
from oracle import synth, cpu_ref
pkg = g.load_package(); pkg.load_library()
cfg = sys.argv[1] if len(sys.argv) > 1 else "E"
p = synth.make_config(cfg)
def gpu_trace():
    P = pkg.LidarProblem(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
    P.reset_lm()
    out = []
    for i in range(10):
        s = P.iterate(1)
        out.append((s["cost_last"], s["accepted"], s["damping_last"]))
        if s["termination"] != 0 and i > 0 and s["iterations"] == 0:
            break
    P.close()
    return out
a = gpu_trace(); b = gpu_trace()
for i, (x, y) in enumerate(zip(a, b)):
    print(f"it {i}: gpu run1 cost {x[0]:.12e} acc {x[1]} u {x[2]:.3e} | run2 cost {y[0]:.12e} acc {y[1]} u {y[2]:.3e} | rel {abs(x[0]-y[0])/abs(x[0]):.2e}")
for it in (1, 2, 3, 4, 6, 10):
    _, c = cpu_ref.lidar_lm(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], max_iter=it, rel_tol=-1.0, threads=16)
    P = pkg.LidarProblem(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"]); o = pkg.lidar_default_opts(); o.rel_tol = -1.0; P.reset_lm(o); s = P.iterate(it); P.close()
    print(f"after {it} iterations: cpu cost {c['cost_last']:.12e} (acc {int(c['accepted'])})  gpu cost {s['cost_last']:.12e} (acc {s['accepted']})  rel {abs(c['cost_last']-s['cost_last'])/abs(c['cost_last']):.2e}")
