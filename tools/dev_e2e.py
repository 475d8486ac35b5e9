"""Development timing of the one-shot ABI calls from pinned host buffers (GPU box); under torchrun it initialises
the library's NCCL communicator first.  LVBA_SETUP_TIMING=1 prints the set-up laps."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import __graft_entry__ as g
from oracle import synth
rank = int(os.environ.get("RANK", 0)); lrank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lrank)
pkg = g.load_package(); pkg.load_library()
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    uid = [pkg.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
    pkg.comm_init(world, rank, uid[0], lrank)
p = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C")
def pin(a): return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hp = {k: pin(p[k]) for k in ("vox_ptr", "pose_idx", "clusters", "poses", "q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr")}
for i in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); _, sa = pkg.lidar_lm(hp["vox_ptr"], hp["pose_idx"], hp["clusters"], hp["poses"]); t1 = time.perf_counter()
    _, _, _, sb = pkg.visual_lm(hp["q"], hp["t"], hp["X"], hp["plane_nd"], hp["obs_ptr"], hp["obs_cam"], hp["obs_uv"], hp["intr"], p["sigma_px"], p["sigma_plane"]); t2 = time.perf_counter()
    print(f"rank {rank} call {i}: lidar {1e3*(t1-t0):.1f} ms (setup {sa['ms_setup']:.1f}, iters {sa['iterations']}, dev {sa['ms_build']+sa['ms_solve']+sa['ms_residual']:.1f})  visual {1e3*(t2-t1):.1f} ms (setup {sb['ms_setup']:.1f}, iters {sb['iterations']}, dev {sb['ms_build']+sb['ms_solve']+sb['ms_residual']:.1f})", flush=True)
if world > 1:
    pkg.comm_destroy(); dist.destroy_process_group()
