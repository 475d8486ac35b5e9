"""world_size-2 gloo test (CPU): the sharding rule of SURVEY.md §8(e) — every voxel / landmark goes to the
owner of its lowest pose index, partial H / g / cost are summed with an all-reduce — reproduces the
single-process result.  The per-shard arithmetic here is the numpy oracle; the shard rule is the
library's own host function lvba_shard_owner."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as graft
    from oracle import lidar_oracle as lo, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = graft.load_package(); pkg.build_library(); pkg.load_library()
    p = synth.make_problem(40, 500, 0, seed=21, visual=False)
    W = 40
    vp, pi, cl = p["vox_ptr"], p["pose_idx"], p["clusters"]
    mine = [a for a in range(len(vp) - 1) if pkg.shard_owner(int(pi[vp[a]]), W, world) == rank]
    sl = np.concatenate([np.arange(vp[a], vp[a + 1]) for a in mine])
    lvp = np.concatenate([[0], np.cumsum([vp[a + 1] - vp[a] for a in mine])]).astype(np.int64)
    r, g, blocks = lo.acc_evaluate2(lvp, pi[sl], cl[sl], p["poses"], W)
    H = torch.from_numpy(lo.assemble_dense(blocks, W)); gt = torch.from_numpy(g.copy()); rt = torch.tensor([r, float(len(mine))], dtype=torch.float64)
    dist.all_reduce(H); dist.all_reduce(gt); dist.all_reduce(rt)
    if rank == 0:
        r0, g0, b0 = lo.acc_evaluate2(vp, pi, cl, p["poses"], W)
        H0 = lo.assemble_dense(b0, W)
        q.put((abs(rt[0].item() - r0) / r0, float(np.abs(gt.numpy() - g0).max() / np.abs(g0).max()),
               float(np.abs(H.numpy() - H0).max() / np.abs(H0).max()), rt[1].item(), len(vp) - 1))
    dist.destroy_process_group()


def test_two_rank_shard_sum_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    er, eg, eh, nsum, nvox = res
    assert nsum == nvox                       # every voxel owned by exactly one rank
    assert er <= 1e-12 and eg <= 1e-12 and eh <= 1e-12
