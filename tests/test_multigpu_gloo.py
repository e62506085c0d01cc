"""world_size-2 gloo test of the N>1 host path (no GPU): batch sharding, per-rank sweeps (the
oracle stands in for the device sweep -- this test covers the HOST logic only), all-gather with
the rank-major slab layout NCCL delivers, assembly into the global [T, d, batch] layout."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, T, batch, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rxinfer_jl_b200 as rx
    from oracle import lgssm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mod = lgssm.notebook_model(2)
    _, y = lgssm.generate_data(mod, T, batch, seed=5)
    lo, hi = rx.sharding.shard_bounds(batch, world, rank)
    r = lgssm.smooth_reference_schedule(y[:, :, lo:hi], **mod)
    mean = torch.from_numpy(r["mean"]).float().contiguous()
    cov = torch.from_numpy(r["cov"]).float().contiguous()
    gm, gc = rx.sharding.allgather_posteriors(None, mean, cov, world, backend="torch")
    full_m = rx.sharding.assemble_gathered(gm)
    full_c = rx.sharding.assemble_gathered(gc)
    if rank == 0:
        q.put((full_m.numpy(), full_c.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_matches_single_process():
    from oracle import lgssm
    T, batch, world = 12, 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    full_m, full_c = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    mod = lgssm.notebook_model(2)
    _, y = lgssm.generate_data(mod, T, batch, seed=5)
    ref = lgssm.smooth_reference_schedule(y, **mod)
    assert np.abs(full_m - ref["mean"]).max() < 1e-5
    assert np.abs(full_c - ref["cov"]).max() < 1e-4
