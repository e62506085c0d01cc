"""Worker of tests/test_peer_gather_gpu.py::test_two_processes_cuda_ipc (launched with torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import rxinfer_jl_b200 as rx                      # noqa: E402
from oracle import lgssm                           # noqa: E402
from rxinfer_jl_b200.sharding import PeerGroup, assemble_gathered   # noqa: E402
from util import TOL_MEAN, f32_model, rel_l2       # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    ctx = rx.Context(dev)
    mod = f32_model(lgssm.notebook_model(4))
    T, b = 50, 128
    _, y = lgssm.generate_data(mod, T, world * b, seed=11)
    ys = torch.as_tensor(np.ascontiguousarray(y[:, :, rank * b:(rank + 1) * b]), device=f"cuda:{dev}")
    kw = dict(A=mod["A"], B=mod["B"], P=mod["P"], Q=mod["Q"], m0=mod["m0"], S0=mod["S0"])
    own = ctx.lgssm(ys, **kw, smooth=True)
    grp = PeerGroup(ctx, T, 4, b)
    ref = lgssm.smooth_reference_schedule(y, **mod)
    for replicate in (False, True, False):
        grp.mean.zero_(); grp.cov.zero_()
        torch.cuda.synchronize(); dist.barrier()
        grp.smooth_gather(ys, mod, replicate_cov=replicate)
        assert torch.equal(grp.mean[rank], own["mean"]) and torch.equal(grp.cov[rank], own["cov"])
        assert rel_l2(assemble_gathered(grp.mean).cpu().numpy(), ref["mean"]) < TOL_MEAN
        c = assemble_gathered(grp.cov)
        assert torch.equal(c[..., 0], c[..., world * b - 1])          # shared model: chain-independent covariances
        dist.barrier()
    # masked data => per-chain path => slab push
    rng = np.random.default_rng(3)
    mask = (rng.random((T, world * b)) > 0.3).astype(np.uint8)
    ms = torch.as_tensor(np.ascontiguousarray(mask[:, rank * b:(rank + 1) * b]), device=f"cuda:{dev}")
    grp.mean.zero_(); grp.cov.zero_()
    torch.cuda.synchronize(); dist.barrier()
    grp.smooth_gather(ys, mod, mask=ms)
    refm = lgssm.smooth_reference_schedule(y, **mod, mask=mask)
    assert rel_l2(assemble_gathered(grp.mean).cpu().numpy(), refm["mean"]) < TOL_MEAN
    dist.barrier()
    grp.close()
    print("PEER_WORKER_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
