"""rxg_allgather_posteriors on one GPU (a 1-rank NCCL communicator): the full gather and the
RXG_COV_REPLICATE variant (means over NCCL, chain-independent covariances replicated locally) must fill
the gathered buffers with identical bits.  Multi-rank behaviour is exercised by bench.py --gpus N (it
asserts the same equality on every rank) and, for the host logic, by tests/test_multigpu_gloo.py."""
import numpy as np
import pytest
import torch

from oracle import lgssm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [64, 203])          # 203: rows of the gathered slabs are not 16-byte aligned
def test_replicated_cov_equals_full_gather(rx, ctx, batch):
    from rxinfer_jl_b200.context import comm_unique_id
    if not getattr(ctx, "_comm1", False):
        ctx.comm_init(1, 0, comm_unique_id())
        ctx._comm1 = True
    mod = {k: np.asarray(v, np.float32) for k, v in lgssm.notebook_model(4).items()}
    _, y = lgssm.generate_data({k: v.astype(np.float64) for k, v in mod.items()}, 37, batch, seed=3)
    yd = torch.as_tensor(y, device="cuda")
    r = ctx.lgssm(yd, **mod, smooth=True)
    gm, gc = ctx.allgather_posteriors(r["mean"], r["cov"], 1)
    assert torch.equal(gm[0], r["mean"]) and torch.equal(gc[0], r["cov"])
    gm2, gc2 = ctx.allgather_posteriors(r["mean"], r["cov"], 1, replicate_cov=True)
    assert torch.equal(gm2, gm) and torch.equal(gc2, gc)
    # from the de-duplicated [T, d, d] table (RXG_COV_SHARED_OUT)
    rs = ctx.lgssm(yd, **mod, smooth=True, cov_shared_out=True)
    assert rs["cov"].dim() == 3
    gm3, gc3 = ctx.allgather_posteriors(rs["mean"], rs["cov"], 1, replicate_cov=True)
    assert torch.equal(gm3, gm) and torch.equal(gc3, gc)
    with pytest.raises(ValueError):
        ctx.allgather_posteriors(rs["mean"], rs["cov"], 1)
