"""Multi-GPU: shard the batch axis over ranks, no communication during the sweeps, one all-gather
of posterior marginals at the end.  One process per GPU, launched with
``torch.distributed.run``; ``torch.distributed`` is used for rendezvous only (NCCL unique-id
broadcast), the gather itself runs through ``rxg_allgather_posteriors`` on the context's stream.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int):
    """Contiguous split: rank g owns chains [lo, hi); remainders go to the low ranks."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def assemble_gathered(g: torch.Tensor) -> torch.Tensor:
    """[G, T, ..., b_local] (rank-major slabs, as NCCL delivers them) -> [T, ..., G * b_local]."""
    G = g.shape[0]
    perm = list(range(1, g.dim() - 1)) + [0, g.dim() - 1]
    out = g.permute(*perm).contiguous()
    return out.reshape(*out.shape[:-2], G * g.shape[-1])


def init_comm(ctx, group=None):
    """Create the NCCL communicator behind ``ctx`` (id generated on rank 0, broadcast over the
    default process group)."""
    from .context import comm_unique_id
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(world, rank, box[0])
    return world, rank


def allgather_posteriors(ctx, mean, cov, world, backend="rxg"):
    """Returns rank-major gathered slabs ([G, T, d, b], [G, T, d, d, b])."""
    if backend == "rxg":
        return ctx.allgather_posteriors(mean, cov, world)
    # host-logic path for gloo tests: same layout through torch.distributed
    gm = [torch.empty_like(mean) for _ in range(world)]
    dist.all_gather(gm, mean)
    gc = None
    if cov is not None:
        gc = [torch.empty_like(cov) for _ in range(world)]
        dist.all_gather(gc, cov)
        gc = torch.stack(gc)
    return torch.stack(gm), gc
