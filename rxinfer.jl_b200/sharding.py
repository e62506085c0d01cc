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


def padded_shard(batch: int, world: int) -> int:
    """Common slab width of the gathered layout [G][T][...][b_pad]: the device gathers (NCCL all-gather,
    peer-mapped stores) need the same count on every rank, so a batch that does not divide evenly is
    padded to ceil(batch / world) chains per rank; ``assemble_gathered(..., batch=batch)`` drops the pad."""
    return -(-batch // world)


def assemble_gathered(g: torch.Tensor, batch: int | None = None) -> torch.Tensor:
    """[G, T, ..., b_local] (rank-major slabs, as the gather delivers them) -> [T, ..., G * b_local].
    With ``batch`` given, slabs are ``padded_shard`` wide and rank r holds ``shard_bounds(batch, G, r)``
    real chains in its leading columns: the pad columns are dropped."""
    G = g.shape[0]
    if batch is not None and batch != G * g.shape[-1]:
        parts = []
        for r in range(G):
            lo, hi = shard_bounds(batch, G, r)
            parts.append(g[r][..., : hi - lo])
        return torch.cat(parts, dim=-1).contiguous()
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
        # equal counts on every rank are required (ncclAllGather / peer stores): check instead of hanging
        n = torch.tensor([mean.shape[-1]], device=mean.device)
        lo, hi = n.clone(), n.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if int(lo) != int(hi):
            raise ValueError(f"allgather_posteriors: shards differ in width ({int(lo)}..{int(hi)}); pad them to "
                             "sharding.padded_shard(batch, world) chains per rank")
        if world != getattr(ctx, "comm_nranks", world):
            raise ValueError("allgather_posteriors: world size differs from the communicator's")
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
