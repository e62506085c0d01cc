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


class PeerGroup:
    """Peer-mapped gathered buffers of one rank (``rxg_peer_*``): allocates this rank's ``[G, T, d, b]`` /
    ``[G, T, d, d, b]`` buffers and flag words, exchanges the CUDA IPC handles over ``torch.distributed`` (plumbing
    only) and maps the peers' buffers, so that the fused sweep can store its posteriors into every rank's buffer.

    ``PeerGroup.local(ctxs, ...)`` builds the same thing for several contexts inside ONE process (tests on a single
    GPU: every "rank" is a context with its own stream; no IPC involved)."""

    def __init__(self, ctx, T, d, b_local, with_cov=True, group=None, _local=None):
        from .context import DeviceBuffer
        self.ctx, self.T, self.d, self.b = ctx, T, d, b_local
        if _local is None:
            self.world, self.rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_initialized() else (1, 0)
        else:
            self.world, self.rank = _local
        G = self.world
        self.buf_mean = DeviceBuffer(ctx, 4 * G * T * d * b_local)
        self.buf_cov = DeviceBuffer(ctx, 4 * G * T * d * d * b_local) if with_cov else None
        self.buf_flags = DeviceBuffer(ctx, 4 * 8, zero=True)
        self.mean = self.buf_mean.tensor(G, T, d, b_local)
        self.cov = self.buf_cov.tensor(G, T, d, d, b_local) if with_cov else None
        self._opened = []
        if _local is not None:
            return                      # PeerGroup.local wires the pointers
        mine = (self.buf_mean.export(), self.buf_cov.export() if with_cov else None, self.buf_flags.export())
        if G > 1:
            allh = [None] * G
            dist.all_gather_object(allh, mine, group=group)
        else:
            allh = [mine]
        self.mean_ptrs, self.cov_ptrs, flag_ptrs = [], [], []
        for g in range(G):
            if g == self.rank:
                self.mean_ptrs.append(self.buf_mean.ptr)
                self.cov_ptrs.append(self.buf_cov.ptr if with_cov else 0)
                flag_ptrs.append(self.buf_flags.ptr)
            else:
                hm, hc, hf = allh[g]
                pm, pf = ctx.peer_open(hm), ctx.peer_open(hf)
                pc = ctx.peer_open(hc) if with_cov else 0
                self._opened += [p for p in (pm, pc, pf) if p]
                self.mean_ptrs.append(pm); self.cov_ptrs.append(pc); flag_ptrs.append(pf)
        ctx.peer_group(G, self.rank, flag_ptrs)
        if G > 1:
            dist.barrier(group=group)

    @classmethod
    def local(cls, ctxs, T, d, b_local, with_cov=True):
        groups = [cls(c, T, d, b_local, with_cov, _local=(len(ctxs), r)) for r, c in enumerate(ctxs)]
        for gr in groups:
            gr.mean_ptrs = [o.buf_mean.ptr for o in groups]
            gr.cov_ptrs = [o.buf_cov.ptr if with_cov else 0 for o in groups]
            gr.ctx.peer_group(len(ctxs), gr.rank, [o.buf_flags.ptr for o in groups])
        return groups

    def smooth_gather(self, y, model, *, replicate_cov=False, **kw):
        cov_ptrs = self.cov_ptrs if self.cov is not None else None
        return self.ctx.lgssm_smooth_gather(y, model["A"], model["B"], model["P"], model["Q"], model["m0"], model["S0"],
                                            self.mean_ptrs, cov_ptrs, replicate_cov=replicate_cov, **kw)

    def close(self):
        for p in self._opened:
            try:
                self.ctx.peer_close(p)
            except Exception:
                pass
        self._opened = []


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
