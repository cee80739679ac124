"""Multi-GPU execution = source sharding: the window graph is replicated on every GPU,
every rank applies every edge batch to its replica, each rank owns a slice of the source vertices,
and NOTHING is exchanged on the hot path.  The only collectives of a job are the max-over-ranks of
the timings and one gather of estimate vectors at the very end.  These helpers are backend
agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations
import numpy as np


def shard_sources(sources, rank: int, world: int, per_rank: int | None = None, interleave: bool = False):
    """Contiguous block of sources for `rank`.  With per_rank given (weak scaling) every rank gets
    exactly that many; otherwise the list is split as evenly as possible (strong scaling).
    interleave (strong scaling only): rank r takes sources r, r + world, r + 2 world, ... -- the job's list is ranked by
    degree, and a refresh costs more for the high-degree sources, so dealing the list round-robin evens the ranks out."""
    sources = np.asarray(sources, dtype=np.int32)
    if interleave:
        if per_rank is not None:
            raise ValueError("interleave applies to the strong-scaling split only")
        return np.ascontiguousarray(sources[rank::world])
    if per_rank is not None:
        if len(sources) < world * per_rank:
            raise ValueError(f"need {world * per_rank} sources for {world} ranks x {per_rank}, have {len(sources)}")
        return sources[rank * per_rank:(rank + 1) * per_rank]
    base, extra = divmod(len(sources), world)
    lo = rank * base + min(rank, extra)
    return sources[lo: lo + base + (1 if rank < extra else 0)]


def max_over_ranks(values, device="cpu"):
    """element-wise MAX over ranks of a list of floats (timings are reported as the slowest rank)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_estimates(local, dst: int = 0):
    """gather a [n_local_sources, V] float64 tensor from every rank onto `dst` (list, rank order)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    # ranks may own different numbers of sources: exchange the row counts first
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    out = None
    if rank == dst:
        out = [torch.empty((int(c.item()), local.shape[1]), dtype=local.dtype, device=local.device) for c in counts]
    if all(int(c.item()) == local.shape[0] for c in counts):
        dist.gather(local.contiguous(), out, dst=dst)
    else:  # ragged: point-to-point
        if rank == dst:
            out[dst].copy_(local)
            for r in range(world):
                if r != dst:
                    dist.recv(out[r], src=r)
        else:
            dist.send(local.contiguous(), dst=dst)
    return out
