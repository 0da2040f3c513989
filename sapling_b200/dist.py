"""Multi-GPU sharding of the query path: the index is replicated on every GPU, the query stream is cut
into one contiguous slice per rank, and there is NO collective on the query path.  The only communication is the
optional gather of results to one rank and the max-over-ranks reduction of timings.

One process per GPU (torchrun); `torch.distributed` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_bounds(nq: int, world: int) -> List[Tuple[int, int]]:
    """Slice g = [g*nq/G, (g+1)*nq/G): contiguous, disjoint, covering, sizes differing by at most one."""
    if world < 1:
        raise ValueError("world must be >= 1")
    return [((g * nq) // world, ((g + 1) * nq) // world) for g in range(world)]


def my_shard(nq: int, rank: int, world: int) -> Tuple[int, int]:
    return shard_bounds(nq, world)[rank]


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Elapsed time of a multi-GPU step = the slowest rank's (never the wall clock of one rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_query(engine: Callable[[np.ndarray], np.ndarray], kmers: np.ndarray, rank: int, world: int, dist=None,
                  gather_to: Optional[int] = 0, device=None):
    """Every rank answers its own slice of `kmers` with `engine` (e.g. ``Sapling.queryBatch`` of its replica).

    Returns (lo, hi, local_results, all_results): all_results is the full answer vector on rank `gather_to` (None on
    the others, and None everywhere when gather_to is None).  The gather is a convenience for callers that want the
    answers in one place; throughput runs leave the results sharded.
    """
    lo, hi = my_shard(len(kmers), rank, world)
    local = np.ascontiguousarray(engine(kmers[lo:hi]), dtype=np.int64)
    if gather_to is None or world == 1 or dist is None:
        return lo, hi, local, (local if world == 1 and gather_to is not None else None)
    import torch
    bounds = shard_bounds(len(kmers), world)
    longest = max(h - l for l, h in bounds)
    buf = torch.full((longest,), -2, dtype=torch.int64, device=device or "cpu")
    buf[: hi - lo] = torch.from_numpy(local).to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == gather_to else None
    dist.gather(buf, parts, dst=gather_to)
    if rank != gather_to:
        return lo, hi, local, None
    out = np.empty(len(kmers), dtype=np.int64)
    for (l, h), p in zip(bounds, parts):
        out[l:h] = p[: h - l].cpu().numpy()
    return lo, hi, local, out
