"""Window sharding across the GPUs of one node.  Windows are independent units:
rank r of G owns the contiguous range [r*W/G, (r+1)*W/G); there is no collective on the solve
path, only one gather of the results to rank 0 afterwards (torch.distributed: NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def window_range(n_windows: int, rank: int, world: int) -> tuple[int, int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (n_windows * rank) // world, (n_windows * (rank + 1)) // world


def shard_batch(batch, rank: int, world: int):
    lo, hi = window_range(batch.n_windows, rank, world)
    return batch.slice(lo, hi)


def gather_to_root(local, dst: int = 0):
    """Gather per-rank result tensors (first dim = windows, ragged across ranks allowed) to `dst`.
    `local` is a torch tensor on the device of the process group's backend.  Returns the
    concatenation on dst, None elsewhere.  Single process: returns `local`."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])
