"""View-parallel data parallelism: one process per GPU, every rank holds a full replica
of the Gaussians and renders views {r, r+G, ...}; one all-reduce (sum) of the flat gradient buffer per
optimiser step over NCCL / NVLink.  The reference has no distributed code (train.py renders one view
per Adam step in one process); this is the only collective on the path and it is a real exchange step.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> list[int]:
    """Round-robin ownership: rank r renders views r, r+world, ... (24 views divide by 1/2/4/8)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_views, world))


def allreduce_flat_(flat: torch.Tensor, average: bool = False) -> torch.Tensor:
    """In-place sum (or mean) of the contiguous fp32 gradient buffer across ranks: ONE collective call."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    assert flat.is_contiguous()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(dist.get_world_size())
    return flat


def allreduce_param_grads_(params, average: bool = True):
    """The exchange of a view-parallel optimiser step for a parameter DICTIONARY (the reference keeps its parameters in
    one, train.py:120-160): every ``.grad`` is packed into one flat fp32 buffer in the dictionary's order, reduced with
    ONE collective (mean by default, so the step size does not depend on the number of ranks) and scattered back in
    place.  Parameters without a gradient are skipped -- identically on every rank, or the buffers would not line up."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return params
    names = [k for k, p in params.items() if p.grad is not None]
    if not names:
        return params
    flat = torch.cat([params[k].grad.reshape(-1).float() for k in names])
    allreduce_flat_(flat, average=average)
    o = 0
    for k in names:
        g = params[k].grad
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    return params
