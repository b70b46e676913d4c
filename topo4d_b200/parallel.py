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


# Flat gradient buffers that live in NVLink SYMMETRIC MEMORY (every rank's copy mapped into every peer, plus the NVSwitch
# multicast address when the fabric offers one).  For them the exchange is the two-shot in-switch reduction
# (multimem.ld_reduce of the rank's 1/G slice straight out of all peers, multimem.st of the result into all of them) that
# PyTorch's symmetric-memory runtime ships, instead of NCCL's ring/tree kernels: measured inside the 8-rank step of bench.py
# (tools/probe_allreduce.py, 14.9 MB): +54 us per step against +103 us for ncclAllReduce (0.414 vs 0.463 ms; no exchange 0.360 ms).
# At 2 and 4 ranks NCCL is as fast or faster (0.615 vs 0.627 ms at 4), so the symmetric path is only taken from 8 ranks up.
_SYMM: dict[int, object] = {}          # data_ptr -> rendezvous handle (keeps the mapping alive)
SYMM_MIN_WORLD = 8


def symmetric_flat(numel: int, device) -> torch.Tensor | None:
    """A zero-copy exchangeable fp32 buffer, or None when symmetric memory is unavailable / not worth it (then use a plain
    tensor: allreduce_flat_ falls back to NCCL).  Collective: every rank must call it with the same `numel`."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() < SYMM_MIN_WORLD:
        return None
    try:
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(numel, dtype=torch.float32, device=device)
        hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
        _SYMM[buf.data_ptr()] = hdl
        return buf
    except Exception:  # noqa: BLE001  (driver / fabric without symmetric-memory support)
        return None


def allreduce_flat_(flat: torch.Tensor, average: bool = False) -> torch.Tensor:
    """In-place sum (or mean) of the contiguous fp32 gradient buffer across ranks: ONE collective call."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    assert flat.is_contiguous()
    hdl = _SYMM.get(flat.data_ptr())
    if hdl is not None:
        name = dist.group.WORLD.group_name
        if getattr(hdl, "multicast_ptr", 0):
            torch.ops.symm_mem.multimem_all_reduce_(flat, "sum", name)
        else:
            torch.ops.symm_mem.two_shot_all_reduce_(flat, "sum", name)
    else:
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
