"""Coordinate-sharded data parallelism: one process per GPU, every rank holds the full
(<= 1.1 MB) weights and a contiguous shard of each coordinate batch; the only data-path collective is ONE
all-reduce (SUM) of the flat fp32 weight-gradient buffer per step (complex gradients as re,im pairs —
complex dtypes are not NCCL types).  The reference has no distributed code; this is the B200-native
counterpart of its chunked loops (wire_occupancy.py:137-154).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n work items; shard sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    for t in list(module.parameters()) + list(module.buffers()):
        data = torch.view_as_real(t.data) if t.is_complex() else t.data
        dist.broadcast(data, src, group=group)


def _real_view(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


def flatten_grads(params: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([_real_view(p.grad).reshape(-1) for p in params])


def unflatten_into_grads(flat: torch.Tensor, params: Sequence[torch.Tensor]) -> None:
    off = 0
    for p in params:
        g = _real_view(p.grad)
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def allreduce_gradients(params: Iterable[torch.Tensor], world: int, group=None, average: bool = True) -> None:
    """One collective per step over one flat buffer. With a mean-over-local-shard loss and equal shard
    sizes, averaging over ranks yields the gradient of the mean over the global batch."""
    params = [p for p in params if p.grad is not None]
    if not params or world <= 1:
        return
    flat = flatten_grads(params)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    unflatten_into_grads(flat, params)


def weighted_allreduce_gradients(params: Iterable[torch.Tensor], n_local: int, group=None) -> None:
    """Unequal shards: scale by n_local / n_global so the result is the gradient of the global mean."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    flat = flatten_grads(params)
    cnt = torch.tensor([float(n_local)], device=flat.device)
    flat.mul_(float(n_local))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    flat.div_(cnt)
    unflatten_into_grads(flat, params)
