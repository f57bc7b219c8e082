"""Coordinate-sharded data parallelism: one process per GPU, every rank holds the full
(<= 1.1 MB) weights and a contiguous shard of each coordinate batch; the only data-path collective is ONE
all-reduce (SUM) of the flat fp32 weight-gradient buffer per step (complex gradients as re,im pairs —
complex dtypes are not NCCL types).  The reference has no distributed code; this is the B200-native
counterpart of its chunked loops (wire_occupancy.py:137-154).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional, Sequence, Tuple

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


# ------------------------------------------------------------------------------------------------------------------
# gradient exchange fused with Adam over NVLink peer memory (csrc/peer_kernels.cuh, include/wire_b200.h)
# ------------------------------------------------------------------------------------------------------------------
class _RawCudaBuffer:
    """Exposes a raw device allocation to torch (``torch.as_tensor``) through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def peer_exchange_possible(device: torch.device, group=None):
    """(ok, reason): True when every rank of `group` runs on the same host and this rank's GPU has peer access to every other
    rank's GPU — the precondition of ``PeerGradExchange`` (cudaIpc handles do not cross hosts).  Collective: every rank of
    the group must call it; all ranks get the same answer."""
    import socket
    world = dist.get_world_size(group)
    info = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), int(torch.device(device).index or 0)), group=group)
    rank = dist.get_rank(group)
    ok, why = True, ""
    if len({h for h, _ in info}) != 1:
        ok, why = False, "ranks span several hosts"
    elif world > 16:
        ok, why = False, "more than 16 ranks"
    else:
        me = info[rank][1]
        for r, (_, d) in enumerate(info):
            if r != rank and d != me and not torch.cuda.can_device_access_peer(me, d):
                ok, why = False, f"no peer access between GPU {me} and GPU {d}"
                break
    flags = [None] * world
    dist.all_gather_object(flags, (ok, why), group=group)
    for o, w in flags:
        if not o:
            return False, w
    return True, ""


class PeerGradExchange:
    """One peer-mapped gradient buffer per rank; every rank's Adam kernel sums all of them with P2P loads.

    All ranks of `group` must live on ONE node with NVLink/PCIe peer access between their GPUs (the 8 B200s of a box).
    ``grad`` is this rank's flat fp32 gradient buffer (a torch view of the peer allocation: point the backward pass at
    it); ``bases`` is the ctypes array of every rank's buffer as mapped here, which the C ABI's ``wire_adam_step_peer`` /
    ``wire_peer_wait_done`` take.  The exchange itself involves no torch.distributed call; the group is only used once,
    here, to swap the 64-byte cudaIpc handles."""

    def __init__(self, n_floats: int, device: torch.device, group=None):
        from . import _lib
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise _lib.WireB200Error("PeerGradExchange supports up to 16 ranks (one NVSwitch box)")
        self.device = device
        self.n_floats = (n_floats + 3) // 4 * 4
        handle = ctypes.create_string_buffer(64)
        base = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.wire_peer_alloc(self.n_floats, ctypes.byref(base), handle), "wire_peer_alloc")
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._own = base.value
            self._opened: List[int] = []
            self.bases = (ctypes.c_void_p * self.world)()
            err = None
            for r in range(self.world):
                if r == self.rank:
                    self.bases[r] = self._own
                else:
                    ptr = ctypes.c_void_p()
                    try:
                        _lib.check(self.lib.wire_peer_open(handles[r], ctypes.byref(ptr)), f"wire_peer_open(rank {r})")
                    except _lib.WireB200Error as exc:
                        err = str(exc)
                        break
                    self.bases[r] = ptr.value
                    self._opened.append(ptr.value)
            # every rank must reach the same verdict, or some would wait in the barrier kernels for peers that gave up
            errs: List[Optional[str]] = [None] * self.world
            dist.all_gather_object(errs, err, group=group)
            if any(e is not None for e in errs):
                for ptr in self._opened:
                    self.lib.wire_peer_close(ptr)
                self.lib.wire_peer_free(self._own)
                self._own, self._opened = None, []
                raise _lib.WireB200Error("peer mapping failed: " + next(e for e in errs if e is not None))
        header = int(self.lib.wire_peer_header_bytes())
        self._raw = _RawCudaBuffer(self._own + header, self.n_floats)
        self.grad = torch.as_tensor(self._raw, device=device)
        dist.barrier(group=group)  # every rank has mapped every buffer before the first step touches them

    def close(self) -> None:
        if getattr(self, "_own", None) is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            try:
                dist.barrier(group=self.group)  # nobody may still be reading a buffer that is about to be unmapped
            except Exception:
                pass
            for ptr in self._opened:
                self.lib.wire_peer_close(ptr)
            self.lib.wire_peer_free(self._own)
        self._own, self._opened, self.grad = None, [], None
