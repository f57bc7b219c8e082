"""On-device coordinate pipeline and metrics around the WIRE hot path.

The reference assembles every batch on the host — ``torch.randperm`` on the CPU, a CPU gather of coordinate rows, one
``.cuda()`` copy per chunk (``wire_image_denoise.py:142-147``, ``wire_occupancy.py:137-144``) — scatters the prediction
back with ``rec[:, b_indices] = pixelvalues`` and computes PSNR / IoU with further torch ops (``modules/utils.py:67-82``,
``modules/volutils.py:74-91``).  ``GridBatcher`` keeps the signal resident in HBM and builds a batch from linear indices
with ONE kernel of the C ABI (``wire_grid_batch``): coordinates are *generated* from the index (bit-exact with
``utils.get_coords`` / the drivers' ``torch.linspace`` grids), targets are gathered.  ``run_epoch`` is the reference's
epoch loop on top of ``Trainer.step_indexed``, sharded across ranks by coordinate.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from . import functional as F
from ._lib import WireB200Error, check

LINSPACE_NUMPY = 0   # utils.get_coords: np.linspace in float64, cast to float32 (wire_occupancy.py:70)
LINSPACE_TORCH = 1   # torch.linspace in float32 on the CPU (wire_image_denoise.py:63-66, wire_SISR.py, wire_ct.py)


class GridBatcher:
    """A signal sampled on an (H, W) or (H, W, T) grid, resident on the device.

    ``signal``: float32 ``[H*W(*T), out_features]`` in the reference's flattening (``im.reshape(H*W*T, out)``); linear
    index ``(i*W + j)*T + k`` has coordinate ``(x_j, y_i, z_k)`` — ``np.meshgrid``'s default 'xy' order, as in
    ``utils.get_coords`` (``modules/utils.py:163-176``)."""

    def __init__(self, shape: Sequence[int], signal: Optional[torch.Tensor] = None, linspace: str = "numpy",
                 device: Optional[torch.device] = None):
        if len(shape) not in (2, 3):
            raise WireB200Error("GridBatcher needs an (H, W) or (H, W, T) grid")
        self.lib = _lib.load()
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.total = 1
        for s in self.shape:
            self.total *= s
        self.kind = {"numpy": LINSPACE_NUMPY, "torch": LINSPACE_TORCH}[linspace]
        self._dims = (ctypes.c_int32 * 3)(*(list(self.shape) + [1] * (3 - self.ndim)))
        if signal is not None:
            if signal.dtype != torch.float32 or signal.device.type != "cuda":
                raise WireB200Error("signal must be a float32 CUDA tensor: wire_b200 has no CPU path")
            signal = signal.reshape(self.total, -1).contiguous()
            device = signal.device
        if device is None or torch.device(device).type != "cuda":
            raise WireB200Error("GridBatcher needs a CUDA device")
        self.device = torch.device(device)
        self.signal = signal
        self.out_features = signal.shape[1] if signal is not None else 0
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)

    # -- batch assembly ------------------------------------------------------------------------------------------
    def _idx_args(self, idx, start, count) -> Tuple[int, int, int]:
        if idx is not None:
            if idx.dtype != torch.int64 or idx.device != self.device or not idx.is_contiguous():
                raise WireB200Error("idx must be a contiguous int64 tensor on the batcher's device")
            return idx.data_ptr(), 0, idx.numel()
        if start is None or count is None:
            raise WireB200Error("give idx, or start and count")
        return 0, int(start), int(count)

    def assemble_into(self, coords: Optional[torch.Tensor], target: Optional[torch.Tensor], idx: Optional[torch.Tensor] = None,
                      start: Optional[int] = None, count: Optional[int] = None) -> int:
        """coords[r] = grid point of index idx[r] (or start + r), target[r] = signal[index]; returns the batch size."""
        ptr, base, n = self._idx_args(idx, start, count)
        if target is not None and self.signal is None:
            raise WireB200Error("this GridBatcher has no signal to gather targets from")
        for t, w in ((coords, self.ndim), (target, self.out_features)):
            if t is not None and (t.dtype != torch.float32 or t.device != self.device or not t.is_contiguous() or t.numel() < n * w):
                raise WireB200Error("output buffers must be contiguous float32 tensors on the batcher's device, large enough")
        with torch.cuda.device(self.device):
            check(self.lib.wire_grid_batch(self._dims, self.ndim, self.kind, ptr, base, n,
                                           self.signal.data_ptr() if target is not None else None, self.out_features,
                                           coords.data_ptr() if coords is not None else None,
                                           target.data_ptr() if target is not None else None, self.err.data_ptr(), F._stream()),
                  "wire_grid_batch")
        return n

    def assemble(self, idx: Optional[torch.Tensor] = None, start: Optional[int] = None, count: Optional[int] = None):
        _, _, n = self._idx_args(idx, start, count)
        coords = torch.empty((n, self.ndim), dtype=torch.float32, device=self.device)
        target = torch.empty((n, self.out_features), dtype=torch.float32, device=self.device) if self.signal is not None else None
        self.assemble_into(coords, target, idx, start, count)
        return coords, target

    def coords(self, start: int = 0, count: Optional[int] = None) -> torch.Tensor:
        """The grid's coordinates [count, ndim] (all of them by default) — ``utils.get_coords`` generated on the device."""
        count = self.total - start if count is None else count
        out = torch.empty((count, self.ndim), dtype=torch.float32, device=self.device)
        self.assemble_into(out, None, None, start, count)
        return out

    def scatter(self, rec: torch.Tensor, src: torch.Tensor, idx: Optional[torch.Tensor] = None, start: Optional[int] = None,
                count: Optional[int] = None) -> None:
        """rec[idx] = src  (``rec[:, b_indices, :] = pixelvalues``, wire_image_denoise.py:150-151)."""
        ptr, base, n = self._idx_args(idx, start, count)
        width = src.numel() // max(n, 1)
        if rec.dtype != torch.float32 or src.dtype != torch.float32 or not rec.is_contiguous() or not src.is_contiguous():
            raise WireB200Error("scatter needs contiguous float32 tensors")
        with torch.cuda.device(self.device):
            check(self.lib.wire_scatter_rows(ptr, base, n, src.data_ptr(), width, rec.data_ptr(), rec.numel() // width,
                                             self.err.data_ptr(), F._stream()), "wire_scatter_rows")

    def check_indices(self) -> None:
        """Raises if any batch since the last call saw an out-of-range index (one host sync)."""
        if int(self.err.item()):
            self.err.zero_()
            raise WireB200Error("GridBatcher: linear index outside the grid")


# -- metrics ---------------------------------------------------------------------------------------------------------
def iou_counts(preds: torch.Tensor, gt: torch.Tensor, thres: Optional[float] = None, in_place: bool = True,
               group=None, reduce: bool = False) -> torch.Tensor:
    """``volutils.get_I_and_U`` (modules/volutils.py:79-91) as one pass on the device: returns an int64 tensor
    [intersection, union].  Like the reference, ``preds`` is thresholded IN PLACE unless ``in_place=False``.
    ``reduce=True`` sums the counts over the ranks of ``group`` (each rank passing its own shard)."""
    lib = _lib.load()
    if preds.dtype != torch.float32 or gt.dtype != torch.float32 or preds.device.type != "cuda" or gt.device != preds.device:
        raise WireB200Error("iou_counts needs float32 CUDA tensors on one device")
    if preds.numel() != gt.numel() or not preds.is_contiguous() or not gt.is_contiguous():
        raise WireB200Error("iou_counts: preds and gt must be contiguous and of equal size")
    counts = torch.zeros(2, dtype=torch.int64, device=preds.device)
    with torch.cuda.device(preds.device):
        check(lib.wire_iou_counts(preds.data_ptr(), gt.data_ptr(), preds.numel(), float(thres if thres is not None else 0.0),
                                  int(thres is not None), int(in_place), counts.data_ptr(), F._stream()), "wire_iou_counts")
    if reduce and dist.is_available() and dist.is_initialized():
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def get_IoU(preds: torch.Tensor, gt: torch.Tensor, thres: Optional[float] = None) -> torch.Tensor:
    """``volutils.get_IoU`` (modules/volutils.py:74-76): intersection / union, a device scalar."""
    c = iou_counts(preds, gt, thres)
    return c[0] / c[1]


def sq_err_stats(x: torch.Tensor, xhat: torch.Tensor) -> torch.Tensor:
    """float64 device tensor [sum (x - xhat)^2, max x] in one pass."""
    lib = _lib.load()
    if x.dtype != torch.float32 or xhat.dtype != torch.float32 or x.device.type != "cuda" or xhat.device != x.device:
        raise WireB200Error("sq_err_stats needs float32 CUDA tensors on one device")
    if x.numel() != xhat.numel() or not x.is_contiguous() or not xhat.is_contiguous():
        raise WireB200Error("sq_err_stats: x and xhat must be contiguous and of equal size")
    stats = torch.tensor([0.0, float("-inf")], dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.wire_sq_err_stats(x.data_ptr(), xhat.data_ptr(), x.numel(), stats.data_ptr(), F._stream()), "wire_sq_err_stats")
    return stats


def psnr(x: torch.Tensor, xhat: torch.Tensor) -> torch.Tensor:
    """``utils.psnr`` (modules/utils.py:67-82): 10 log10(max(x) / mean((x - xhat)^2)), a float64 device scalar."""
    s = sq_err_stats(x, xhat)
    return 10.0 * torch.log10(s[1] / (s[0] / x.numel()))


def mse(x: torch.Tensor, xhat: torch.Tensor) -> torch.Tensor:
    """``((gt - rec)**2).mean()`` (wire_image_denoise.py:161-167) as a float64 device scalar."""
    return sq_err_stats(x, xhat)[0] / x.numel()


# -- the reference's epoch loop -----------------------------------------------------------------------------------------
def run_epoch(trainer, batcher: GridBatcher, maxpoints: int, indices: Optional[torch.Tensor] = None,
              generator: Optional[torch.Generator] = None, rec: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One epoch of ``wire_occupancy.py:136-158`` / ``wire_image_denoise.py:141-157``: a random permutation of the grid,
    chunks of ``maxpoints``, one fused training step per chunk; returns the mean chunk loss (device scalar, this rank).

    ``indices``: the epoch's permutation (int64, host or device; e.g. the reference's own ``torch.randperm(H*W*T)`` for a
    trajectory-identical run).  Default: a permutation drawn on the device — every rank must then pass a generator in the
    same state.  Data parallel: each chunk is split into contiguous per-rank shards (``parallel.shard_range``); the loss is
    normalised by the chunk size, so unequal shards are exact.  ``rec`` ([total, out]) receives the predictions made
    during the epoch (``im_estim[b_indices] = pixelvalues``); with several ranks it is summed over ranks at the end."""
    from .parallel import shard_range
    dev = batcher.device
    N = batcher.total
    world, rank = trainer.world, (dist.get_rank(trainer.group) if trainer.world > 1 else 0)
    if indices is None:
        if world > 1 and generator is None:
            raise WireB200Error("data-parallel run_epoch needs the epoch's permutation (indices=) or a generator that is in the "
                                "same state on every rank: each rank would otherwise draw its own permutation")
        indices = torch.randperm(N, device=dev, generator=generator)
    indices = indices.to(dev, non_blocking=True)
    if rec is not None and world > 1:
        rec.zero_()
    total = torch.zeros((), dtype=torch.float32, device=dev)
    nchunks = 0
    for b in range(0, N, maxpoints):
        chunk = indices[b:min(N, b + maxpoints)]
        lo, hi = shard_range(chunk.numel(), rank, world)
        loss = trainer.step_indexed(batcher, chunk[lo:hi], n_global=chunk.numel() if world > 1 else None, rec=rec)
        total += loss
        nchunks += 1
    if rec is not None and world > 1:
        dist.all_reduce(rec, op=dist.ReduceOp.SUM, group=trainer.group)
    return total / max(nchunks, 1)
