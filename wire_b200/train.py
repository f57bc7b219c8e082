"""Fused training step for a WIRE INR (fused loss + flat Adam + CUDA graph).

Replaces the body of the reference's training loops

    pixelvalues = model(b_coords)
    loss = ((pixelvalues - gt[:, b_indices, :])**2).mean()
    optim.zero_grad(); loss.backward(); optim.step()          # wire_image_denoise.py:148-157

with one call, ``loss = trainer.step(coords, target)``:  C-ABI forward → MSE gradient kernel → C-ABI backward
into ONE flat fp32 gradient buffer → (one NCCL all-reduce of that buffer when data-parallel) → one fused Adam
kernel over the flat (re,im) parameter buffer — Adam on ``view_as_real`` parameters, exactly what
``torch.optim.Adam`` does for complex parameters.  The model's ``nn.Parameter``s are re-pointed at views of the
flat buffer, so ``state_dict()``, ``model(coords)`` and the reference drivers keep seeing the trained weights.
With ``graph=True`` (default) the whole step is captured once and replayed as a CUDA graph; the step counter and
the learning rate live on the device (``set_lr`` implements the drivers' ``LambdaLR`` schedules).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from . import functional as F
from ._lib import NetGrads, WireB200Error, check


class Trainer:
    def __init__(self, model, lr: float = 5e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 graph: bool = True, process_group=None, peer_exchange: Optional[bool] = None):
        self.model = model
        self.lib = _lib.load()
        layers = list(model.net)
        if model.hidden_layers < 1 or len(layers) != model.hidden_layers + 2:
            raise WireB200Error("Trainer needs the standard WIRE stack (first layer, >=1 hidden layers, final Linear)")
        if any(layer.scalars_trainable for layer in layers[:-1]) and not model.fused_scalar_grads_ok():
            raise WireB200Error("trainable omega_0 / scale_0 in the fused Trainer need precision='mixed16' (in_features <= 3, "
                                "out_features <= 4); other configurations train through model(coords) + torch.optim")
        self.desc = F.make_desc(model.two_d, model.in_features, model.width, model.hidden_layers, model.out_features,
                                model.precision)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graph = graph
        dev = layers[0].linear.weight.device
        if dev.type != "cuda":
            raise WireB200Error("Trainer needs the model on a CUDA device: wire_b200 has no CPU path")
        self.device = dev

        # ---- flat parameter / gradient / Adam buffers (16-byte aligned slots) ----
        two_d = bool(self.desc.two_d)
        self._train_params: List[torch.nn.Parameter] = []
        self._slots: List[tuple] = []          # (layer index or -1 for the final Linear, field of wire_layer_grads / wire_net_grads)
        for l, layer in enumerate(layers[:-1]):
            self._train_params += [layer.linear.weight, layer.linear.bias]
            self._slots += [(l, "weight"), (l, "bias")]
            if two_d:
                self._train_params += [layer.scale_orth.weight, layer.scale_orth.bias]
                self._slots += [(l, "weight2"), (l, "bias2")]
            # trainable=True (modules/wire.py:66,80-81): the layer's own omega_0 / scale_0 join the flat buffers
            if layer.omega_0.requires_grad:
                self._train_params.append(layer.omega_0)
                self._slots.append((l, "omega0"))
            if layer.scale_0.requires_grad:
                self._train_params.append(layer.scale_0)
                self._slots.append((l, "scale0"))
        self._train_params += [layers[-1].weight, layers[-1].bias]
        self._slots += [(-1, "final_weight"), (-1, "final_bias")]
        if any(p is None for p in self._train_params):
            raise WireB200Error("Trainer requires bias=True layers")
        sizes = [p.numel() * (2 if p.is_complex() else 1) for p in self._train_params]
        self._starts, off = [], 0
        for s in sizes:
            self._starts.append(off)
            off += (s + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        # data-parallel exchange.  peer_exchange=None (default): when every rank of the group sits on ONE host and every pair of
        # their GPUs has peer access (the 8 B200s of an NVSwitch box), the flat gradient lives in a peer-mapped buffer and the
        # Adam kernel sums every rank's gradients itself over NVLink (parallel.PeerGradExchange) — the ranks must then issue
        # their steps in lock-step (a rank that pauses longer than WIRE_B200_PEER_TIMEOUT_S, default 600 s, makes the others
        # trap).  Anything else (several hosts, no peer access, mapping failure, WIRE_B200_PEER=0, peer_exchange=False): one
        # torch.distributed all-reduce of the flat buffer per step.  peer_exchange=True insists and raises if it cannot.
        self.peer = None
        if self.world > 1 and peer_exchange is not False:
            from .parallel import PeerGradExchange, peer_exchange_possible
            want = peer_exchange is True or os.environ.get("WIRE_B200_PEER", "1") != "0"
            ok, why = peer_exchange_possible(dev, process_group) if want else (False, "disabled by WIRE_B200_PEER=0")
            if ok:
                try:
                    self.peer = PeerGradExchange(off, dev, process_group)
                except WireB200Error as exc:   # every rank fails or succeeds together (PeerGradExchange agrees on it)
                    ok, why = False, str(exc)
            if not ok and peer_exchange is True:
                raise WireB200Error(f"peer gradient exchange requested but not possible: {why}")
            self.peer_fallback_reason = None if ok else why
        if self.peer is not None:
            self.flat_grad = self.peer.grad
        else:
            self.flat_grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self._grad_views = []
        with torch.no_grad():
            for p, o, s in zip(self._train_params, self._starts, sizes):
                view = self.flat[o:o + s]
                src = torch.view_as_real(p.data).reshape(-1) if p.is_complex() else p.data.reshape(-1)
                view.copy_(src)
                p.data = torch.view_as_complex(view.view(*p.shape, 2)) if p.is_complex() else view.view(p.shape)
                self._grad_views.append(self.flat_grad[o:o + s])
        if self.world > 1:
            # replicas must start from the same parameters (the peer path keeps them bit-identical from then on)
            dist.broadcast(self.flat, 0, group=process_group)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.scratch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_ring = torch.zeros(256, dtype=torch.float32, device=dev)   # loss of optimiser step s at [s % 256]
        # wire_net_backward_mse (loss gradient computed inside the top backward kernel) is parity-tested but measured equal to
        # the separate 6-us loss kernel (the top kernel's I/O warp gets longer: 1.042 vs 1.042 ms/step), so it is opt-in
        self._fused_mse = os.environ.get("WIRE_B200_FUSED_MSE", "0") == "1"
        self._issued = 0                                                        # optimiser steps issued (host mirror of step_dev)
        # staging pairs for pinned host inputs: the copy of step i + depth - 1 may run while step i computes
        self._stage_depth = max(2, int(os.environ.get("WIRE_B200_STAGE_DEPTH", "2")))
        self._n = None
        self._key = None
        self._n_global = None
        self._states = {}
        self._copy_stream = None
        self._pool = None
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    # ------------------------------------------------------------------------------------------
    def set_lr(self, lr: float) -> None:
        self.lr_dev.fill_(float(lr))

    def set_loss_avgpool(self, H: int, W: int, scale: int) -> None:
        """Super-resolution loss (wire_SISR.py:151-161): ``step(coords_hr, gt_lr)`` then takes the H*W high-resolution
        coordinates and the LOW-resolution target [(H//scale)*(W//scale), out]; the prediction is average-pooled by
        ``scale`` (``torch.nn.AvgPool2d(scale)``) inside the fused loss kernel before the squared error."""
        if H // scale < 1 or W // scale < 1:
            raise WireB200Error("scale larger than the image")
        self._pool = (int(H), int(W), int(scale))
        self._states.clear()
        self._key = None

    @property
    def steps_done(self) -> int:
        return int(self.step_dev.item())

    # per-batch-size state: buffers, workspace, C-ABI pointer tables, captured graph.  Epoch loops alternate between the
    # chunk size and one ragged last chunk (wire_occupancy.py:141), so a few sizes are kept instead of re-allocating.
    _MAX_STATES = 4

    def _prepare(self, n: int, n_global: Optional[int] = None) -> None:
        key = (n, n_global)
        self._n, self._key = n, key
        st = self._states.get(key)
        if st is not None:
            self._states[key] = self._states.pop(key)  # most recently used last
            self.__dict__.update(st)
            return
        while len(self._states) >= self._MAX_STATES:
            self._states.pop(next(iter(self._states)))
        dev, d = self.device, self.desc
        st = {"_graph": None, "_n_global": n_global}
        st["coords_buf"] = torch.empty((n, d.in_features), dtype=torch.float32, device=dev)
        n_target = n if self._pool is None else (self._pool[0] // self._pool[2]) * (self._pool[1] // self._pool[2])
        st["target_buf"] = torch.empty((n_target, d.out_features), dtype=torch.float32, device=dev)
        st["out_buf"] = torch.empty((n, d.out_features), dtype=torch.float32, device=dev)
        st["gout_buf"] = torch.empty((n, d.out_features), dtype=torch.float32, device=dev)
        nbytes = self.lib.wire_net_workspace_bytes(ctypes.byref(d), n, 1)
        if nbytes == 0:
            check(1, "wire_net_workspace_bytes")
        st["ws"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(self.lib.wire_net_workspace_init(ctypes.byref(d), n, 1, st["ws"].data_ptr(), nbytes, F._stream()),
              "wire_net_workspace_init")
        tensors = self.model.flat_params()
        st["_P"] = F._fill_net_params(d, F._check_net_tensors(d, tensors))
        G = NetGrads()
        for (l, field), view in zip(self._slots, self._grad_views):
            setattr(G if l < 0 else G.layer[l], field, view.data_ptr())
        if self.peer is None:
            # the Adam kernel clears every gradient element as it consumes it (zero_grad=1): no memset in the step
            G.clear_mode = _lib.GRADS_PREZEROED
        else:
            # peers read this buffer during their Adam kernels: it is cleared after peer_wait, by ONE memset
            G.clear_mode, G.flat_base, G.flat_floats = _lib.GRADS_CLEAR_FLAT, self.flat_grad.data_ptr(), self.flat_grad.numel()
        st["_G"] = G
        self._states[key] = st
        self.__dict__.update(st)

    def _peer_wait(self) -> None:
        # peers must have finished reading last step's gradients before they are overwritten
        check(self.lib.wire_peer_wait_done(self.peer.bases, self.peer.world, self.peer.rank, self.step_dev.data_ptr(), F._stream()),
              "wire_peer_wait_done")

    def _fwd_bwd(self) -> None:
        d, n, st = self.desc, self._n, F._stream()
        lib = self.lib
        if self._pool is not None:
            self.loss_dev.zero_()
        check(lib.wire_net_forward(ctypes.byref(d), ctypes.byref(self._P), self.coords_buf.data_ptr(), n, self.out_buf.data_ptr(),
                                   self.ws.data_ptr(), self.ws.numel(), 1, st), "wire_net_forward")
        if self._pool is not None:
            Hh, Ww, sc = self._pool
            check(lib.wire_avgpool_mse_loss_grad(self.out_buf.data_ptr(), self.target_buf.data_ptr(), Hh, Ww, d.out_features, sc,
                                                 self.gout_buf.data_ptr(), self.loss_dev.data_ptr(), st), "wire_avgpool_mse_loss_grad")
        if self.peer is not None:
            self._peer_wait()
        if self._pool is None and not self._fused_mse:
            n_norm = n if self._n_global is None else self._n_global
            check(lib.wire_mse_loss_grad_ring(self.out_buf.data_ptr(), self.target_buf.data_ptr(), n * d.out_features,
                                              n_norm * d.out_features, self.gout_buf.data_ptr(), self.loss_ring.data_ptr(),
                                              self.loss_ring.numel(), self.step_dev.data_ptr(), st), "wire_mse_loss_grad_ring")
        elif self._pool is None:
            # MSE fused into the top of the backward pass: grad_out is never materialised on the mixed16 path, and the loss
            # of optimiser step s lands in loss_ring[s % R] (the next slot is cleared) — no loss / reset kernels, and the host
            # can read it for the next R - 1 steps without putting a copy on the compute stream's critical path
            n_norm = n if self._n_global is None else self._n_global
            check(lib.wire_net_backward_mse(ctypes.byref(d), ctypes.byref(self._P), self.coords_buf.data_ptr(), n,
                                            self.out_buf.data_ptr(), self.target_buf.data_ptr(), n_norm * d.out_features,
                                            self.loss_ring.data_ptr(), self.loss_ring.numel(), self.step_dev.data_ptr(),
                                            self.gout_buf.data_ptr(), self.ws.data_ptr(), self.ws.numel(), ctypes.byref(self._G), None, st),
                  "wire_net_backward_mse")
            return
        check(lib.wire_net_backward(ctypes.byref(d), ctypes.byref(self._P), self.coords_buf.data_ptr(), n, self.gout_buf.data_ptr(),
                                    self.ws.data_ptr(), self.ws.numel(), ctypes.byref(self._G), None, st), "wire_net_backward")

    def _adam(self, scale: Optional[float] = None) -> None:
        b1, b2 = self.betas
        # equal shards with a local mean: average the ranks' gradients; shards of a global mean (n_global): they add up
        if scale is None:
            scale = 1.0 if getattr(self, "_n_global", None) is not None else 1.0 / self.world
        if self.peer is not None:
            check(self.lib.wire_adam_step_peer(self.flat.data_ptr(), self.peer.bases, self.peer.world, self.peer.rank,
                                               self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.flat.numel(),
                                               self.lr_dev.data_ptr(), b1, b2, self.eps, self.weight_decay, self.step_dev.data_ptr(),
                                               scale, self.scratch.data_ptr(), F._stream()), "wire_adam_step_peer")
            return
        check(self.lib.wire_adam_step_dev(self.flat.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self.flat.numel(), self.lr_dev.data_ptr(), b1, b2, self.eps,
                                          self.weight_decay, self.step_dev.data_ptr(), scale, self.scratch.data_ptr(), 1,
                                          F._stream()), "wire_adam_step_dev")

    def _exchange_and_adam(self, scale: Optional[float] = None) -> None:
        if self.world > 1 and self.peer is None:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self._adam(scale)

    def _whole_step(self) -> None:
        self._fwd_bwd()
        self._exchange_and_adam()

    def _run(self) -> None:
        """The training step on the current state's device buffers (eager the first time, then a CUDA-graph replay)."""
        if not self.use_graph or (self.world > 1 and self.peer is None):
            self._whole_step()
        elif self._graph is None:
            # one eager step first (lazy one-time setup inside the C ABI must not happen during capture), then the capture
            # (which records the step without executing it): this call performs exactly one training step
            self._whole_step()
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self._whole_step()
            torch.cuda.current_stream().wait_stream(side)
            self._graph = g
            self._states[self._key]["_graph"] = g
        else:
            self._graph.replay()

    def _empty_step(self) -> None:
        """This rank's shard of the batch is empty: contribute zero gradients, but take part in the exchange and in Adam."""
        if self.peer is not None:
            self._peer_wait()
        self.flat_grad.zero_()
        self.loss_dev.zero_()
        R = self.loss_ring.numel()
        self.loss_ring[(self._issued + 1) % R].zero_()   # what the loss kernel of a non-empty step does for the next slot
        self._exchange_and_adam(scale=1.0)   # an empty shard only occurs with n_global batches: the ranks' gradients add up
        self._issued += 1

    def step(self, coords: torch.Tensor, target: torch.Tensor, n_global: Optional[int] = None) -> torch.Tensor:
        """One training iteration on (coords [..., in], target [..., out]); host or device tensors.
        Returns the mean-squared error of this rank's batch as a device scalar (no host sync).

        Data parallel: by default every rank passes an equally sized shard and the ranks' gradients are averaged.
        With ``n_global`` (the coordinate count of the whole batch over all ranks) the loss is normalised by the global
        count instead — shards may then be unequal, and the returned value is this rank's PART of the global mean."""
        d = self.desc
        n = coords.numel() // d.in_features
        if coords.dtype != torch.float32 or target.dtype != torch.float32:
            raise WireB200Error("coords and target must be float32")
        n_target = n
        if self._pool is not None:
            Hh, Ww, sc = self._pool
            if n != Hh * Ww or n_global is not None:
                raise WireB200Error("the pooled loss needs the full H*W coordinate grid on one rank")
            n_target = (Hh // sc) * (Ww // sc)
        if target.numel() != n_target * d.out_features:
            raise WireB200Error("target does not match coords")
        with torch.cuda.device(self.device):
            if n == 0:
                self._empty_step()
                return self._loss_of_last_step()
            if (n, n_global) != self._key:
                self._prepare(n, n_global)
            self._load_inputs(coords.reshape(n, d.in_features), target.reshape(n_target, d.out_features))
            self._run()
            self._issued += 1
        return self._loss_of_last_step()

    def _loss_of_last_step(self) -> torch.Tensor:
        """Device scalar holding the loss of the step just issued: a slot of the loss ring, valid until ``len(loss_ring) - 1``
        further steps have been issued (copy it, e.g. to pinned memory on a side stream, if it is needed for longer)."""
        if self._pool is not None:
            return self.loss_dev[0].clone()   # the pooled-loss kernel accumulates into one scalar that the next step clears
        return self.loss_ring[(self._issued - 1) % self.loss_ring.numel()]

    def _load_inputs(self, coords: torch.Tensor, target: torch.Tensor) -> None:
        """Bring this step's inputs into the step's fixed input buffers.  Device tensors: one D2D copy each.  Pinned host
        tensors (the reference's ``b_coords = coords[...].cuda()`` per chunk, wire_image_denoise.py:145-147): the
        host->device copy runs on a COPY STREAM into one of two staging buffers, so it overlaps the previous step's
        kernels; the compute stream then only does a device-to-device copy (microseconds) before the step's graph."""
        if coords.device.type == "cuda" or not (coords.is_pinned() and target.is_pinned()):
            self.coords_buf.copy_(coords, non_blocking=True)
            self.target_buf.copy_(target, non_blocking=True)
            return
        st = self._states[self._key]
        if "stage" not in st:
            depth = self._stage_depth
            st["stage"] = [(torch.empty_like(self.coords_buf), torch.empty_like(self.target_buf)) for _ in range(depth)]
            st["stage_free"] = [None] * depth   # event: the compute stream has consumed this staging pair
            st["stage_i"] = 0
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        i = st["stage_i"]
        st["stage_i"] = (i + 1) % len(st["stage"])
        sc, stg = st["stage"][i]
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            if st["stage_free"][i] is not None:
                self._copy_stream.wait_event(st["stage_free"][i])
            sc.copy_(coords, non_blocking=True)
            stg.copy_(target, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        cur.wait_event(ready)
        self.coords_buf.copy_(sc, non_blocking=True)
        self.target_buf.copy_(stg, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        st["stage_free"][i] = done

    def step_indexed(self, batcher, idx: Optional[torch.Tensor] = None, start: Optional[int] = None, count: Optional[int] = None,
                     n_global: Optional[int] = None, rec: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training iteration on the grid points ``idx`` (int64 device tensor of linear indices; or the range
        ``start .. start+count``) of a ``wire_b200.data.GridBatcher``: coordinates are generated and targets gathered on the
        device by one kernel, straight into the step's input buffers — the reference's
        ``b_coords = coords[b_indices].cuda(); ... gt[b_indices]`` (wire_occupancy.py:142-149) without host work.
        ``rec`` ([total, out], optional) receives the step's predictions: ``rec[b_indices] = pixelvalues``."""
        d = self.desc
        if batcher.ndim != d.in_features or batcher.out_features != d.out_features:
            raise WireB200Error("GridBatcher does not match the model's in/out features")
        n = idx.numel() if idx is not None else int(count)
        with torch.cuda.device(self.device):
            if n == 0:
                self._empty_step()
                return self._loss_of_last_step()
            if (n, n_global) != self._key:
                self._prepare(n, n_global)
            batcher.assemble_into(self.coords_buf, self.target_buf, idx, start, count)
            self._run()
            self._issued += 1
            if rec is not None:
                batcher.scatter(rec, self.out_buf, idx, start, count)
        return self._loss_of_last_step()

    def step_sisr(self, coords_hr: torch.Tensor, gt_lr: torch.Tensor, gt_hr: Optional[torch.Tensor] = None):
        """One iteration of the super-resolution loop (wire_SISR.py:154-177) after ``set_loss_avgpool(H, W, scale)``:

            rec_hr = model(coords_hr); rec = AvgPool2d(scale)(rec_hr); loss = ((gt_lr - rec)**2).mean()     # :157-161
            with torch.no_grad(): rec_hr = model(coords_hr); mse = ((gt - rec_hr)**2).mean()                 # :163-168
            optim.zero_grad(); loss.backward(); optim.step()                                                 # :173-175

        The reference's second, ``no_grad`` forward sees the same weights and coordinates as the first (the optimiser steps
        after it), so it reproduces the first one bit for bit; here ONE forward serves both: the prediction the training
        forward wrote is returned as ``rec_hr`` (a view of the step's output buffer, valid until the next step) and, when
        ``gt_hr`` is given, its mean squared error against the high-resolution image is computed on the device.
        Returns ``(loss, rec_hr, mse_hr or None)`` — device tensors, no host sync."""
        if self._pool is None:
            raise WireB200Error("call set_loss_avgpool(H, W, scale) first")
        loss = self.step(coords_hr, gt_lr)
        rec_hr = self.out_buf
        mse_hr = None
        if gt_hr is not None:
            from . import data
            g = gt_hr.reshape(rec_hr.shape)
            if g.device != rec_hr.device:
                g = g.to(rec_hr.device, non_blocking=True)
            mse_hr = data.mse(g.contiguous(), rec_hr)
        return loss, rec_hr, mse_hr

    def close(self) -> None:
        """Release the peer-mapped gradient buffer (collective: every rank must call it)."""
        self._states.clear()
        self._graph = None
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    @torch.no_grad()
    def predict(self, coords: torch.Tensor) -> torch.Tensor:
        return self.model(coords)
