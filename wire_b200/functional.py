"""torch.autograd.Function wrappers over the C ABI: the only place where torch tensors become raw pointers.

``wire_net``       — whole stack, replaces ``wire.INR.forward`` (modules/wire.py:161-165) /
                     ``wire2d.INR.forward`` (modules/wire2d.py:121-125) and their autograd graph
``gabor_layer``    — one layer, replaces ``ComplexGaborLayer.forward`` (modules/wire.py:88-93) /
                     ``ComplexGaborLayer2D.forward`` (modules/wire2d.py:56-67)
PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic runs in the
hand-written kernels behind ``libwire_b200.so``.  There is no eager fallback.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LayerGrads, LayerParams, NetDesc, NetGrads, NetParams, WireB200Error, check

_PRECISIONS = {"tf32": _lib.PRECISION_TF32, "fp32": _lib.PRECISION_FP32, "mixed16": _lib.PRECISION_MIXED16}


def precision_id(name: str) -> int:
    try:
        return _PRECISIONS[name]
    except KeyError:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {name!r}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str, dtype) -> torch.Tensor:
    if not t.is_cuda:
        raise WireB200Error(f"{what} must be a CUDA tensor: wire_b200 has no CPU path (got device {t.device})")
    if t.dtype != dtype:
        raise WireB200Error(f"{what} must be {dtype}, got {t.dtype}")
    return t.contiguous()


# --------------------------------------------------------------------------------------------
# workspace pool: scratch for activations is checked out per forward and returned when the
# autograd node dies, so two live graphs never share saved activations.
# --------------------------------------------------------------------------------------------
class _Lease:
    def __init__(self, pool, key, buf):
        self.pool, self.key, self.buf = pool, key, buf

    def release(self):
        if self.buf is not None:
            self.pool._give_back(self.key, self.buf)
            self.buf = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class WorkspacePool:
    def __init__(self):
        self._free: Dict[tuple, List[torch.Tensor]] = {}
        self._lock = threading.Lock()

    def _give_back(self, key, buf):
        with self._lock:
            lst = self._free.setdefault(key, [])
            if len(lst) < 2:
                lst.append(buf)

    def lease_net(self, desc: NetDesc, n: int, training: bool, device) -> _Lease:
        lib = _lib.load()
        key = ("net", desc.two_d, desc.in_features, desc.width, desc.hidden_layers, desc.out_features,
               desc.precision, int(n) if training else min(int(n), int(lib.wire_b200_infer_chunk_rows())), bool(training),
               str(device))
        with self._lock:
            lst = self._free.get(key)
            buf = lst.pop() if lst else None
        if buf is None:
            nbytes = lib.wire_net_workspace_bytes(ctypes.byref(desc), n, int(training))
            if nbytes == 0:
                check(1, "wire_net_workspace_bytes")
            buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            check(lib.wire_net_workspace_init(ctypes.byref(desc), n, int(training), buf.data_ptr(), nbytes, _stream()),
                  "wire_net_workspace_init")
        return _Lease(self, key, buf)

    def clear(self):
        with self._lock:
            self._free.clear()


POOL = WorkspacePool()


def make_desc(two_d: bool, in_features: int, width: int, hidden_layers: int, out_features: int, precision: str) -> NetDesc:
    return NetDesc(int(bool(two_d)), int(in_features), int(width), int(hidden_layers), int(out_features),
                   precision_id(precision))


# --------------------------------------------------------------------------------------------
# whole network
# --------------------------------------------------------------------------------------------
# flat parameter order handed to WireNetFn.apply (per layer): weight, bias, [weight2, bias2], omega_0, scale_0 ;
# then final weight, final bias
def _per_layer(two_d: bool) -> int:
    return 6 if two_d else 4


def _fill_net_params(desc: NetDesc, tensors: Sequence[torch.Tensor]) -> NetParams:
    two_d = bool(desc.two_d)
    per = _per_layer(two_d)
    n_layers = desc.hidden_layers + 1
    if len(tensors) != per * n_layers + 2:
        raise WireB200Error(f"expected {per * n_layers + 2} parameter tensors, got {len(tensors)}")
    P = NetParams()
    for l in range(n_layers):
        t = tensors[l * per:(l + 1) * per]
        lp = P.layer[l]
        lp.weight, lp.bias = _ptr(t[0]), _ptr(t[1])
        if two_d:
            lp.weight2, lp.bias2, lp.omega0, lp.scale0 = _ptr(t[2]), _ptr(t[3]), _ptr(t[4]), _ptr(t[5])
        else:
            lp.weight2, lp.bias2, lp.omega0, lp.scale0 = None, None, _ptr(t[2]), _ptr(t[3])
    P.final_weight, P.final_bias = _ptr(tensors[-2]), _ptr(tensors[-1])
    return P


def _check_net_tensors(desc: NetDesc, tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    two_d = bool(desc.two_d)
    per = _per_layer(two_d)
    M, K0 = desc.width, desc.in_features
    out: List[torch.Tensor] = []
    for l in range(desc.hidden_layers + 1):
        t = tensors[l * per:(l + 1) * per]
        wd = torch.float32 if l == 0 else torch.complex64
        wshape = (M, K0) if l == 0 else (M, M)
        n_w = 4 if two_d else 2
        for i in range(n_w):
            x = _require_cuda(t[i], f"layer {l} parameter {i}", wd)
            exp = wshape if i % 2 == 0 else (M,)
            if tuple(x.shape) != exp:
                raise WireB200Error(f"layer {l} parameter {i}: expected shape {exp}, got {tuple(x.shape)}")
            out.append(x)
        for i in range(n_w, n_w + 2):
            out.append(_require_cuda(t[i], f"layer {l} omega_0/scale_0", torch.float32))
    fw = _require_cuda(tensors[-2], "final weight", torch.complex64)
    fb = _require_cuda(tensors[-1], "final bias", torch.complex64)
    if tuple(fw.shape) != (desc.out_features, M) or tuple(fb.shape) != (desc.out_features,):
        raise WireB200Error("final layer shape mismatch")
    out += [fw, fb]
    return out


class WireNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc: NetDesc, coords: torch.Tensor, *params: torch.Tensor):
        lib = _lib.load()
        coords_c = _require_cuda(coords, "coords", torch.float32)
        if coords_c.shape[-1] != desc.in_features:
            raise WireB200Error(f"coords last dim {coords_c.shape[-1]} != in_features {desc.in_features}")
        tensors = _check_net_tensors(desc, params)
        flat = coords_c.reshape(-1, desc.in_features)
        n = flat.shape[0]
        training = any(ctx.needs_input_grad[1:])
        out = torch.empty((n, desc.out_features), dtype=torch.float32, device=flat.device)
        P = _fill_net_params(desc, tensors)
        with torch.cuda.device(flat.device):
            lease = POOL.lease_net(desc, n, training, flat.device)
            check(lib.wire_net_forward(ctypes.byref(desc), ctypes.byref(P), flat.data_ptr(), n, out.data_ptr(),
                                       lease.buf.data_ptr(), lease.buf.numel(), int(training), _stream()),
                  "wire_net_forward")
        if training:
            ctx.desc, ctx.lease, ctx.n = desc, lease, n
            ctx.coords_shape = coords.shape
            ctx.save_for_backward(flat, *tensors)
        else:
            lease.release()
        return out.reshape(*coords.shape[:-1], desc.out_features)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        lib = _lib.load()
        desc, n = ctx.desc, ctx.n
        flat, *tensors = ctx.saved_tensors
        if ctx.lease.buf is None:
            raise WireB200Error("workspace of this forward pass was already released (backward called twice "
                                "without retain_graph?)")
        g = _require_cuda(grad_out, "grad_out", torch.float32).reshape(n, desc.out_features)
        two_d = bool(desc.two_d)
        per = _per_layer(two_d)
        n_w = 4 if two_d else 2
        # one flat fp32 buffer for every gradient (what a data-parallel all-reduce wants to see)
        sizes = []
        for l in range(desc.hidden_layers + 1):
            for i in range(n_w):
                t = tensors[l * per + i]
                sizes.append(t.numel() * (2 if t.is_complex() else 1))
        sizes += [tensors[-2].numel() * 2, tensors[-1].numel() * 2]
        # trainable omega_0 / scale_0 (modules/wire.py:66,80-81): one float slot each, filled by the fused backward kernels
        scal_slots = {}
        for l in range(desc.hidden_layers + 1):
            for j in range(2):
                if ctx.needs_input_grad[2 + l * per + n_w + j]:
                    scal_slots[(l, j)] = len(sizes)
                    sizes.append(1)
        # every slot starts on a 16-byte boundary (view_as_complex needs an even offset; kernels like float4)
        starts, off = [], 0
        for s in sizes:
            starts.append(off)
            off += (s + 3) // 4 * 4
        flatg = torch.empty(off, dtype=torch.float32, device=flat.device)
        views = [flatg[o:o + s] for o, s in zip(starts, sizes)]
        G = NetGrads()
        vi = 0
        grads: List[Optional[torch.Tensor]] = []
        for l in range(desc.hidden_layers + 1):
            lg = G.layer[l]
            ptrs = []
            for i in range(n_w):
                t = tensors[l * per + i]
                v = views[vi]
                vi += 1
                ptrs.append(v.data_ptr())
                grads.append(torch.view_as_complex(v.view(*t.shape, 2)) if t.is_complex() else v.view(t.shape))
            lg.weight, lg.bias = ptrs[0], ptrs[1]
            if two_d:
                lg.weight2, lg.bias2 = ptrs[2], ptrs[3]
            for j in range(2):   # omega_0, scale_0: None unless trainable (non-trainable in every reference driver)
                k = scal_slots.get((l, j))
                if k is None:
                    grads.append(None)
                else:
                    setattr(lg, "omega0" if j == 0 else "scale0", views[k].data_ptr())
                    grads.append(views[k].view(tensors[l * per + n_w + j].shape))
        G.final_weight, G.final_bias = views[vi].data_ptr(), views[vi + 1].data_ptr()
        G.clear_mode, G.flat_base, G.flat_floats = _lib.GRADS_CLEAR_FLAT, flatg.data_ptr(), flatg.numel()   # one memset
        grads.append(torch.view_as_complex(views[vi].view(*tensors[-2].shape, 2)))
        grads.append(torch.view_as_complex(views[vi + 1].view(*tensors[-1].shape, 2)))
        g_coords = None
        if ctx.needs_input_grad[1]:
            g_coords = torch.empty_like(flat)
        P = _fill_net_params(desc, tensors)
        with torch.cuda.device(flat.device):
            check(lib.wire_net_backward(ctypes.byref(desc), ctypes.byref(P), flat.data_ptr(), n, g.data_ptr(),
                                        ctx.lease.buf.data_ptr(), ctx.lease.buf.numel(), ctypes.byref(G),
                                        _ptr(g_coords), _stream()),
                  "wire_net_backward")
        if g_coords is not None:
            g_coords = g_coords.reshape(ctx.coords_shape)
        return (None, g_coords, *grads)


_WS_TENSORS = {"y": 0, "z": 1, "w": 2, "gz": 3, "gw": 4, "gz0": 5, "gw0": 6}


def workspace_read(desc: NetDesc, n: int, workspace: torch.Tensor, which: str, index: int = 0) -> torch.Tensor:
    """One tensor of a training workspace as dense float32 (C ABI ``wire_net_workspace_read``): ``which`` in
    y / z / w / gz / gw / gz0 / gw0.  Complex tensors come back as complex64 ``[n, M]``, the first layer's real gradients
    as float32 ``[n, M]``.  ``workspace`` is the buffer a ``wire_net_forward(training=1)`` (+ ``wire_net_backward``) ran on —
    for the autograd route, ``out.grad_fn.lease.buf`` of the tensor ``model(coords)`` returned."""
    lib = _lib.load()
    real = which in ("gz0", "gw0")
    out = torch.empty((n, desc.width if real else 2 * desc.width), dtype=torch.float32, device=workspace.device)
    with torch.cuda.device(workspace.device):
        check(lib.wire_net_workspace_read(ctypes.byref(desc), n, workspace.data_ptr(), workspace.numel(), _WS_TENSORS[which],
                                          int(index), out.data_ptr(), _stream()), "wire_net_workspace_read")
    return out if real else torch.view_as_complex(out.view(n, desc.width, 2))


def wire_net(desc: NetDesc, coords: torch.Tensor, params: Sequence[torch.Tensor]) -> torch.Tensor:
    return WireNetFn.apply(desc, coords, *params)


# --------------------------------------------------------------------------------------------
# single layer
# --------------------------------------------------------------------------------------------
def _layer_ws(desc: NetDesc, is_first: bool, in_features: int, n: int, device) -> torch.Tensor:
    lib = _lib.load()
    nbytes = lib.wire_gabor_layer_workspace_bytes(ctypes.byref(desc), int(is_first), in_features, n)
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _layer_params(two_d, weight, bias, weight2, bias2, omega0, scale0) -> LayerParams:
    lp = LayerParams()
    lp.weight, lp.bias = _ptr(weight), _ptr(bias)
    lp.weight2, lp.bias2 = (_ptr(weight2), _ptr(bias2)) if two_d else (None, None)
    lp.omega0, lp.scale0 = _ptr(omega0), _ptr(scale0)
    return lp


class GaborLayerFn(torch.autograd.Function):
    """y = gabor(x W^T + b) for one layer; x real [.., K] (first layer) or complex64 [.., K]."""

    @staticmethod
    def forward(ctx, desc: NetDesc, is_first: bool, x, weight, bias, weight2, bias2, omega0, scale0):
        lib = _lib.load()
        two_d = bool(desc.two_d)
        xd = torch.float32 if is_first else torch.complex64
        xc = _require_cuda(x, "layer input", xd)
        K = xc.shape[-1]
        M = desc.width
        weight = _require_cuda(weight, "weight", xd)
        bias = _require_cuda(bias, "bias", xd)
        if two_d:
            weight2 = _require_cuda(weight2, "scale_orth.weight", xd)
            bias2 = _require_cuda(bias2, "scale_orth.bias", xd)
        omega0 = _require_cuda(omega0, "omega_0", torch.float32)
        scale0 = _require_cuda(scale0, "scale_0", torch.float32)
        if tuple(weight.shape) != (M, K):
            raise WireB200Error(f"weight shape {tuple(weight.shape)} != {(M, K)}")
        flat = xc.reshape(-1, K)
        n = flat.shape[0]
        training = any(ctx.needs_input_grad)
        y = torch.empty((n, M), dtype=torch.complex64, device=flat.device)
        z = w = None
        if training:
            z = torch.empty((n, M), dtype=xd, device=flat.device)
            if two_d:
                w = torch.empty((n, M), dtype=xd, device=flat.device)
        ws = _layer_ws(desc, is_first, K, n, flat.device)
        lp = _layer_params(two_d, weight, bias, weight2, bias2, omega0, scale0)
        with torch.cuda.device(flat.device):
            check(lib.wire_gabor_layer_forward(ctypes.byref(desc), int(is_first), K, ctypes.byref(lp), flat.data_ptr(), n,
                                               y.data_ptr(), _ptr(z), _ptr(w), ws.data_ptr(), ws.numel(), _stream()),
                  "wire_gabor_layer_forward")
        if training:
            ctx.desc, ctx.is_first, ctx.K, ctx.n, ctx.x_shape = desc, is_first, K, n, x.shape
            ctx.save_for_backward(flat, z, w if two_d else flat.new_empty(0), weight, bias,
                                  weight2 if two_d else flat.new_empty(0), bias2 if two_d else flat.new_empty(0),
                                  omega0, scale0)
        return y.reshape(*x.shape[:-1], M)

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        desc, is_first, K, n = ctx.desc, ctx.is_first, ctx.K, ctx.n
        two_d = bool(desc.two_d)
        flat, z, w, weight, bias, weight2, bias2, omega0, scale0 = ctx.saved_tensors
        gy = _require_cuda(grad_y, "grad_y", torch.complex64).reshape(n, desc.width)
        gW, gb = torch.empty_like(weight), torch.empty_like(bias)
        gW2 = gb2 = None
        if two_d:
            gW2, gb2 = torch.empty_like(weight2), torch.empty_like(bias2)
        gx = torch.empty_like(flat) if ctx.needs_input_grad[2] else None
        lg = LayerGrads()
        lg.weight, lg.bias, lg.weight2, lg.bias2 = _ptr(gW), _ptr(gb), _ptr(gW2), _ptr(gb2)
        lp = _layer_params(two_d, weight, bias, weight2, bias2, omega0, scale0)
        ws = _layer_ws(desc, is_first, K, n, flat.device)
        with torch.cuda.device(flat.device):
            check(lib.wire_gabor_layer_backward(ctypes.byref(desc), int(is_first), K, ctypes.byref(lp), flat.data_ptr(),
                                                z.data_ptr(), _ptr(w) if two_d else None, gy.data_ptr(), n, _ptr(gx),
                                                ctypes.byref(lg), ws.data_ptr(), ws.numel(), _stream()),
                  "wire_gabor_layer_backward")
        if gx is not None:
            gx = gx.reshape(ctx.x_shape)
        g_om = g_s0 = None
        if ctx.needs_input_grad[7] or ctx.needs_input_grad[8]:   # trainable omega_0 / scale_0 (modules/wire.py:80-81)
            acc = torch.zeros(2, dtype=torch.float64, device=flat.device)
            zr = torch.view_as_real(z) if z.is_complex() else z
            wr = (torch.view_as_real(w) if w.is_complex() else w) if two_d else None
            with torch.cuda.device(flat.device):
                check(lib.wire_gabor_scalar_grads(int(is_first), int(two_d), desc.width, zr.data_ptr(), _ptr(wr),
                                                  torch.view_as_real(gy).data_ptr(), n, omega0.data_ptr(), scale0.data_ptr(),
                                                  acc.data_ptr(), _stream()), "wire_gabor_scalar_grads")
            acc = acc.float()
            g_om = acc[0:1] if ctx.needs_input_grad[7] else None
            g_s0 = acc[1:2] if ctx.needs_input_grad[8] else None
        return (None, None, gx, gW, gb, gW2, gb2, g_om, g_s0)


def gabor_layer(desc, is_first, x, weight, bias, weight2, bias2, omega0, scale0):
    return GaborLayerFn.apply(desc, is_first, x, weight, bias, weight2, bias2, omega0, scale0)


class FinalLinearRealFn(torch.autograd.Function):
    """Re(h W_f^T + b_f) with autograd (C ABI ``wire_final_linear_forward`` / ``_backward``): the output layer of the
    layer-by-layer route (``modules/wire.py:156-165``), used when the fused whole-network kernels do not apply
    (trainable omega_0 / scale_0, user-edited stacks)."""

    @staticmethod
    def forward(ctx, desc: NetDesc, h, weight, bias):
        out = final_linear_real(desc, h, weight, bias)
        ctx.desc, ctx.h_shape = desc, h.shape
        ctx.save_for_backward(h, weight)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        desc = ctx.desc
        h, weight = ctx.saved_tensors
        hc = h.contiguous().reshape(-1, h.shape[-1])
        n = hc.shape[0]
        go = _require_cuda(grad_out, "grad_out", torch.float32).reshape(n, desc.out_features)
        weight = weight.contiguous()
        gh = torch.empty_like(hc) if ctx.needs_input_grad[1] else None
        gW = torch.empty_like(weight)
        gb = torch.empty(desc.out_features, dtype=torch.complex64, device=hc.device)
        with torch.cuda.device(hc.device):
            check(lib.wire_final_linear_backward(ctypes.byref(desc), weight.data_ptr(), hc.data_ptr(), go.data_ptr(), n, _ptr(gh),
                                                 gW.data_ptr(), gb.data_ptr(), _stream()), "wire_final_linear_backward")
        if gh is not None:
            gh = gh.reshape(ctx.h_shape)
        return (None, gh, gW, gb)


def final_linear_real_autograd(desc: NetDesc, h, weight, bias):
    return FinalLinearRealFn.apply(desc, h, weight, bias)


def final_linear_real(desc: NetDesc, h: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Re(h W^T + b) with the CUDA kernel (no autograd: see ``final_linear_real_autograd``)."""
    lib = _lib.load()
    hc = _require_cuda(h, "h", torch.complex64)
    flat = hc.reshape(-1, hc.shape[-1])
    out = torch.empty((flat.shape[0], desc.out_features), dtype=torch.float32, device=flat.device)
    weight = _require_cuda(weight, "final weight", torch.complex64)
    bias = _require_cuda(bias, "final bias", torch.complex64)
    with torch.cuda.device(flat.device):
        check(lib.wire_final_linear_forward(ctypes.byref(desc), weight.data_ptr(), bias.data_ptr(), flat.data_ptr(),
                                            flat.shape[0], out.data_ptr(), _stream()), "wire_final_linear_forward")
    return out.reshape(*h.shape[:-1], desc.out_features)


# --------------------------------------------------------------------------------------------
# RealGaborLayer activation (modules/wire.py:38-42)
# --------------------------------------------------------------------------------------------
class RealGaborFn(torch.autograd.Function):
    """y = cos(omega_0 f) * exp(-(scale_0 s)^2) — one fused kernel each way (C ABI ``wire_real_gabor_forward`` / ``_backward``)."""

    @staticmethod
    def forward(ctx, f, s, omega0: float, scale0: float):
        lib = _lib.load()
        fc = _require_cuda(f, "freqs output", torch.float32)
        sc = _require_cuda(s, "scale output", torch.float32)
        if fc.shape != sc.shape:
            raise WireB200Error("freqs and scale outputs must have the same shape")
        y = torch.empty_like(fc)
        with torch.cuda.device(fc.device):
            check(lib.wire_real_gabor_forward(fc.data_ptr(), sc.data_ptr(), fc.numel(), float(omega0), float(scale0), y.data_ptr(),
                                              _stream()), "wire_real_gabor_forward")
        ctx.save_for_backward(fc, sc)
        ctx.consts = (float(omega0), float(scale0))
        return y

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        fc, sc = ctx.saved_tensors
        gy = _require_cuda(grad_y, "grad_y", torch.float32)
        gf, gs = torch.empty_like(fc), torch.empty_like(sc)
        with torch.cuda.device(fc.device):
            check(lib.wire_real_gabor_backward(fc.data_ptr(), sc.data_ptr(), gy.data_ptr(), fc.numel(), ctx.consts[0], ctx.consts[1],
                                               gf.data_ptr(), gs.data_ptr(), _stream()), "wire_real_gabor_backward")
        return gf, gs, None, None


def real_gabor(f, s, omega0, scale0):
    return RealGaborFn.apply(f, s, omega0, scale0)


class RealGaborLayerFn(torch.autograd.Function):
    """The whole ``RealGaborLayer.forward`` (modules/wire.py:29-42): both real Linears and the activation in the kernels of this
    repo (C ABI ``wire_real_gabor_layer_forward`` / ``_backward``, FP32 FMAs) — no library GEMM."""

    @staticmethod
    def forward(ctx, x, w_freqs, b_freqs, w_scale, b_scale, omega0: float, scale0: float):
        lib = _lib.load()
        xc = _require_cuda(x, "layer input", torch.float32)
        wf = _require_cuda(w_freqs, "freqs.weight", torch.float32)
        ws = _require_cuda(w_scale, "scale.weight", torch.float32)
        bf = None if b_freqs is None else _require_cuda(b_freqs, "freqs.bias", torch.float32)
        bs = None if b_scale is None else _require_cuda(b_scale, "scale.bias", torch.float32)
        M, K = wf.shape
        if tuple(ws.shape) != (M, K) or xc.shape[-1] != K:
            raise WireB200Error(f"RealGaborLayer shapes do not match: x {tuple(xc.shape)}, freqs {tuple(wf.shape)}, scale {tuple(ws.shape)}")
        flat = xc.reshape(-1, K)
        n = flat.shape[0]
        training = any(ctx.needs_input_grad)
        y = torch.empty((n, M), dtype=torch.float32, device=flat.device)
        f = torch.empty_like(y) if training else None
        sv = torch.empty_like(y) if training else None
        with torch.cuda.device(flat.device):
            check(lib.wire_real_gabor_layer_forward(flat.data_ptr(), n, K, M, wf.data_ptr(), _ptr(bf), ws.data_ptr(), _ptr(bs),
                                                    float(omega0), float(scale0), y.data_ptr(), _ptr(f), _ptr(sv), _stream()),
                  "wire_real_gabor_layer_forward")
        if training:
            ctx.save_for_backward(flat, f, sv, wf, ws)
            ctx.meta = (n, K, M, float(omega0), float(scale0), x.shape, b_freqs is not None, b_scale is not None)
        return y.reshape(*x.shape[:-1], M)

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        flat, f, sv, wf, ws = ctx.saved_tensors
        n, K, M, omega0, scale0, x_shape, has_bf, has_bs = ctx.meta
        gy = _require_cuda(grad_y, "grad_y", torch.float32).reshape(n, M)
        gx = torch.empty_like(flat) if ctx.needs_input_grad[0] else None
        gwf, gws = torch.empty_like(wf), torch.empty_like(ws)
        gbf = torch.empty(M, dtype=torch.float32, device=flat.device)
        gbs = torch.empty(M, dtype=torch.float32, device=flat.device)
        sf, ss = torch.empty_like(f), torch.empty_like(sv)
        with torch.cuda.device(flat.device):
            check(lib.wire_real_gabor_layer_backward(flat.data_ptr(), f.data_ptr(), sv.data_ptr(), gy.data_ptr(), n, K, M, wf.data_ptr(),
                                                     ws.data_ptr(), omega0, scale0, _ptr(gx), gwf.data_ptr(), gbf.data_ptr(),
                                                     gws.data_ptr(), gbs.data_ptr(), sf.data_ptr(), ss.data_ptr(), _stream()),
                  "wire_real_gabor_layer_backward")
        if gx is not None:
            gx = gx.reshape(x_shape)
        return gx, gwf, (gbf if has_bf else None), gws, (gbs if has_bs else None), None, None


def real_gabor_layer(x, w_freqs, b_freqs, w_scale, b_scale, omega0, scale0):
    return RealGaborLayerFn.apply(x, w_freqs, b_freqs, w_scale, b_scale, omega0, scale0)
