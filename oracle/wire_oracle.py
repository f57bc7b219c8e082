"""CPU oracle for the WIRE hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file.  The product (``wire_b200``) never does; it fails loudly when its CUDA
library is missing.

What it restates (file:line relative to the reference checkout, Annatk26/wire):

* ``modules/wire.py:44-93``   ComplexGaborLayer  — Linear (real if first, else cfloat) then
  ``exp(1j*omega_0*lin - |scale_0*lin|^2)``
* ``modules/wire.py:94-167``  wire.INR           — width ``int(hidden/sqrt(2))``, first layer real,
  ``hidden_layers`` complex Gabor layers, complex final Linear, ``.real`` of the output
* ``modules/wire2d.py:6-67``  ComplexGaborLayer2D — second Linear ``scale_orth`` on the same input,
  ``exp(1j*omega_0*lin) * exp(-scale_0^2 (|lin|^2+|orth|^2))``
* ``modules/wire2d.py:70-127`` wire2d.INR        — width ``int(hidden/2)``

The arithmetic of the reference lives in PyTorch (pinned ``torch==1.13.1`` in the reference's
``requirements.txt:5``; this image has 2.11) — ``nn.Linear`` on complex64 plus element-wise ATen ops
and complex autograd — so the faithful restatement is the same op sequence on CPU tensors
(``TorchOracle`` below, used as the timed "port" CPU baseline), plus an autograd-free closed form in
NumPy float64 (``forward_np`` / ``backward_np``) that pins the Wirtinger
convention independently of autograd.

Parity pinning: the reference ships NO tests/golden vectors for this path.  The
oracle is therefore pinned against outputs of the reference itself, generated in the build
container by ``oracle/make_golden.py`` (imports ``/root/reference/modules/{wire,wire2d}.py``) and
committed under ``tests/golden/``; ``tests/test_oracle.py`` checks both restatements against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn


def hidden_width(kind: str, hidden_features: int) -> int:
    """modules/wire.py:119 (``int(h/sqrt 2)``) and modules/wire2d.py:92 (``int(h/2)``)."""
    return int(hidden_features / np.sqrt(2)) if kind == "wire" else int(hidden_features / 2)


class GaborLayer(nn.Module):
    """One WIRE layer (wire.py:44-93; with ``two_d`` wire2d.py:6-67). Same parameter names."""

    def __init__(self, in_features, out_features, is_first, omega0, sigma0, two_d, cdtype):
        super().__init__()
        rdtype = torch.float32 if cdtype == torch.complex64 else torch.float64
        self.two_d = two_d
        self.omega_0 = nn.Parameter(omega0 * torch.ones(1, dtype=rdtype), requires_grad=False)
        self.scale_0 = nn.Parameter(sigma0 * torch.ones(1, dtype=rdtype), requires_grad=False)
        self.linear = nn.Linear(in_features, out_features, dtype=rdtype if is_first else cdtype)
        if two_d:
            self.scale_orth = nn.Linear(in_features, out_features, dtype=rdtype if is_first else cdtype)

    def forward(self, x):
        lin = self.linear(x)
        if not self.two_d:
            omega = self.omega_0 * lin
            scale = self.scale_0 * lin
            return torch.exp(1j * omega - scale.abs().square())
        orth = self.scale_orth(x)
        freq = torch.exp(1j * self.omega_0 * lin)
        arg = lin.abs().square() + orth.abs().square()
        return freq * torch.exp(-self.scale_0 * self.scale_0 * arg)


class TorchOracle(nn.Module):
    """wire.INR / wire2d.INR restated (state_dict keys identical to the reference's)."""

    def __init__(self, kind, in_features, hidden_features, hidden_layers, out_features,
                 first_omega_0=30.0, hidden_omega_0=30.0, scale=10.0, cdtype=torch.complex64):
        super().__init__()
        assert kind in ("wire", "wire2d")
        width = hidden_width(kind, hidden_features)
        two_d = kind == "wire2d"
        layers: List[nn.Module] = [GaborLayer(in_features, width, True, first_omega_0, scale, two_d, cdtype)]
        for _ in range(hidden_layers):
            layers.append(GaborLayer(width, width, False, hidden_omega_0, scale, two_d, cdtype))
        layers.append(nn.Linear(width, out_features, dtype=cdtype))
        self.net = nn.Sequential(*layers)
        self.kind, self.width = kind, width

    def forward(self, coords):
        return self.net(coords).real


def deterministic_state(model: nn.Module, seed: int):
    """Weights independent of torch's RNG/version: U(+-1/sqrt(fan_in)) from numpy, re and im independently
    (the same distribution nn.Linear's default init produces)."""
    rs = np.random.RandomState(seed)
    state = {}
    for k, v in model.state_dict().items():
        if k.endswith("omega_0") or k.endswith("scale_0"):
            state[k] = v.clone()
            continue
        fan_in = v.shape[1] if v.dim() == 2 else None
        if fan_in is None:  # bias: fan_in of the matching weight
            fan_in = model.state_dict()[k.replace("bias", "weight")].shape[1]
        b = 1.0 / np.sqrt(fan_in)
        re = rs.uniform(-b, b, size=tuple(v.shape)).astype(np.float32)
        if v.is_complex():
            im = rs.uniform(-b, b, size=tuple(v.shape)).astype(np.float32)
            state[k] = torch.from_numpy(re + 1j * im).to(torch.complex64)
        else:
            state[k] = torch.from_numpy(re)
    return state



# --------------------------------------------------------------------------------------------
# closed form (NumPy, float64/complex128) — SURVEY.md appendix A
# --------------------------------------------------------------------------------------------
def _layers_from_state(state: Dict[str, np.ndarray]):
    idx = sorted({int(k.split(".")[1]) for k in state if k.startswith("net.")})
    last = idx[-1]
    layers = []
    for i in idx[:-1]:
        p = f"net.{i}."
        layers.append(dict(
            W=np.asarray(state[p + "linear.weight"]), b=np.asarray(state[p + "linear.bias"]),
            W2=np.asarray(state[p + "scale_orth.weight"]) if p + "scale_orth.weight" in state else None,
            b2=np.asarray(state[p + "scale_orth.bias"]) if p + "scale_orth.bias" in state else None,
            omega=float(np.asarray(state[p + "omega_0"]).reshape(-1)[0]),
            scale=float(np.asarray(state[p + "scale_0"]).reshape(-1)[0])))
    final = dict(W=np.asarray(state[f"net.{last}.weight"]), b=np.asarray(state[f"net.{last}.bias"]))
    return layers, final


def forward_np(state: Dict[str, np.ndarray], coords: np.ndarray, return_saved: bool = False):
    """Forward of the whole stack in complex128. A.1: z = x W^T + b (plain transpose),
    y = exp(j w0 z - s0^2 (|z|^2 [+ |w|^2])), out = Re(h Wf^T + bf)."""
    layers, final = _layers_from_state(state)
    x = np.asarray(coords, dtype=np.float64).reshape(-1, coords.shape[-1])
    saved = []
    for L in layers:
        z = x @ L["W"].astype(np.complex128 if np.iscomplexobj(L["W"]) else np.float64).T + L["b"]
        w = None
        mag = np.abs(z) ** 2
        if L["W2"] is not None:
            w = x @ L["W2"].T + L["b2"]
            mag = mag + np.abs(w) ** 2
        y = np.exp(1j * L["omega"] * z - L["scale"] ** 2 * mag)
        saved.append(dict(x=x, z=z, w=w, y=y))
        x = y
    o = x @ final["W"].astype(np.complex128).T + final["b"]
    out = o.real.reshape(*coords.shape[:-1], -1)
    return (out, saved) if return_saved else out


def backward_np(state: Dict[str, np.ndarray], coords: np.ndarray, grad_out: np.ndarray):
    """Closed-form gradients in PyTorch's convention (grad = dL/dRe + j dL/dIm). A.2:
    p = conj(y) g_y ; g_z = -j w0 p - 2 s0^2 z Re p ; g_w = -2 s0^2 w Re p ;
    first layer (real): g_z = w0 Im p - 2 s0^2 z Re p ;
    Linear: g_x = g_z conj(W), g_W = g_z^T conj(x), g_b = sum_n g_z."""
    layers, final = _layers_from_state(state)
    _, saved = forward_np(state, coords, return_saved=True)
    g_o = np.asarray(grad_out, dtype=np.float64).reshape(-1, grad_out.shape[-1])
    idx_final = len(layers)
    grads: Dict[str, np.ndarray] = {}
    h = saved[-1]["y"]
    grads[f"net.{idx_final}.weight"] = g_o.T.astype(np.complex128) @ np.conj(h)
    grads[f"net.{idx_final}.bias"] = g_o.sum(0).astype(np.complex128)
    g_y = g_o.astype(np.complex128) @ np.conj(final["W"].astype(np.complex128))
    for i in range(len(layers) - 1, -1, -1):
        L, S = layers[i], saved[i]
        p = np.conj(S["y"]) * g_y
        s2 = L["scale"] ** 2
        first = not np.iscomplexobj(L["W"])
        if first:
            g_z = L["omega"] * p.imag - 2.0 * s2 * S["z"] * p.real
        else:
            g_z = -1j * L["omega"] * p - 2.0 * s2 * S["z"] * p.real
        x = S["x"]
        grads[f"net.{i}.linear.weight"] = g_z.T @ np.conj(x)
        grads[f"net.{i}.linear.bias"] = g_z.sum(0)
        g_x = g_z @ np.conj(L["W"])
        if L["W2"] is not None:
            g_w = -2.0 * s2 * S["w"] * p.real
            grads[f"net.{i}.scale_orth.weight"] = g_w.T @ np.conj(x)
            grads[f"net.{i}.scale_orth.bias"] = g_w.sum(0)
            g_x = g_x + g_w @ np.conj(L["W2"])
        g_y = g_x
    grads["coords"] = np.real(g_y).reshape(coords.shape)
    return grads


# --------------------------------------------------------------------------------------------
# single stages of the closed form (the per-kernel parity tests feed each CUDA kernel's own inputs through these)
# --------------------------------------------------------------------------------------------
def gabor_np(z, w, omega, scale):
    """modules/wire.py:91-93 / modules/wire2d.py:62-67 on given pre-activations: exp(j w0 z - s0^2 (|z|^2 [+ |w|^2]))."""
    mag = np.abs(z) ** 2 + (0.0 if w is None else np.abs(w) ** 2)
    return np.exp(1j * omega * np.asarray(z, dtype=np.complex128) - scale ** 2 * mag)


def layer_forward_np(L, x):
    """One layer of ``forward_np`` from a given input x (L: an entry of ``_layers_from_state``): returns z, w (None for wire), y."""
    cplx = np.iscomplexobj(L["W"])
    x = np.asarray(x, dtype=np.complex128 if cplx else np.float64)
    z = x @ L["W"].astype(np.complex128 if cplx else np.float64).T + L["b"]
    w = None if L["W2"] is None else x @ L["W2"].astype(np.complex128 if cplx else np.float64).T + L["b2"]
    return z, w, gabor_np(z, w, L["omega"], L["scale"])


def gabor_backward_np(L, z, w, g_y):
    """A.2 on given pre-activations: p = conj(y) g_y; hidden: g_z = -j w0 p - 2 s0^2 z Re p; first layer (real z):
    g_z = w0 Im p - 2 s0^2 z Re p; wire2d: g_w = -2 s0^2 w Re p.  Returns g_z, g_w (None for wire)."""
    y = gabor_np(z, w, L["omega"], L["scale"])
    p = np.conj(y) * g_y
    s2 = L["scale"] ** 2
    if np.iscomplexobj(L["W"]):
        g_z = -1j * L["omega"] * p - 2.0 * s2 * z * p.real
    else:
        g_z = L["omega"] * p.imag - 2.0 * s2 * np.real(z) * p.real
    g_w = None if w is None else -2.0 * s2 * w * p.real
    return g_z, g_w


def linear_backward_np(W, x, g_z):
    """Complex (or real) Linear backward, PyTorch convention: g_x = g_z conj(W), g_W = g_z^T conj(x), g_b = sum_n g_z."""
    return g_z @ np.conj(W), g_z.T @ np.conj(x), g_z.sum(0)


# --------------------------------------------------------------------------------------------
# metrics the parity harness needs (restated, not imported)
# --------------------------------------------------------------------------------------------
def psnr(x: np.ndarray, xhat: np.ndarray) -> float:
    """modules/utils.py:67-82: 10*log10(max(x) / mse) (note: max, not max^2)."""
    denom = float(np.mean((np.asarray(x) - np.asarray(xhat)) ** 2))
    return 10.0 * math.log10(float(np.max(x)) / denom)


def iou(a: np.ndarray, b: np.ndarray, thres: Optional[float] = 0.5) -> float:
    """modules/volutils.py:74-91: threshold both, |A∧B| / |A∨B|."""
    a = np.asarray(a) > thres
    b = np.asarray(b) > thres
    return float(np.logical_and(a, b).sum()) / float(np.logical_or(a, b).sum())


def image_coords(H: int, W: int) -> torch.Tensor:
    """wire_image_denoise.py:63-66: linspace(-1,1) x linspace(-1,1), meshgrid 'xy', hstack -> [1,HW,2]."""
    x = torch.linspace(-1, 1, W)
    y = torch.linspace(-1, 1, H)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    return torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None, ...]


def volume_coords(H: int, W: int, T: int) -> torch.Tensor:
    """modules/utils.py:163-176 get_coords: np.meshgrid default 'xy' order over three axes."""
    X, Y, Z = np.meshgrid(np.linspace(-1, 1, W), np.linspace(-1, 1, H), np.linspace(-1, 1, T))
    c = np.hstack((X.reshape(-1, 1), Y.reshape(-1, 1), Z.reshape(-1, 1)))
    return torch.tensor(c.astype(np.float32))


# ======================================================================================================================
# Data pipeline either side of the hot path — restated for the parity tests of
# wire_b200/csrc/data_kernels.cuh.  Pinned by tests/golden/data_pipeline.npz (oracle/make_golden_data.py runs the
# reference's own modules/utils.py and modules/volutils.py).
# ======================================================================================================================
def get_coords_np(H: int, W: int, T: Optional[int] = None) -> np.ndarray:
    """modules/utils.py:163-176 ``get_coords``: np.meshgrid (default 'xy') of float64 linspaces, hstack, cast to f32."""
    if T is None:
        X, Y = np.meshgrid(np.linspace(-1, 1, W), np.linspace(-1, 1, H))
        coords = np.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))
    else:
        X, Y, Z = np.meshgrid(np.linspace(-1, 1, W), np.linspace(-1, 1, H), np.linspace(-1, 1, T))
        coords = np.hstack((X.reshape(-1, 1), Y.reshape(-1, 1), Z.reshape(-1, 1)))
    return coords.astype(np.float32)


def image_coords_torch(H: int, W: int) -> np.ndarray:
    """wire_image_denoise.py:63-66 (also wire_SISR.py, wire_ct.py): torch float32 CPU linspace, meshgrid 'xy', hstack."""
    x = torch.linspace(-1, 1, W)
    y = torch.linspace(-1, 1, H)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    return torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1))).numpy()


def iou_counts_np(preds: np.ndarray, gt: np.ndarray, thres: Optional[float] = None):
    """modules/volutils.py:79-91 ``get_I_and_U``: thresholds ``preds`` IN PLACE (two masked assignments, in this order),
    then counts logical_and / logical_or.  Returns (intersection, union) as Python ints."""
    if thres is not None:
        preds[preds < thres] = 0.0
        preds[preds >= thres] = 1.0
    return int(np.logical_and(preds, gt).sum()), int(np.logical_or(preds, gt).sum())


def psnr_np(x: np.ndarray, xhat: np.ndarray) -> float:
    """modules/utils.py:67-82 ``psnr``: 10 log10(max(x) / mean((x - xhat)^2))."""
    err = x - xhat
    return float(10 * np.log10(np.max(x) / np.mean(pow(err, 2))))


def gabor_scalar_grads_np(x, weight, bias, weight2, bias2, omega0, s0, gy):
    """Closed-form gradients of a layer's own ``omega_0`` / ``scale_0`` (``trainable=True``, modules/wire.py:66,80-81 and
    modules/wire2d.py:27,42-43) in float64: with z = x W^T + b, w = x W2^T + b2 (wire2d), y = exp(j w0 z - s0^2 (|z|^2 + |w|^2))
    and p = conj(y) g_y:   g_omega0 = sum Im(conj(z) p),   g_scale0 = -2 s0 sum (|z|^2 + |w|^2) Re(p).
    Pinned by tests/golden/trainable_scalars.npz (the reference's layers under autograd).  Returns (y, g_omega0, g_scale0)."""
    x = np.asarray(x).astype(np.complex128)
    z = x @ np.asarray(weight).astype(np.complex128).T + np.asarray(bias).astype(np.complex128)
    t = np.abs(z) ** 2
    if weight2 is not None:
        w = x @ np.asarray(weight2).astype(np.complex128).T + np.asarray(bias2).astype(np.complex128)
        t = t + np.abs(w) ** 2
    y = np.exp(1j * omega0 * z - s0 * s0 * t)
    p = np.conj(y) * np.asarray(gy).astype(np.complex128)
    return y, float(np.sum((np.conj(z) * p).imag)), float(-2.0 * s0 * np.sum(t * p.real))


def real_gabor_np(x, w_freqs, b_freqs, w_scale, b_scale, omega0, s0):
    """modules/wire.py:38-42 ``RealGaborLayer.forward`` in float64: cos(omega_0 freqs(x)) * exp(-(scale_0 scale(x))^2)."""
    x = np.asarray(x, dtype=np.float64)
    f = x @ np.asarray(w_freqs, dtype=np.float64).T + np.asarray(b_freqs, dtype=np.float64)
    s = x @ np.asarray(w_scale, dtype=np.float64).T + np.asarray(b_scale, dtype=np.float64)
    return np.cos(omega0 * f) * np.exp(-((s0 * s) ** 2))


# ======================================================================================================================
# Radon forward operator of the CT driver — PARITY UNPINNED: the reference computes it with
# ``kornia.geometry.rotate`` (modules/lin_inverse.py:19-40), kornia (pinned 0.6.9 in the reference's requirements) is neither
# under /root/reference nor installed in the build image, and the reference ships no sinogram fixture.  Restated from kornia's
# published source: rotate(tensor, angle) = warp_affine with get_rotation_matrix2d(center=((W-1)/2, (H-1)/2), angle [deg],
# scale 1) = [[a, b, (1-a) cx - b cy], [-b, a, b cx + (1-a) cy]], a = cos, b = sin; warp_affine inverts the matrix in
# normalised coordinates and samples with F.affine_grid / F.grid_sample(mode='bilinear', padding_mode='zeros',
# align_corners=True).  So out[i][j] = bilinear(im, c + R^T ((j, i) - c)), R^T = [[cos, -sin], [sin, cos]].
# ======================================================================================================================
def rotate_torch(imten: torch.Tensor, angles_deg: torch.Tensor) -> torch.Tensor:
    """kornia.geometry.rotate(imten [B,C,H,W], angles [B]) restated with affine_grid / grid_sample (autograd-capable)."""
    B, C, H, W = imten.shape
    t = angles_deg.to(imten.dtype) * (math.pi / 180.0)
    cs, sn = torch.cos(t), torch.sin(t)
    ry = (H - 1) / max(W - 1, 1)
    rx = (W - 1) / max(H - 1, 1)
    zero = torch.zeros_like(cs)
    theta = torch.stack([torch.stack([cs, -sn * ry, zero], -1), torch.stack([sn * rx, cs, zero], -1)], 1)   # [B, 2, 3]
    grid = torch.nn.functional.affine_grid(theta, [B, C, H, W], align_corners=True)
    return torch.nn.functional.grid_sample(imten, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def radon_torch(imten: torch.Tensor, angles_deg: torch.Tensor, is_3d: bool = False) -> torch.Tensor:
    """modules/lin_inverse.py:19-40 line for line, with ``rotate_torch`` in the place of ``kornia.geometry.rotate``."""
    nangles = len(angles_deg)
    imten_rep = torch.repeat_interleave(imten, nangles, 0)
    imten_rot = rotate_torch(imten_rep, angles_deg)
    if is_3d:
        return imten_rot.sum(2).squeeze().permute(1, 0, 2)
    return imten_rot.sum(2).squeeze()


def radon_np(im: np.ndarray, angles_deg) -> np.ndarray:
    """Direct float64 loops over the same formula (small cases): im [H][W] -> sinogram [nangles][W]."""
    H, W = im.shape
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    out = np.zeros((len(angles_deg), W))
    for a, deg in enumerate(angles_deg):
        t = np.deg2rad(float(deg))
        c, s = np.cos(t), np.sin(t)
        for j in range(W):
            acc = 0.0
            for i in range(H):
                xs = cx + c * (j - cx) - s * (i - cy)
                ys = cy + s * (j - cx) + c * (i - cy)
                x0, y0 = int(np.floor(xs)), int(np.floor(ys))
                ax, ay = xs - x0, ys - y0
                for yy, wy in ((y0, 1 - ay), (y0 + 1, ay)):
                    for xx, wx in ((x0, 1 - ax), (x0 + 1, ax)):
                        if 0 <= xx < W and 0 <= yy < H:
                            acc += wx * wy * im[yy, xx]
            out[a, j] = acc
    return out
