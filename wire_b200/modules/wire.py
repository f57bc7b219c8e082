"""WIRE with the 1-D complex Gabor wavelet — CUDA-backed counterpart of the reference's ``modules/wire.py``.

Same constructor signatures, attribute names, parameter names/dtypes/shapes and ``state_dict`` keys as
``ComplexGaborLayer`` (modules/wire.py:44-93) and ``INR`` (modules/wire.py:94-167), so
``model.net[i]``, ``model.net[0].omega_0``, ``state_dict()``/``load_state_dict()``, ``torch.optim.Adam``
and ``count_parameters`` keep working.  ``INR.forward`` runs the whole stack through the fused
sm_100a kernels (``wire_b200.functional.wire_net``); a single layer called on its own
(``model.net[idx](x)``, modules/utils.py:251-252) runs ``wire_b200.functional.gabor_layer``.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import functional as F

# 'mixed16' (default): FP16 activations / BF16 gradients in HBM, FP32 accumulate — the fastest path, parity-tested to the
# same bars as 'tf32' (FP16 carries TF32's 11-bit significand); 'tf32': 32-bit operands; 'fp32': CUDA-core yardstick.
DEFAULT_PRECISION = "mixed16"


class _GaborBase(nn.Module):
    """Shared plumbing of ComplexGaborLayer / ComplexGaborLayer2D."""

    two_d = False

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega0=10.0, sigma0=40.0,
                 trainable=False, precision=DEFAULT_PRECISION):
        super().__init__()
        self.is_first = is_first
        self.in_features = in_features
        self.out_features = out_features
        self.precision = precision
        dtype = torch.float if is_first else torch.cfloat
        # modules/wire.py:80-81: f32[1] Parameters that show up in parameters()/state_dict()
        self.omega_0 = nn.Parameter(omega0 * torch.ones(1), trainable)
        self.scale_0 = nn.Parameter(sigma0 * torch.ones(1), trainable)
        self.linear = nn.Linear(in_features, out_features, bias=bias, dtype=dtype)
        if self.two_d:
            self.scale_orth = nn.Linear(in_features, out_features, bias=bias, dtype=dtype)

    def _bias(self, lin: nn.Linear) -> torch.Tensor:
        if lin.bias is not None:
            return lin.bias
        return torch.zeros(self.out_features, dtype=lin.weight.dtype, device=lin.weight.device)

    @property
    def scalars_trainable(self) -> bool:
        return bool(self.omega_0.requires_grad or self.scale_0.requires_grad)


    def flat_params(self):
        """Parameter tensors in the order ``functional.wire_net`` expects."""
        ps = [self.linear.weight, self._bias(self.linear)]
        if self.two_d:
            ps += [self.scale_orth.weight, self._bias(self.scale_orth)]
        return ps + [self.omega_0, self.scale_0]

    def forward(self, input):
        desc = F.make_desc(self.two_d, self.in_features, self.out_features, 1, 1, self.precision)
        so = self.scale_orth if self.two_d else None
        return F.gabor_layer(desc, self.is_first, input, self.linear.weight, self._bias(self.linear),
                             so.weight if so is not None else None, self._bias(so) if so is not None else None,
                             self.omega_0, self.scale_0)


class ComplexGaborLayer(_GaborBase):
    """exp(1j*omega_0*lin - |scale_0*lin|^2) after a Linear — modules/wire.py:44-93."""

    two_d = False


class RealGaborLayer(nn.Module):
    """modules/wire.py:6-42 — real Gabor layer ``cos(omega_0 freqs(x)) * exp(-(scale_0 scale(x))^2)`` (not used by ``INR``).
    Same constructor and parameter names (``freqs``, ``scale``; ``omega_0`` / ``scale_0`` are plain floats, as in the
    reference).  ``freqs`` / ``scale`` are ``nn.Linear`` modules only as parameter containers (state_dict keys): the two
    Linears, the activation and the whole backward pass run in this repo's FP32 kernels (``functional.real_gabor_layer``)."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega0=10.0, sigma0=10.0, trainable=False):
        super().__init__()
        self.omega_0 = omega0
        self.scale_0 = sigma0
        self.is_first = is_first
        self.in_features = in_features
        self.freqs = nn.Linear(in_features, out_features, bias=bias)
        self.scale = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input):
        if not input.is_cuda:
            raise F.WireB200Error("RealGaborLayer input must be a CUDA tensor: wire_b200 has no CPU path")
        return F.real_gabor_layer(input, self.freqs.weight, self.freqs.bias, self.scale.weight, self.scale.bias,
                                  self.omega_0, self.scale_0)


class FinalLinear(nn.Linear):
    """The complex output Linear (modules/wire.py:156).  Inside ``INR.forward`` it is fused into the last
    Gabor kernel; called on its own (layer-by-layer evaluation) it returns the complex output like
    ``nn.Linear`` does, computed by the CUDA kernel when no gradient is required."""

    def forward(self, input):
        desc = F.make_desc(False, 1, self.in_features, 1, self.out_features, DEFAULT_PRECISION)
        bias = self.bias if self.bias is not None else torch.zeros(self.out_features, dtype=self.weight.dtype, device=self.weight.device)
        # Re and Im of h W^T + b as two real-output passes of the CUDA kernel: Im(h W^T + b) = Re(h (-jW)^T + (-jb)).  Under
        # autograd each pass is a FinalLinearRealFn node (the [n, M]-sized work stays in the hand-written kernels; only the
        # multiplication of the [out, M] weight by -j is a torch op) — there is no eager nn.Linear path.
        if torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad):
            re = F.final_linear_real_autograd(desc, input, self.weight, bias)
            im = F.final_linear_real_autograd(desc, input, self.weight * (-1j), bias * (-1j))
        else:
            re = F.final_linear_real(desc, input, self.weight, bias)
            im = F.final_linear_real(desc, input, self.weight * (-1j), bias * (-1j))
        return torch.complex(re, im)


class _INRBase(nn.Module):
    layer_cls = ComplexGaborLayer
    two_d = False

    def _build(self, in_features, width, hidden_layers, out_features, first_omega_0, hidden_omega_0, scale, precision):
        self.complex = True
        self.wavelet = 'gabor'
        self.pos_encode = False  # legacy attribute read by modules/utils.py:246
        # |y| = exp(-omega Im z - s^2 |z|^2) peaks at exp(omega^2 / (4 s^2)); mixed16 stores y in FP16 (max 65504 = e^11.09), so
        # hyper-parameters with omega_0 / scale_0 > ~6 (none of the reference's drivers: 7/6, 8/9, 20/10, 3/4) would overflow
        # where the complex64 reference stays finite: such models run the 32-bit-operand kernels instead.
        if precision == "mixed16":
            ratio = max(abs(float(first_omega_0)), abs(float(hidden_omega_0))) / max(abs(float(scale)), 1e-12)
            if ratio * ratio / 4.0 > 9.0:
                import warnings
                warnings.warn(f"wire_b200: omega_0 / scale_0 = {ratio:.2f} lets |y| reach e^{ratio * ratio / 4:.1f}, beyond the FP16 "
                              "activations of precision='mixed16'; using precision='tf32' for this model")
                precision = "tf32"
        self.precision = precision
        self.in_features, self.width = in_features, width
        self.hidden_layers, self.out_features = hidden_layers, out_features
        net = [self.layer_cls(in_features, width, omega0=first_omega_0, sigma0=scale, is_first=True,
                              trainable=False, precision=precision)]
        for _ in range(hidden_layers):
            net.append(self.layer_cls(width, width, omega0=hidden_omega_0, sigma0=scale, precision=precision))
        net.append(FinalLinear(width, out_features, dtype=torch.cfloat))
        self.net = nn.Sequential(*net)

    def flat_params(self):
        ps = []
        for layer in list(self.net)[:-1]:
            ps += layer.flat_params()
        final = self.net[-1]
        return ps + [final.weight, final.bias]

    def fused_scalar_grads_ok(self) -> bool:
        """Trainable omega_0 / scale_0 (modules/wire.py:66,80-81) are differentiated inside the fused backward kernels of the
        mixed16 path (extra epilogue reductions); the other precisions and shapes outside that path use the per-layer route."""
        return (self.precision == "mixed16" and self.in_features <= 3 and self.out_features <= 4 and self.width <= 992)

    def forward(self, coords):
        layers = list(self.net)
        trainable_scalars = any(getattr(layer, "scalars_trainable", False) for layer in layers[:-1])
        if (self.hidden_layers < 1 or len(layers) != self.hidden_layers + 2
                or (trainable_scalars and torch.is_grad_enabled() and not self.fused_scalar_grads_ok())):
            # no hidden layer, a user-edited Sequential, or trainable omega_0 / scale_0 outside the mixed16 path: compose the
            # per-layer kernels, each an autograd node of its own
            x = coords
            for layer in layers[:-1]:
                x = layer(x)
            desc = F.make_desc(self.two_d, self.in_features, self.width, 1, self.out_features, self.precision)
            if torch.is_grad_enabled() and (x.requires_grad or layers[-1].weight.requires_grad):
                return F.final_linear_real_autograd(desc, x, layers[-1].weight, layers[-1].bias)
            return F.final_linear_real(desc, x, layers[-1].weight, layers[-1].bias)
        desc = F.make_desc(self.two_d, self.in_features, self.width, self.hidden_layers, self.out_features,
                           self.precision)
        return F.wire_net(desc, coords, self.flat_params())


class INR(_INRBase):
    """modules/wire.py:94-167 — same positional signature (incl. the fork's ``scaled_hidden_features``)."""

    def __init__(self, in_features, hidden_features, scaled_hidden_features=None, hidden_layers=1, out_features=1,
                 outermost_linear=True, first_omega_0=30, hidden_omega_0=30., scale=10.0, scale_tensor=[],
                 pos_encode=False, multi_scale=False, sidelength=512, fn_samples=None, use_nyquist=True,
                 precision=DEFAULT_PRECISION):
        super().__init__()
        self.nonlin = ComplexGaborLayer
        # "Since complex numbers are two real numbers, reduce the number of hidden parameters by 2" (:119)
        width = int(hidden_features / np.sqrt(2))
        self._build(in_features, width, hidden_layers, out_features, first_omega_0, hidden_omega_0, scale, precision)
