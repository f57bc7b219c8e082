"""WIRE with the 2-D complex Gabor wavelet — CUDA-backed counterpart of the reference's ``modules/wire2d.py``.

``ComplexGaborLayer2D`` (modules/wire2d.py:6-67) adds a second, orthogonal scale Linear ``scale_orth``
on the same input: ``exp(1j*omega_0*lin) * exp(-scale_0^2 (|lin|^2 + |orth|^2))``.  ``INR``
(modules/wire2d.py:70-127) uses width ``int(hidden/2)`` and has no ``scaled_hidden_features`` argument.
"""
from __future__ import annotations

import torch

from .wire import DEFAULT_PRECISION, FinalLinear, _GaborBase, _INRBase  # noqa: F401


class ComplexGaborLayer2D(_GaborBase):
    two_d = True

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega0=10.0, sigma0=10.0,
                 trainable=False, precision=DEFAULT_PRECISION):
        super().__init__(in_features, out_features, bias=bias, is_first=is_first, omega0=omega0, sigma0=sigma0,
                         trainable=trainable, precision=precision)


class INR(_INRBase):
    layer_cls = ComplexGaborLayer2D
    two_d = True

    def __init__(self, in_features, hidden_features, hidden_layers, out_features, outermost_linear=True,
                 first_omega_0=10, hidden_omega_0=10., scale=10.0, pos_encode=False, sidelength=512,
                 fn_samples=None, use_nyquist=True, precision=DEFAULT_PRECISION):
        super().__init__()
        self.nonlin = ComplexGaborLayer2D
        # "reduce the number of hidden parameters by 4" (modules/wire2d.py:92)
        width = int(hidden_features / 2)
        self._build(in_features, width, hidden_layers, out_features, first_omega_0, hidden_omega_0, scale, precision)
