"""Model factory — CUDA-backed counterpart of the reference's ``modules/models.py:27-77``.

Keeps the fork's parameter names and order.  Two defects of the fork's boundary are absorbed
compatibly: ``scaled_hidden_features`` gets a default (none of the five ``wire_*.py``
drivers pass it), and dispatch is per module by keyword (the fork forwards 14 positionals that
``wire2d.INR`` cannot accept, which makes ``nonlin='wire2d'`` unreachable there).
Only the WIRE family is in scope; the comparison baselines (siren, gauss, relu, mfn, bspline_*) are not.
"""
from __future__ import annotations

from . import wire, wire2d

model_dict = {'wire': wire, 'wire2d': wire2d}


def get_INR(nonlin, in_features, hidden_features, scaled_hidden_features=None, hidden_layers=None,
            out_features=None, outermost_linear=True, first_omega_0=30, hidden_omega_0=30, scale=10,
            scale_tensor=[], pos_encode=False, sidelength=512, fn_samples=None, use_nyquist=True, **extra):
    """Same arguments as the reference's ``get_INR``; returns an ``nn.Module`` whose ``forward(coords)``
    maps f32 ``[..., in_features]`` to f32 ``[..., out_features]``.  ``extra`` may carry ``precision``
    ('mixed16' default, 'tf32', 'fp32')."""
    if nonlin not in model_dict:
        raise ValueError(f"nonlin={nonlin!r} is outside the WIRE hot path served by wire_b200 "
                         f"(supported: {sorted(model_dict)})")
    if hidden_layers is None or out_features is None:
        raise TypeError("get_INR() missing required argument: 'hidden_layers' and 'out_features'")
    if nonlin == 'wire':
        return wire.INR(in_features, hidden_features, scaled_hidden_features, hidden_layers, out_features,
                        outermost_linear, first_omega_0, hidden_omega_0, scale, scale_tensor, pos_encode,
                        False, sidelength, fn_samples, use_nyquist, **extra)
    return wire2d.INR(in_features, hidden_features, hidden_layers, out_features, outermost_linear,
                      first_omega_0, hidden_omega_0, scale, pos_encode, sidelength, fn_samples, use_nyquist,
                      **extra)
