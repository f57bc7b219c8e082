"""Run the reference's drivers unchanged on top of wire_b200.

``patch_reference(modules_pkg)`` rebinds the reference's own entry points for the hot path —
``modules.wire.INR`` / ``ComplexGaborLayer`` (modules/wire.py:44-167), ``modules.wire2d.INR`` /
``ComplexGaborLayer2D`` (modules/wire2d.py:6-127), ``modules.models.get_INR`` (modules/models.py:27-77) and
``modules.lin_inverse.radon`` (modules/lin_inverse.py:19-40)
— to the CUDA-backed implementations, so `wire_image_denoise.py`, `wire_SISR.py`, `wire_occupancy.py`,
`wire_ct.py` and `wire_multi_sr.py` keep their source untouched.
"""
from __future__ import annotations

import importlib
import sys


def patch_reference(modules_pkg=None):
    from .modules import models as m_models, wire as m_wire, wire2d as m_wire2d

    if modules_pkg is None:
        modules_pkg = sys.modules.get("modules") or importlib.import_module("modules")
    name = modules_pkg.__name__
    patched = []
    for sub, ours, attrs in ((f"{name}.wire", m_wire, ("INR", "ComplexGaborLayer")),
                             (f"{name}.wire2d", m_wire2d, ("INR", "ComplexGaborLayer2D"))):
        try:
            mod = sys.modules.get(sub) or importlib.import_module(sub)
        except Exception:
            continue
        for a in attrs:
            setattr(mod, a, getattr(ours, a))
            patched.append(f"{sub}.{a}")
    try:
        ref_models = sys.modules.get(f"{name}.models") or importlib.import_module(f"{name}.models")
    except Exception:
        ref_models = None
    if ref_models is not None:
        ref_get = getattr(ref_models, "get_INR", None)

        def get_INR(nonlin, *args, **kwargs):
            if nonlin in ("wire", "wire2d"):
                return m_models.get_INR(nonlin, *args, **kwargs)
            return ref_get(nonlin, *args, **kwargs)  # out-of-scope nonlinearities stay with the reference

        ref_models.get_INR = get_INR
        if hasattr(ref_models, "model_dict"):
            ref_models.model_dict["wire"] = m_wire
            ref_models.model_dict["wire2d"] = m_wire2d
        patched.append(f"{name}.models.get_INR")
    # the CT driver's forward operator (modules/lin_inverse.py:19-40); the reference module itself needs kornia at import
    try:
        ref_lin = sys.modules.get(f"{name}.lin_inverse") or importlib.import_module(f"{name}.lin_inverse")
    except Exception:
        ref_lin = None
    if ref_lin is not None:
        from . import lin_inverse as m_lin
        ref_lin.radon = m_lin.radon
        patched.append(f"{name}.lin_inverse.radon")
    return patched
