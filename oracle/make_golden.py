"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (test infrastructure; build container only).

Imports ``/root/reference/modules/{wire,wire2d}.py`` unmodified, loads deterministic weights, runs the
forward and ``loss.backward()`` in complex64 and complex128, and commits small fixtures to
``tests/golden/*.npz``.  ``/root/reference`` does not exist on the GPU box: tests only read the fixtures.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from wire_oracle import deterministic_state  # noqa: E402  (shared with the tests)

REF = os.environ.get("WIRE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")

# name, kind, in, hidden_features, H, out, first_omega, hidden_omega, scale, N, coordinate box
CASES = [
    ("wire_small", "wire", 2, 24, 2, 3, 7.0, 7.0, 6.0, 96, 1.0),
    ("wire_occ_small", "wire", 3, 40, 3, 1, 20.0, 20.0, 10.0, 80, 1.0),
    ("wire2d_small", "wire2d", 2, 32, 2, 3, 8.0, 8.0, 9.0, 72, 1.0),
    ("wire_odd_width", "wire", 2, 27, 1, 2, 5.0, 6.0, 4.0, 33, 1.0),      # M = 19 (odd), N ragged
    ("wire_denoise_212", "wire", 2, 300, 2, 3, 7.0, 7.0, 6.0, 64, 1.0),    # the headline width M = 212
    ("wire2d_sisr_128", "wire2d", 2, 256, 2, 3, 8.0, 8.0, 9.0, 48, 1.0),   # SISR width M = 128
]


def build_reference(kind, in_f, hidden, H, out_f, w0, w0h, s0):
    sys.path.insert(0, REF)
    from modules import wire as ref_wire, wire2d as ref_wire2d  # the reference, unmodified
    if kind == "wire":
        return ref_wire.INR(in_f, hidden, None, H, out_f, True, w0, w0h, s0)
    return ref_wire2d.INR(in_f, hidden, H, out_f, True, w0, w0h, s0)


def run(model, coords, grad_out, double):
    if double:
        for p in model.parameters():
            p.data = p.data.to(torch.complex128 if p.is_complex() else torch.float64)
        coords, grad_out = coords.double(), grad_out.double()
    coords = coords.clone().requires_grad_(True)
    layer_out = []
    x = coords
    for layer in model.net:
        x = layer(x)
        layer_out.append(x)
    out = x.real
    (out * grad_out).sum().backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return out.detach(), [t.detach() for t in layer_out], grads, coords.grad.detach()


def main():
    os.makedirs(OUT, exist_ok=True)
    for ci, (name, kind, in_f, hidden, H, out_f, w0, w0h, s0, N, box) in enumerate(CASES):
        rs = np.random.RandomState(1000 + ci)
        coords = torch.from_numpy(rs.uniform(-box, box, size=(1, N, in_f)).astype(np.float32))
        grad_out = torch.from_numpy(rs.normal(size=(1, N, out_f)).astype(np.float32))
        blob = {"coords": coords.numpy(), "grad_out": grad_out.numpy(),
                "meta": np.array([in_f, hidden, H, out_f, N], dtype=np.int64),
                "hyper": np.array([w0, w0h, s0], dtype=np.float64), "kind": np.array(kind)}
        big = name.endswith("_212") or name.endswith("_128")
        for tag, double in (("c64", False), ("c128", True)):
            model = build_reference(kind, in_f, hidden, H, out_f, w0, w0h, s0)
            state = deterministic_state(model, seed=7 + ci)
            model.load_state_dict(state, strict=True)
            out, layer_out, grads, gcoords = run(model, coords, grad_out, double)
            blob[f"out_{tag}"] = out.numpy()
            blob[f"gcoords_{tag}"] = gcoords.numpy()
            if not big:
                for li, t in enumerate(layer_out):
                    blob[f"layer{li}_{tag}"] = t.numpy()
            for k, g in grads.items():
                g = g.numpy()
                if big and g.size > 8192:  # keep fixtures small: a fixed random subsample of the big matrices
                    idx = np.random.RandomState(99).choice(g.size, size=4096, replace=False)
                    blob[f"gradidx.{k}"] = idx.astype(np.int64)
                    g = g.reshape(-1)[idx]
                blob[f"grad_{tag}.{k}"] = g
            if tag == "c64" and not big:
                for k, v in state.items():
                    blob[f"param.{k}"] = v.numpy()
        blob["seed"] = np.array(7 + ci)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  out[:3]={blob['out_c64'].reshape(-1)[:3]}")


if __name__ == "__main__":
    main()
