"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import os

import numpy as np
import torch

import wire_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    """Network fixtures (tests/golden/wire*.npz); data_pipeline.npz is the fixture of the coordinate pipeline."""
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "wire*.npz")))


def load_data_golden():
    return np.load(os.path.join(GOLDEN_DIR, "data_pipeline.npz"), allow_pickle=False)


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    in_f, hidden, H, out_f, N = (int(v) for v in g["meta"])
    w0, w0h, s0 = (float(v) for v in g["hyper"])
    return dict(g=g, kind=str(g["kind"]), in_f=in_f, hidden=hidden, H=H, out_f=out_f, N=N, w0=w0, w0h=w0h, s0=s0,
                seed=int(g["seed"]))


def oracle_model(c, cdtype=torch.complex64):
    """TorchOracle with the deterministic weights the golden file was generated with."""
    m = O.TorchOracle(c["kind"], c["in_f"], c["hidden"], c["H"], c["out_f"], c["w0"], c["w0h"], c["s0"],
                      cdtype=torch.complex64)
    m.load_state_dict(O.deterministic_state(m, c["seed"]), strict=True)
    if cdtype == torch.complex128:
        for p in m.parameters():
            p.data = p.data.to(torch.complex128 if p.is_complex() else torch.float64)
    return m


def rel_err(a, b):
    """||a-b|| / ||b|| over everything (the 'relative RMS error' of SURVEY.md §7)."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = float(np.linalg.norm(b.reshape(-1)))
    return float(np.linalg.norm((a - b).reshape(-1))) / (den if den > 0 else 1.0)


def golden_grad(c, tag, key, full):
    """Compare helper: golden grads of the big cases are stored as a fixed subsample."""
    g = c["g"]
    ref = g[f"grad_{tag}.{key}"]
    full = np.asarray(full)
    if f"gradidx.{key}" in g.files and ref.size != full.size:
        return full.reshape(-1)[g[f"gradidx.{key}"]], ref
    return full, ref


def run_oracle(model, coords, grad_out):
    coords = coords.clone().requires_grad_(True)
    out = model(coords)
    (out * grad_out).sum().backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return out.detach(), grads, coords.grad.detach()


def record(section, key, value):
    """Measured errors of this run -> gpurun_out/parity_measured.json (evidence copied into profiles/; never read by a test)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "parity_measured.json")
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data.setdefault(section, {})[key] = value
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except (OSError, ValueError):
        pass
