"""A/B of an environment switch on gradients: runs the same forward + backward with VAR=0 and VAR=1 (AB_VAR, default
WIRE_B200_BIAS_SUM) for several widths of wire / wire2d and prints the worst relative difference over all parameter gradients.
    AB_VAR=WIRE_B200_FWGRAD_STREAM python tools/ab_grad_check.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import wire_b200, util, wire_oracle as O
for kind in ("wire", "wire2d"):
    for hid in (64, 128, 192, 256, 90):
        M = hid if kind == "wire2d" else hid  # complex width is derived inside get_INR; pass raw widths through hidden_features
        torch.manual_seed(0)
        m = wire_b200.get_INR(nonlin=kind, in_features=2, hidden_features=hid, hidden_layers=2, out_features=3,
                              first_omega_0=8.0, hidden_omega_0=8.0, scale=9.0, precision="mixed16").cuda()
        n = 40000 + 37
        c = torch.rand(1, n, 2, device="cuda") * 2 - 1
        g = torch.randn(1, n, 3, device="cuda") / n
        res = {}
        for flag in ("0", "1"):
            os.environ[os.environ.get("AB_VAR", "WIRE_B200_BIAS_SUM")] = flag
            m.zero_grad()
            (m(c) * g).sum().backward()
            res[flag] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        worst = 0
        for k in res["0"]:
            a, b = res["0"][k], res["1"][k]
            if a.is_complex(): a, b = torch.view_as_real(a), torch.view_as_real(b)
            e = float((a - b).norm() / a.norm().clamp_min(1e-30))
            worst = max(worst, e)
            if "bias" in k: print(kind, hid, k, "rel diff", e)
        print(kind, hid, "width", m.net[1].linear.weight.shape if hasattr(m.net[1], "linear") else "", "worst", worst, flush=True)
