import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200
dev = torch.device("cuda", 0)
kw = dict(nonlin="wire", in_features=3, hidden_features=300, hidden_layers=3, out_features=1, first_omega_0=20.0, hidden_omega_0=20.0, scale=10.0)
torch.manual_seed(0)
a = wire_b200.get_INR(**kw, precision="mixed16").to(dev)
b = wire_b200.get_INR(**kw, precision="fp32").to(dev)
b.load_state_dict(a.state_dict())
for n in (25000, 24992, 25088, 100000, 200000, 3125, 12500):
    g = torch.Generator().manual_seed(n)
    c = (torch.rand(1, n, 3, generator=g) * 2 - 1).to(dev)
    t = torch.rand(1, n, 1, generator=g).to(dev)
    res = []
    for m in (a, b):
        m.zero_grad(set_to_none=True)
        out = m(c)
        loss = ((out - t) ** 2).mean(); loss.backward()
        res.append((out.detach(), [torch.view_as_real(p.grad).clone() if p.grad.is_complex() else p.grad.clone() for p in m.parameters() if p.grad is not None]))
    oe = float((res[0][0] - res[1][0]).norm() / res[1][0].norm())
    ge = max(float((x - y).norm() / (y.norm() + 1e-30)) for x, y in zip(res[0][1], res[1][1]))
    print(f"n={n}: out rel err {oe:.3e}  max grad rel err {ge:.3e}", flush=True)
