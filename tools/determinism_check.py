"""Run-to-run reproducibility of the CUDA path: forward outputs must be bit-identical, gradients may differ only by the
order of fp32 atomics (~1e-6 relative).  Anything larger points at a race."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wire_b200

def rel(a, b):
    a = torch.view_as_real(a) if a.is_complex() else a
    b = torch.view_as_real(b) if b.is_complex() else b
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

for precision in ("tf32", "mixed16"):
    for kind, hidden in (("wire", 300), ("wire2d", 256)):
        for n in (3000, 262144):
            torch.manual_seed(0)
            m = wire_b200.get_INR(kind, 2, hidden, None, 2, 3, True, 7.0, 7.0, 6.0, precision=precision).cuda()
            coords = (torch.rand(1, n, 2, device="cuda") * 2 - 1)
            target = torch.rand(1, n, 3, device="cuda")
            outs, grads = [], []
            for rep in range(4):
                out = m(coords)
                loss = ((out - target) ** 2).mean()
                g = torch.autograd.grad(loss, [p for p in m.parameters() if p.requires_grad])
                outs.append(out.detach().clone()); grads.append([x.clone() for x in g])
            torch.cuda.synchronize()
            out_same = all(torch.equal(outs[0], o) for o in outs[1:])
            gmax = max(rel(a, b) for gs in grads[1:] for a, b in zip(gs, grads[0]))
            print(f"{precision:8s} {kind:7s} n={n:7d}: outputs bit-identical={out_same}  max grad rel diff={gmax:.2e}", flush=True)
            if gmax > 1e-5:
                names = [k for k, p in m.named_parameters() if p.requires_grad]
                for i, k in enumerate(names):
                    print("      ", k, " ".join(f"{rel(gs[i], grads[0][i]):.1e}" for gs in grads[1:]), flush=True)
