"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): two eager Trainer steps + one module fwd/bwd for wire,
wire (occupancy shape) and wire2d at a ragged batch size, mixed16 and tf32.
    compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200  # noqa: E402

dev = torch.device("cuda", 0)
n = int(os.environ.get("SANITY_N", "1300"))
for precision in ("mixed16", "tf32"):
    for kind, in_f, hidden, H, out_f in (("wire", 2, 300, 2, 3), ("wire", 3, 300, 3, 1), ("wire2d", 2, 256, 2, 3)):
        torch.manual_seed(0)
        m = wire_b200.get_INR(kind, in_f, hidden, None, H, out_f, True, 7.0, 7.0, 6.0, precision=precision).to(dev)
        c = torch.rand(1, n, in_f, device=dev) * 2 - 1
        t = torch.rand(1, n, out_f, device=dev)
        loss = ((m(c) - t) ** 2).mean()
        loss.backward()
        tr = wire_b200.Trainer(m, lr=5e-3, graph=False)
        for _ in range(2):
            l = tr.step(c, t)
        torch.cuda.synchronize()
        print(f"{precision} {kind} in={in_f} H={H}: module loss {float(loss):.5f}, trainer loss {float(l):.5f}", flush=True)
print("SANITY OK")
