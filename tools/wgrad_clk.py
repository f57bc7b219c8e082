"""Per-CTA phase stamps of tc_wgrad (WIRE_B200_WGRAD_CLK=1: K-loop cycles per chunk, epilogue, teardown) on one eager step.
    python tools/wgrad_clk.py wire2d 256 1048576"""
import os, sys, torch
sys.path.insert(0, ".")
import wire_b200
kind, hidden, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
m = wire_b200.get_INR(kind, 2, hidden, None, 2, 3, True, 10.0, 10.0, 10.0, precision="mixed16").cuda()
tr = wire_b200.Trainer(m, lr=5e-3, graph=False)
c = torch.rand(1, n, 2, device="cuda") * 2 - 1
t = torch.rand(1, n, 3, device="cuda")
for _ in range(3): tr.step(c, t)
torch.cuda.synchronize()
os.environ["WIRE_B200_WGRAD_CLK"] = "1"
for _ in range(2): tr.step(c, t)
torch.cuda.synchronize()
