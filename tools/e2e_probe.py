"""Probe: end-to-end step time of Trainer.step(pinned host inputs) against how far the host may run ahead of the loss read."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import wire_b200

dev = torch.device("cuda", 0)
size = 512; n = size * size
torch.manual_seed(0)
model = wire_b200.get_INR(**B.CFG).to(dev)
_, noisy = B.synthetic_image(size, size)
coords_h = B.image_coords(size, size).pin_memory()
target_h = torch.from_numpy(noisy.reshape(1, n, 3)).pin_memory()
coords, target = coords_h.to(dev), target_h.to(dev)
tr = wire_b200.Trainer(model, lr=5e-3)
for _ in range(5):
    tr.step(coords, target)
torch.cuda.synchronize()
for steps in (20, 200):
    for lag in (1, 2, 4, 8, 32):
        loss_pin = torch.zeros(steps, dtype=torch.float32).pin_memory()
        evs = [torch.cuda.Event() for _ in range(steps)]
        side = torch.cuda.Stream(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); host = 0.0
        for i in range(steps):
            h0 = time.perf_counter()
            loss = tr.step(coords_h, target_h)
            done = torch.cuda.Event(); done.record()
            with torch.cuda.stream(side):
                side.wait_event(done)
                loss_pin[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
                evs[i].record(side)
            host += time.perf_counter() - h0
            if i >= lag:
                evs[i - lag].synchronize(); float(loss_pin[i - lag])
        for i in range(max(0, steps - lag), steps):
            evs[i].synchronize()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        print(f"steps {steps} lag {lag}: e2e {dt*1e3:.4f} ms/step, host enqueue {host/steps*1e3:.4f} ms/step", flush=True)
# device-resident reference
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): tr.step(coords, target)
e1.record(); torch.cuda.synchronize()
print("device-resident 200 steps:", e0.elapsed_time(e1) / 200, "ms/step")
