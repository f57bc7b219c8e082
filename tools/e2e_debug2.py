"""Where does the e2e (host-buffer) step lose time against the device-resident step? (debug helper, GPU box)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200, bench

dev = torch.device("cuda", 0)
model = wire_b200.get_INR(**bench.CFG).to(dev)
n = 512 * 512
_, noisy = bench.synthetic_image(512, 512)
coords_h = bench.image_coords(512, 512).pin_memory()
target_h = torch.from_numpy(noisy.reshape(1, n, 3)).pin_memory()
coords, target = coords_h.to(dev), target_h.to(dev)
tr = wire_b200.Trainer(model, lr=5e-3)
for _ in range(5):
    tr.step(coords, target); tr.step(coords_h, target_h)
K = 100

def loop(name, inputs, read_loss):
    loss_pin = torch.zeros(K).pin_memory()
    evs = [torch.cuda.Event() for _ in range(K)]
    torch.cuda.synchronize(); t0 = time.perf_counter(); host = 0.0
    for i in range(K):
        h0 = time.perf_counter()
        loss = tr.step(*inputs)
        if read_loss:
            loss_pin[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
            evs[i].record()
        host += time.perf_counter() - h0
        if read_loss and i > 0:
            evs[i - 1].synchronize(); float(loss_pin[i - 1])
    torch.cuda.synchronize()
    print(f"{name:40s} {1e3 * (time.perf_counter() - t0) / K:.4f} ms/step   host enqueue {1e3 * host / K:.4f} ms/step")

loop("device inputs, no loss read", (coords, target), False)
loop("device inputs, loss read 1 step later", (coords, target), True)
loop("pinned inputs, no loss read", (coords_h, target_h), False)
loop("pinned inputs, loss read 1 step later", (coords_h, target_h), True)
# raw H2D time of the two buffers
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    tr.coords_buf.copy_(coords_h.reshape(n, 2), non_blocking=True); tr.target_buf.copy_(target_h.reshape(n, 3), non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"raw H2D of one step's inputs: {e0.elapsed_time(e1) / 20:.4f} ms")
