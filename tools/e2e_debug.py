"""Where does host time go in one training step? (debug helper, run on the GPU box)"""
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
opt = torch.optim.Adam(model.parameters(), lr=5e-3)

def sync():
    torch.cuda.synchronize()

def timed(label, fn, acc):
    sync(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
    acc.setdefault(label, []).append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    return r

for _ in range(5):
    out = model(coords); loss = ((out - target) ** 2).mean(); opt.zero_grad(); loss.backward(); opt.step()
acc = {}
for _ in range(20):
    c = timed("h2d", lambda: (coords_h.to(dev, non_blocking=True), target_h.to(dev, non_blocking=True)), acc)
    out = timed("forward", lambda: model(c[0]), acc)
    loss = timed("loss", lambda: ((out - c[1]) ** 2).mean(), acc)
    timed("zero_grad", lambda: opt.zero_grad(set_to_none=True), acc)
    timed("backward", lambda: loss.backward(), acc)
    timed("adam", lambda: opt.step(), acc)
    timed("item", lambda: float(loss.detach()), acc)
for k, v in acc.items():
    v = v[5:]
    print(f"{k:10s} host-enqueue {sum(a for a, _ in v) / len(v):7.3f} ms   enqueue+device {sum(b for _, b in v) / len(v):7.3f} ms")
# async loop
sync(); t0 = time.perf_counter()
for _ in range(20):
    out = model(coords); loss = ((out - target) ** 2).mean(); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
print(f"async loop: host {1e3 * (t1 - t0) / 20:.3f} ms/step, total {1e3 * (t2 - t0) / 20:.3f} ms/step")
