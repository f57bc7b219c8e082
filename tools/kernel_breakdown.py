"""Per-kernel device time of one training step for any config (CUDA events around every launch, eager replay).
    python tools/kernel_breakdown.py wire2d 2 256 2 3 8.0 9.0 1048576 [precision]"""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200

kind, in_f, hidden, H, out_f, w0, s0, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6]), float(sys.argv[7]), int(sys.argv[8])
precision = sys.argv[9] if len(sys.argv) > 9 else "mixed16"
dev = torch.device("cuda", 0)
lib = wire_b200._lib.load()
model = wire_b200.get_INR(kind, in_f, hidden, None, H, out_f, True, w0, w0, s0, precision=precision).to(dev)
tr = wire_b200.Trainer(model, lr=5e-3)
coords = torch.rand(1, n, in_f, device=dev) * 2 - 1
target = torch.rand(1, n, out_f, device=dev)
for _ in range(5):
    tr.step(coords, target)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    tr.step(coords, target)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
tr.use_graph = False
lib.wire_b200_prof_reset(); lib.wire_b200_prof_enable(1)
for _ in range(10):
    tr.step(coords, target)
torch.cuda.synchronize()
out = {}
for k in range(lib.wire_b200_prof_kinds()):
    cnt, t = ctypes.c_uint64(0), ctypes.c_double(0.0)
    lib.wire_b200_prof_get(k, ctypes.byref(cnt), ctypes.byref(t))
    if cnt.value:
        out[lib.wire_b200_prof_name(k).decode()] = (cnt.value // 10, round(t.value / 10, 4))
print(json.dumps({"config": sys.argv[1:], "M": model.width, "ms_per_step_graph": ms, "per_step (launches, ms)": out, "sum_ms": round(sum(v[1] for v in out.values()), 4)}))
