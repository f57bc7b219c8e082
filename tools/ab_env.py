"""A/B of environment switches inside one process: per-kernel device time and graph step time of one training step.
    python tools/ab_env.py wire 2 300 2 3 7.0 6.0 262144 VAR=a,b [VAR2=c,d ...]   (settings are run one variable at a time)"""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200

kind, in_f, hidden, H, out_f, w0, s0, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6]), float(sys.argv[7]), int(sys.argv[8])
dev = torch.device("cuda", 0)
lib = wire_b200._lib.load()
coords = torch.rand(1, n, in_f, device=dev) * 2 - 1
target = torch.rand(1, n, out_f, device=dev)


def run(tag):
    torch.manual_seed(0)
    model = wire_b200.get_INR(kind, in_f, hidden, None, H, out_f, True, w0, w0, s0, precision="mixed16").to(dev)
    tr = wire_b200.Trainer(model, lr=5e-3)
    for _ in range(5):
        tr.step(coords, target)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            tr.step(coords, target)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    tr.use_graph = False
    lib.wire_b200_prof_reset(); lib.wire_b200_prof_enable(1)
    for _ in range(10):
        loss = tr.step(coords, target)
    torch.cuda.synchronize()
    lib.wire_b200_prof_enable(0)
    out = {}
    for k in range(lib.wire_b200_prof_kinds()):
        cnt, t = ctypes.c_uint64(0), ctypes.c_double(0.0)
        lib.wire_b200_prof_get(k, ctypes.byref(cnt), ctypes.byref(t))
        if cnt.value:
            out[lib.wire_b200_prof_name(k).decode()] = round(t.value / 10 * 1000, 1)
    print(json.dumps({"tag": tag, "graph_ms": round(best, 4), "loss": float(loss), "us": out}), flush=True)


run("default")
for spec in sys.argv[9:]:
    var, vals = spec.split("=")
    for v in vals.split(","):
        os.environ[var] = v
        run(f"{var}={v}")
    del os.environ[var]
