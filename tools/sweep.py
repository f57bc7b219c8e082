"""Timing of the other BASELINE.json configs on one B200 (parity for them lives in tests/): 
  [2] wire2d 4x SISR step on 1024x1024 HR coords (one grad forward + one no_grad forward, AvgPool2d(4), MSE vs LR)
  [3] wire occupancy chunk (in 3, out 1, 3 hidden layers, omega0 20, s0 10, 200 000 coords per step)
  [4] width / depth / batch sweep (tensor-pipe utilisation)
Prints one JSON line per case; run on the GPU box: python tools/sweep.py > gpurun_out/sweep.jsonl"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200

dev = torch.device("cuda", 0)


def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def flops(kind, M, H, in_f, out_f):
    return (48 if kind == "wire2d" else 24) * H * M * M + 12 * M * out_f + (8 if kind == "wire2d" else 4) * in_f * M


def trainer_case(name, kind, in_f, hidden, H, out_f, w0, s0, n, steps=10):
    model = wire_b200.get_INR(kind, in_f, hidden, None, H, out_f, True, w0, w0, s0).to(dev)
    tr = wire_b200.Trainer(model, lr=5e-3)
    coords = torch.rand(1, n, in_f, device=dev) * 2 - 1
    target = torch.rand(1, n, out_f, device=dev)
    ms = timeit(lambda: tr.step(coords, target), steps)
    M = model.width
    f = flops(kind, M, H, in_f, out_f) * n
    print(json.dumps({"case": name, "kind": kind, "precision": model.precision, "M": M, "H": H, "n": n, "ms_per_step": ms, "coords_per_s": n / ms * 1e3,
                      "algorithmic_tflops": f / ms * 1e-9, "frac_nominal_tf32": f / ms * 1e-9 / 1100.0}), flush=True)
    del tr, model
    torch.cuda.empty_cache()


def sisr_case():
    """wire_SISR.py:154-177 restated: HR 1024x1024 coords, AvgPool2d(4), MSE vs a 256x256 LR image, plus the
    no_grad metric forward; nn.Module API + torch.optim.Adam (the loss is not a plain per-coordinate MSE)."""
    H = W = 1024
    model = wire_b200.get_INR("wire2d", 2, 256, None, 2, 3, True, 8.0, 8.0, 9.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    x = torch.linspace(-1, 1, W, device=dev); y = torch.linspace(-1, 1, H, device=dev)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    coords = torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None, ...].contiguous()
    lr_img = torch.rand(1, 3, H // 4, W // 4, device=dev)
    pool = torch.nn.AvgPool2d(4)

    def step():
        rec_hr = model(coords)
        rec = pool(rec_hr.reshape(H, W, 3).permute(2, 0, 1)[None, ...])
        loss = ((lr_img - rec) ** 2).mean()
        with torch.no_grad():
            model(coords)
        opt.zero_grad(); loss.backward(); opt.step()

    ms = timeit(step, 5, 2)
    n = H * W
    f = (flops("wire2d", 128, 2, 2, 3) + flops("wire2d", 128, 2, 2, 3) / 3) * n
    print(json.dumps({"case": "wire2d_sisr_1024", "M": 128, "H": 2, "n": n, "ms_per_step": ms, "coords_per_s": n / ms * 1e3,
                      "algorithmic_tflops": f / ms * 1e-9}), flush=True)


if __name__ == "__main__":
    trainer_case("wire_denoise_512", "wire", 2, 300, 2, 3, 7.0, 6.0, 512 * 512)
    trainer_case("wire_occupancy_chunk_2e5", "wire", 3, 300, 3, 1, 20.0, 10.0, 200000)
    trainer_case("wire_occupancy_chunk_2e6", "wire", 3, 300, 3, 1, 20.0, 10.0, 2000000, steps=5)
    sisr_case()
    for hidden in (128, 256, 512, 1024):
        for n in (1 << 16, 1 << 20):
            trainer_case(f"sweep_h{hidden}_n{n}", "wire", 2, hidden, 2, 3, 7.0, 6.0, n, steps=5)
    trainer_case("sweep_h300_H5_n2^18", "wire", 2, 300, 5, 3, 7.0, 6.0, 1 << 18, steps=5)
    trainer_case("wire2d_h256_n2^20", "wire2d", 2, 256, 2, 3, 8.0, 9.0, 1 << 20, steps=5)
