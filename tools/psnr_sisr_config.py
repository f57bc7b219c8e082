"""North-star bar on BASELINE config [2]: the reference's super-resolution loop (wire_SISR.py:118-178: wire2d, hidden 256 -> M = 128,
H = 2, omega0 = 8, sigma0 = 9, 4x, Adam 5e-3 with the 0.2^(k/niters) decay; every iteration = grad forward on the HR grid,
AvgPool2d(4) + MSE against the LR image, a second no_grad forward for the HR metrics, backward, step) on a synthetic HR image,
(a) on the oracle port of the reference (eager complex64 on the GPU, TF32 off), (b) on this repo's CUDA modules through the same
nn.Module + torch.optim loop, (c) on wire_b200.Trainer.step_sisr (fused loss / Adam / graph, the second forward shared).
Reports the best HR PSNR of each.     python tools/psnr_sisr_config.py [niters] [HR size] > profiles/rNN_psnr_sisr.json"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wire_oracle as O
import wire_b200
from test_trajectory_gpu import _oracle, _ours, synthetic_image, DEV

niters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
scale = 4
H2, W2 = H // scale, W // scale
img, _ = synthetic_image(H, W)
gt_img = torch.from_numpy(img).to(DEV)                                               # [H, W, 3]
down = torch.nn.AvgPool2d(scale)
gt = gt_img.reshape(H * W, 3)[None]
gt_lr = down(gt_img.permute(2, 0, 1)[None]).reshape(1, 3, -1).permute(0, 2, 1).contiguous()   # INTER_AREA at an integer factor = box mean
x_hr = torch.linspace(-1, 1, W, device=DEV); y_hr = torch.linspace(-1, 1, H, device=DEV)
X, Y = torch.meshgrid(x_hr, y_hr, indexing="xy")
coords_hr = torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None].contiguous()
cfg = ("wire2d", 2, 256, 2, 3, 8.0, 9.0)
torch.backends.cuda.matmul.allow_tf32 = False


def module_loop(model):
    optim = torch.optim.Adam(lr=5e-3, params=model.parameters())
    sched = torch.optim.lr_scheduler.LambdaLR(optim, lambda k: 0.2 ** min(k / niters, 1))
    best = float("inf")
    mses = torch.zeros(niters, device=DEV)
    for epoch in range(niters):
        rec_hr = model(coords_hr)
        rec = down(rec_hr.reshape(H, W, 3).permute(2, 0, 1)[None])
        loss = ((gt_lr - rec.reshape(1, 3, -1).permute(0, 2, 1)) ** 2).mean()
        with torch.no_grad():
            rec_hr = model(coords_hr)
            mses[epoch] = ((gt - rec_hr) ** 2).mean()
        optim.zero_grad(); loss.backward(); optim.step(); sched.step()
    return -10 * np.log10(float(mses.min()))


out = {"config": f"wire2d 4x SISR, HR {H}x{W} -> LR {H2}x{W2}, M=128, H=2, omega0=8, sigma0=9, {niters} iterations", "psnr_db": {}, "seconds": {}}
ref, init = _oracle(*cfg, seed=33)
ref_runs = int(os.environ.get("REF_RUNS", "1"))   # > 1: the reference's own run-to-run spread (its complex GEMM backward uses atomics)
for i in range(ref_runs):
    if i:
        ref, _ = _oracle(*cfg, seed=33)
    key = "reference_port_c64_gpu_eager" + ("" if i == 0 else f"_run{i + 1}")
    t0 = time.time(); out["psnr_db"][key] = module_loop(ref); torch.cuda.synchronize(); out["seconds"][key] = time.time() - t0
    del ref
for precision in os.environ.get("PRECISIONS", "mixed16").split(","):
    m = _ours(*cfg, init=init, precision=precision)
    key = f"wire_b200_{precision}_module_loop"
    t0 = time.time(); out["psnr_db"][key] = module_loop(m); torch.cuda.synchronize(); out["seconds"][key] = time.time() - t0
    del m
m = _ours(*cfg, init=init, precision="mixed16")
tr = wire_b200.Trainer(m, lr=5e-3)
tr.set_loss_avgpool(H, W, scale)
mses = torch.zeros(niters, device=DEV)
t0 = time.time()
for epoch in range(niters):
    tr.set_lr(5e-3 * 0.2 ** min(epoch / niters, 1))
    loss, rec_hr, mse_hr = tr.step_sisr(coords_hr, gt_lr, gt)
    mses[epoch] = mse_hr
torch.cuda.synchronize(); out["seconds"]["wire_b200_mixed16_trainer_step_sisr"] = time.time() - t0
out["psnr_db"]["wire_b200_mixed16_trainer_step_sisr"] = -10 * np.log10(float(mses.min()))
r = out["psnr_db"]["reference_port_c64_gpu_eager"]
out["diff_db"] = {k: v - r for k, v in out["psnr_db"].items() if k.startswith("wire_b200")}
print(json.dumps(out, indent=1))
