"""North-star bar at the real size: the reference's denoising loop (wire_image_denoise.py:123-178: 2 000 full-batch iterations of
Adam with the LambdaLR decay, best PSNR against the clean image) at BASELINE config [1] — 512 x 512 RGB, hidden 300 -> M = 212,
H = 2, omega0 = 7, sigma0 = 6 — once on the oracle port of the reference (eager complex64 on the GPU, TF32 off) and once on this
repo's CUDA modules from the same weights and permutations.  tests/test_trajectory_gpu.py asserts the same at 256^2 x 200.
    python tools/psnr_full_config.py [niters] [size] > profiles/rNN_psnr_512_2000.json"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wire_oracle as O
from test_trajectory_gpu import _oracle, _ours, synthetic_image, denoise_loop, DEV

niters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 512
img, noisy = synthetic_image(H, W)
coords = O.image_coords(H, W).to(DEV)
gt = torch.from_numpy(img.reshape(1, H * W, 3)).to(DEV)
gt_noisy = torch.from_numpy(noisy.reshape(1, H * W, 3)).to(DEV)
gen = torch.Generator().manual_seed(1234)
perms = [torch.randperm(H * W, generator=gen).to(DEV) for _ in range(niters)]   # full batch: the order only permutes the sum
cfg = ("wire", 2, 300, 2, 3, 7.0, 6.0)
torch.backends.cuda.matmul.allow_tf32 = False
out = {"config": f"{H}x{W} RGB, M=212, H=2, omega0=7, sigma0=6, {niters} full-batch iterations, lr 5e-3 * 0.1^(k/niters)", "psnr_db": {}, "seconds": {}}
ref, init = _oracle(*cfg, seed=21)
t0 = time.time(); out["psnr_db"]["reference_port_c64_gpu_eager"] = denoise_loop(ref, coords, gt_noisy, gt, perms, niters); torch.cuda.synchronize()
out["seconds"]["reference_port_c64_gpu_eager"] = time.time() - t0
del ref
for precision in ("mixed16", "tf32"):
    m = _ours(*cfg, init=init, precision=precision)
    t0 = time.time(); out["psnr_db"]["wire_b200_" + precision] = denoise_loop(m, coords, gt_noisy, gt, perms, niters); torch.cuda.synchronize()
    out["seconds"]["wire_b200_" + precision] = time.time() - t0
    del m
r = out["psnr_db"]["reference_port_c64_gpu_eager"]
out["diff_db"] = {k: v - r for k, v in out["psnr_db"].items() if k.startswith("wire_b200")}
print(json.dumps(out, indent=1))
