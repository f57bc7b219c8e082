"""Spread of the best-PSNR figure of tests/test_parity_gpu.py::test_training_psnr_parity_with_reference_loop over
initialisation seeds and precisions (GPU kernels only; fp32 kernels track the CPU reference to 0.003 dB)."""
import os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wire_oracle as O
import wire_b200
from test_parity_gpu import _train

H = W = 64
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rs = np.random.RandomState(0)
yy, xx = np.meshgrid(np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
img = np.stack([0.5 + 0.25 * np.sin(3 * xx + c) * np.cos(2 * yy - c) + 0.2 * ((xx - 0.2 * c) ** 2 + yy ** 2 < 0.2) for c in range(3)], -1).astype(np.float32)
img = (img - img.min()) / (img.max() - img.min())
noisy = (img + 0.1 * rs.normal(size=img.shape)).astype(np.float32)
coords = O.image_coords(H, W).cuda()
target = torch.from_numpy(noisy.reshape(1, H * W, 3)).cuda()
clean = torch.from_numpy(img.reshape(1, H * W, 3)).cuda()
for seed in range(11, 19):
    ref = O.TorchOracle("wire", 2, 300, 2, 3, 7.0, 7.0, 6.0)
    init = O.deterministic_state(ref, seed)
    row = []
    for precision in ("fp32", "fp32", "tf32", "mixed16"):
        m = wire_b200.get_INR(nonlin="wire", in_features=2, out_features=3, hidden_features=300, hidden_layers=2,
                              first_omega_0=7.0, hidden_omega_0=7.0, scale=6.0, precision=precision)
        m.load_state_dict(init, strict=True); m.cuda()
        row.append(_train(m, coords, target, clean, iters))
    print(f"seed {seed}: fp32 {row[0]:.3f} fp32(again) {row[1]:.3f} tf32 {row[2]:.3f} mixed16 {row[3]:.3f}", flush=True)
