"""Exploration (scratch): which seeds give a healthy small occupancy fit with the ORACLE (GPU eager c64)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch, numpy as np
import wire_oracle as O
import test_trajectory_gpu as T
torch.backends.cuda.matmul.allow_tf32 = False
H = W = Tt = 64; maxpoints = 262144; niters = 600
N = H * W * Tt
vol = T.synthetic_volume(H, W, Tt, seed=0)
imten = torch.from_numpy(vol).reshape(N, 1).cuda()
coords_tab = torch.from_numpy(O.get_coords_np(H, W, Tt)).cuda()
perms = [torch.randperm(N, generator=torch.Generator().manual_seed(300 + e)).cuda() for e in range(niters)]
perms_b = [torch.randperm(N, generator=torch.Generator().manual_seed(7300 + e)).cuda() for e in range(niters)]
for tag, pp, cd, dt in (("c64", perms, torch.complex64, torch.float32), ("c64 other order", perms_b, torch.complex64, torch.float32), ("c128", perms, torch.complex128, torch.float64)):
    ref, init = T._oracle("wire", 3, 300, 3, 1, 20.0, 10.0, seed=32, cdtype=cd)
    t = time.time()
    ious = T.occupancy_loop(ref, coords_tab, imten, pp, niters, maxpoints, dtype=dt)
    print(f"{tag}: iou every 50th {[round(v,4) for v in ious[49::50]]} final {ious[-1]:.4f} ({time.time()-t:.1f}s)", flush=True)
