"""BASELINE.json config [3]: WIRE occupancy fit on a synthetic S^3 volume (default 512^3 = 134 M coords, chunks of 2e5),
coordinate-sharded data parallel on 1/2/4/8 B200 (wire_occupancy.py:107-158 restated on the on-device pipeline).

    python tools/occupancy_bench.py [--size 512] [--chunk 200000] [--steps 200] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/occupancy_bench.py ...

Every rank holds the whole volume (537 MB) and the same device permutation; each training step takes one chunk of the
permutation — weak scaling: chunk * N coordinates per step (chunk per GPU); strong: chunk coordinates per step split over the
ranks (trajectory-identical to one GPU up to summation order) — assembles coordinates + targets on the device
(GridBatcher), runs the fused step (forward, MSE, backward, peer-memory gradient exchange + Adam) and scatters the
prediction.  Timed with CUDA events, max over ranks; afterwards IoU at threshold 0.5 over the WHOLE volume by sharded
inference (volutils.get_IoU semantics).  One JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wire_b200  # noqa: E402
from wire_b200 import data, parallel  # noqa: E402


def synthetic_volume(S, dev):
    """Union of a few ellipsoids and a torus, ~10-20 % occupied; built slab by slab on the device."""
    lin = torch.linspace(-1, 1, S, device=dev)
    rs = np.random.RandomState(0)
    ell = [(rs.uniform(-0.5, 0.5, 3), rs.uniform(0.15, 0.45, 3)) for _ in range(5)]
    vol = torch.empty((S, S, S), dtype=torch.float32, device=dev)
    for i0 in range(0, S, 64):
        y = lin[i0:i0 + 64].view(-1, 1, 1)          # index i -> y_i, j -> x_j, k -> z_k (utils.get_coords order)
        x = lin.view(1, -1, 1)
        z = lin.view(1, 1, -1)
        occ = torch.zeros((y.shape[0], S, S), dtype=torch.bool, device=dev)
        for c, r in ell:
            occ |= ((x - c[0]) / r[0]) ** 2 + ((y - c[1]) / r[1]) ** 2 + ((z - c[2]) / r[2]) ** 2 < 1.0
        occ |= (torch.sqrt(x * x + y * y) - 0.6) ** 2 + z * z < 0.08 ** 2
        vol[i0:i0 + 64] = occ.float()
    return vol


def run_occupancy(dev, world, rank, size=512, chunk=200000, steps=200, warmup=5, scaling="weak", precision="mixed16",
                  iou=True, seed=0, vol=None):
    """One timed occupancy run (the process group, if any, is already initialised); returns the result dict on every rank.
    `vol`: a volume built earlier by synthetic_volume (re-used between the weak and the strong run)."""
    S = size
    N = S ** 3
    if vol is None:
        vol = synthetic_volume(S, dev)
    occupied = float(vol.mean())
    imten = vol.reshape(N, 1)
    batcher = wire_b200.GridBatcher((S, S, S), imten, linspace="numpy")
    torch.manual_seed(seed)
    model = wire_b200.get_INR(nonlin="wire", in_features=3, hidden_features=300, hidden_layers=3, out_features=1,
                              first_omega_0=20.0, hidden_omega_0=20.0, scale=10.0, precision=precision).to(dev)   # wire_occupancy.py:43-45,107-116
    tr = wire_b200.Trainer(model, lr=5e-3)   # (broadcasts the parameters of rank 0)
    gen = torch.Generator(device=dev).manual_seed(1234 + seed)      # same seed, same device type: same permutation on every rank
    perm = torch.randperm(N, device=dev, generator=gen)
    per_step = chunk * (world if scaling == "weak" else 1)
    per_step = min(per_step, N)
    est = torch.zeros(N, 1, device=dev)

    def step(i):
        b = (i * per_step) % (N - per_step + 1)
        c = perm[b:b + per_step]
        lo, hi = parallel.shard_range(per_step, rank, world)
        return tr.step_indexed(batcher, c[lo:hi], n_global=per_step if world > 1 else None, rec=est)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    batcher.check_indices()

    iou_val = None
    if iou:
        # IoU over the whole volume: each rank infers its contiguous shard of the grid, counts are summed over ranks
        lo, hi = parallel.shard_range(N, rank, world)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        blk = 1 << 22
        with torch.no_grad():
            for b in range(lo, hi, blk):
                n = min(blk, hi - b)
                c = batcher.coords(b, n)
                pred = model(c[None, ...]).reshape(n, 1).contiguous()
                counts += data.iou_counts(pred, imten[b:b + n], 0.5)
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        iou_val = float(counts[0] / counts[1])
    M, H = model.width, 3
    flop = 24 * H * M * M + 12 * M * 1 + 4 * 3 * M
    res = {"workload": f"WIRE occupancy {S}^3 ({N} coords), chunks of {chunk}", "n_gpus": world,
           "scaling": scaling, "coords_per_step": per_step, "steps": steps, "ms_per_step": ms,
           "coords_per_s": per_step / ms * 1e3, "epoch_s_at_this_rate": N / (per_step / ms * 1e3),
           "algorithmic_tflops": per_step * flop / ms * 1e-9, "frac_nominal_tf32_per_gpu": per_step * flop / ms * 1e-9 / 1100.0 / world,
           "precision": precision, "exchange": "peer" if tr.peer is not None else ("nccl" if world > 1 else None),
           "occupied_fraction": occupied, "final_chunk_loss_this_rank": float(loss),
           "iou_after_steps": iou_val, "steps_done": tr.steps_done, "seed": seed}
    tr.close()
    del tr, model, est, perm, batcher
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--chunk", type=int, default=200000)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="mixed16")
    ap.add_argument("--no-iou", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_occupancy(dev, world, rank, args.size, args.chunk, args.steps, args.warmup, args.scaling, args.precision,
                        not args.no_iou, args.seed)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
