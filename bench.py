#!/usr/bin/env python
"""bench.py — WIRE fwd+bwd(+Adam) coordinates/second on a 512x512 image fit (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...         # coordinate-sharded data parallel, weak scaling

A step = one full training iteration of wire_image_denoise.py:148-157 on one batch of 262 144 coordinates
per GPU: forward through the complex Gabor stack, MSE loss, backward, Adam step.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "wire_fwd_bwd_coords_per_sec"
UNIT = "coords/s"
# the headline workload: wire_image_denoise.py defaults (hidden 300 -> M = 212, 2 hidden layers, omega0 7, sigma0 6)
CFG = dict(nonlin="wire", in_features=2, hidden_features=300, hidden_layers=2, out_features=3,
           first_omega_0=7.0, hidden_omega_0=7.0, scale=6.0)
LR = 5e-3
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (mean of the H forward launches of a step), from
# `ncu --set full` captures of this command at the default size: (precision, kernel) -> bytes
#   tf32:    profiles/r01_ncu_full_v6_summary.txt (530.7 MB read + 835.0 MB written, hidden layer 1)
#   mixed16: profiles/r02_ncu_full_step_kernels.txt (layer 1: 251.1 + 399.6 MB; layer 2 with the fused final Linear: 242.2 + 183.3 MB)
NCU_TRAFFIC = {("tf32", "tc_rows_gabor_fwd"): 1365.6e6, ("mixed16", "tc_rows_gabor_fwd"): 538.1e6}


def flop_per_coord(M, H, in_f, out_f):
    """F_wire = 24 H M^2 + 12 M out + 4 in M (complex MAC = 8 real flop)."""
    return 24 * H * M * M + 12 * M * out_f + 4 * in_f * M


def synthetic_image(H, W, seed=0):
    """Band-limited sinusoids + hard-edged discs, normalised to [0,1], plus Gaussian noise."""
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
    chans = []
    for c in range(3):
        im = np.zeros_like(xx)
        for _ in range(6):
            fx, fy, ph = rs.uniform(-6, 6), rs.uniform(-6, 6), rs.uniform(0, 2 * np.pi)
            im += rs.uniform(0.2, 1.0) * np.sin(fx * xx + fy * yy + ph)
        for _ in range(3):
            cx, cy, r = rs.uniform(-0.7, 0.7), rs.uniform(-0.7, 0.7), rs.uniform(0.1, 0.3)
            im += 1.5 * ((xx - cx) ** 2 + (yy - cy) ** 2 < r * r)
        chans.append(im)
    img = np.stack(chans, -1)
    img = (img - img.min()) / (img.max() - img.min())
    noisy = img + 0.1 * rs.normal(size=img.shape)
    return img.astype(np.float32), noisy.astype(np.float32)


def image_coords(H, W):
    x = torch.linspace(-1, 1, W)
    y = torch.linspace(-1, 1, H)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    return torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None, ...]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.rows, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(size, steps, warmup, threads=None):
    """The reference's CPU path: the oracle port (same op sequence as modules/wire.py) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wire_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = O.TorchOracle("wire", CFG["in_features"], CFG["hidden_features"], CFG["hidden_layers"], CFG["out_features"],
                          CFG["first_omega_0"], CFG["hidden_omega_0"], CFG["scale"])
    _, noisy = synthetic_image(size, size)
    coords = image_coords(size, size)
    target = torch.from_numpy(noisy.reshape(1, size * size, 3))
    opt = torch.optim.Adam(model.parameters(), lr=LR)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = model(coords)
        loss = ((out - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        float(loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    n = size * size
    return dict(coords_per_s=n / statistics.mean(times), ms_per_step=1e3 * statistics.mean(times), cores=threads,
                sample=f"{size}x{size} image ({n} coords) full-batch fwd+bwd+Adam, fp32/complex64, {steps} timed steps "
                       f"after {warmup} warm-up, torch {torch.__version__} CPU")


def gpu_eager_reference_run(size, dev, steps=5, warmup=2):
    """Like-for-like GPU bar: the same oracle port executed by eager PyTorch on this GPU in complex64
    with TF32 disabled — i.e. what the reference's own code does after `model.cuda()`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wire_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = O.TorchOracle("wire", CFG["in_features"], CFG["hidden_features"], CFG["hidden_layers"], CFG["out_features"],
                          CFG["first_omega_0"], CFG["hidden_omega_0"], CFG["scale"]).to(dev)
    _, noisy = synthetic_image(size, size)
    coords = image_coords(size, size).to(dev)
    target = torch.from_numpy(noisy.reshape(1, size * size, 3)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=LR)

    def step():
        loss = ((model(coords) - target) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "value": size * size / ms * 1e3, "unit": UNIT, "kind": "port (eager torch complex64 on this GPU, allow_tf32=False)",
            "sample": f"{size}x{size} full batch, {steps} timed steps"}


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (the oracle port: same op sequence as
    modules/wire.py, `/root/reference` does not travel to the GPU box) on all host cores, on THIS arm's config — the full
    512x512 batch per step (about 3 s per step on 16 cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = args.size
    r = cpu_reference_run(size, args.steps, max(args.warmup, 1))
    M = int(CFG["hidden_features"] / np.sqrt(2))
    n = size * size
    line = {"impl": "reference", "metric": METRIC, "value": r["coords_per_s"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"WIRE image fit {size}x{size} RGB ({n} coords/GPU full-batch fwd+bwd+Adam), "
                                   f"wire_image_denoise.py defaults: hidden 300 -> M={M}, H=2, omega0=7, sigma0=6",
                       "width": M, "hidden_layers": CFG["hidden_layers"], "coords_per_gpu": n,
                       "api": "oracle port of modules/wire.py INR + torch.optim.Adam on the host CPU (one process, all cores)"},
            "cpu_baseline": {"value": r["coords_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["coords_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def pin_to_gpu_numa_node(local):
    """Bind this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node of its GPU.
    torchrun starts every rank unbound; with eight ranks on one socket the per-step staging copies and the Python thread of
    each rank compete for the same memory controller (SCALE_r01: e2e efficiency 0.71-0.85 against 0.98 device-timed)."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = out[-12:] if len(out) >= 12 else out          # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        return None
    return None


def timed_steps(trainer, coords, target, steps, barrier, world, dev, dist):
    """`steps` training steps on device-resident inputs between CUDA events; max over ranks; ms per step."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        trainer.step(coords, target)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t)
    return ms_total / steps


E2E_LAG = 8   # the host consumes the loss of step i while step i + E2E_LAG is being enqueued


def run_ours(args):
    import torch.distributed as dist
    import wire_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that print from C (NCCL's "NCCL version ..." banner) go to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    affinity = pin_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = wire_b200._lib.load()
    wire_b200._lib.check(lib.wire_b200_device_ok(), "device check")

    size = args.size
    n = size * size
    M = int(CFG["hidden_features"] / np.sqrt(2))
    H = CFG["hidden_layers"]
    torch.manual_seed(0)
    model = wire_b200.get_INR(**CFG, precision=args.precision).to(dev)
    # weak scaling: every rank fits its own 512x512 tile of a (512*world) x 512 synthetic image
    _, noisy = synthetic_image(size, size, seed=rank)
    coords_h = image_coords(size, size).pin_memory()
    target_h = torch.from_numpy(noisy.reshape(1, n, 3)).pin_memory()
    coords = coords_h.to(dev)
    target = target_h.to(dev)
    # the public training API: one call = forward + MSE + backward (+ gradient exchange) + Adam; the constructor broadcasts
    # rank 0's parameters
    trainer = wire_b200.Trainer(model, lr=LR, graph=not args.no_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def counters():
        tot = 0
        for k in range(lib.wire_b200_prof_kinds()):
            cnt = ctypes.c_uint64(0)
            lib.wire_b200_prof_get(k, ctypes.byref(cnt), None)
            tot += cnt.value
        return tot

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        trainer.step(coords, target)
    barrier()
    # kernels launched by one step (a graph replay re-launches exactly what one eager step launches)
    lib.wire_b200_prof_enable(0)
    lib.wire_b200_prof_reset()
    saved = trainer.use_graph
    trainer.use_graph = False
    trainer.step(coords, target)
    trainer.use_graph = saved
    barrier()
    launches_per_step = counters()

    # ---------------- timed region: device-resident inputs, CUDA events, max over ranks ----------------
    ms_step = timed_steps(trainer, coords, target, args.steps, barrier, world, dev, dist) if not args.e2e_first else None

    # ---------------- e2e: pinned host inputs, H2D every step, D2H loss copy every step ----------------
    # Every step copies its inputs host -> device and its loss device -> host (both inside the timed region).  The HOST reads
    # the loss of step i only E2E_LAG steps later (the trainer keeps 256 steps of losses in its device ring, as the drivers'
    # tqdm/loss.item() bookkeeping allows): a rank's Python thread may then run ahead of its GPU by a few steps, which
    # absorbs host jitter instead of passing every hiccup of any rank on to all ranks through the in-kernel barrier.
    loss_pin = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    evs = [torch.cuda.Event() for _ in range(args.steps)]
    losses = []
    side = torch.cuda.Stream(dev)
    for _ in range(3):                                       # warm-up of the host-input route (staging buffers, copy stream)
        trainer.step(coords_h, target_h)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = trainer.step(coords_h, target_h)            # host -> device copies of this step's inputs inside
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(side):                        # device -> host copy of this step's loss, off the compute stream's
            side.wait_event(done)                            # critical path
            loss_pin[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
            evs[i].record(side)
        if i >= E2E_LAG:                                     # consume an earlier step's loss on the host
            evs[i - E2E_LAG].synchronize()
            losses.append(float(loss_pin[i - E2E_LAG]))
    for i in range(max(0, args.steps - E2E_LAG), args.steps):
        evs[i].synchronize()
        losses.append(float(loss_pin[i]))
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt)
    e2e = {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": coords_h.numel() * 4 + target_h.numel() * 4,
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3, "final_loss": losses[-1], "host_read_lag_steps": E2E_LAG,
           "host_affinity": affinity,
           "measured": "first (--e2e-first)" if args.e2e_first else "after the device-timed region (later = lower clocks under the power cap)",
           "note": "trainer.step(host pinned coords, host pinned target): H2D staged on a copy stream; the loss of every step is "
                   "copied to pinned memory on a side stream and read by the host E2E_LAG steps later; wall clock, max over ranks"}
    if ms_step is None:   # --e2e-first (experiment: which region sees the fresher clocks)
        ms_step = timed_steps(trainer, coords, target, args.steps, barrier, world, dev, dist)
    value = world * n / (ms_step * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + e2e region"

    # ---------------- sustained: >= 2000 consecutive steps (a real fit is 2000 iterations), its own clock samples ----------
    sustained = None
    if args.sustained_steps > 0:
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        ms_sus = timed_steps(trainer, coords, target, args.sustained_steps, barrier, world, dev, dist)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"steps": args.sustained_steps, "ms_per_step": ms_sus, "value": world * n / (ms_sus * 1e-3), "unit": UNIT,
                     "clocks": c2}

    # ---------------- the nn.Module + torch.optim.Adam route (what the reference drivers call) ----------------
    module_api = None
    if world == 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        opt = torch.optim.Adam(model.parameters(), lr=LR)
        for _ in range(3):
            loss = ((model(coords) - target) ** 2).mean(); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            loss = ((model(coords) - target) ** 2).mean(); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
        e1.record()
        barrier()
        module_ms = e0.elapsed_time(e1) / args.steps
        module_api = {"ms_per_step": module_ms, "value": n / (module_ms * 1e-3), "unit": UNIT,
                      "note": "the drop-in route the reference's drivers take unchanged: model(coords) (autograd.Function over the "
                              "C ABI) -> torch MSE -> loss.backward() -> torch.optim.Adam.step(), eager, device-resident inputs"}
        del opt

    # ---------------- per-kernel device time (CUDA events on the launching stream, eager replay of the step) ----
    roofline, kernels = None, {}
    # every rank replays the steps (the data-parallel step contains a collective); rank 0 reports its own kernels
    lib.wire_b200_prof_reset()
    lib.wire_b200_prof_enable(1 if rank == 0 else 0)
    trainer.use_graph = False
    barrier()
    for _ in range(args.steps):
        trainer.step(coords, target)
    barrier()
    trainer.use_graph = saved
    step_flop = flop_per_coord(M, H, CFG["in_features"], CFG["out_features"]) * n
    peaks, peaks_src = load_peaks()
    mixed = args.precision == "mixed16"
    # the peak of the PIPE USED: kind::f16 MMAs against the measured cuBLAS BF16 figure (burst: these regions last tens of
    # milliseconds), kind::tf32 MMAs issue at half that rate; FP32 FMAs have no tensor roofline
    tensor_peak = peaks["bf16_tflops"] / (1.0 if mixed else 2.0)
    if rank == 0:
        total_ms = 0.0
        for k in range(lib.wire_b200_prof_kinds()):
            cnt, ms = ctypes.c_uint64(0), ctypes.c_double(0.0)
            lib.wire_b200_prof_get(k, ctypes.byref(cnt), ctypes.byref(ms))
            if cnt.value:
                kernels[lib.wire_b200_prof_name(k).decode()] = {"launches": cnt.value, "ms_total": ms.value,
                                                                "ms_avg": ms.value / cnt.value}
                total_ms += ms.value
        lib.wire_b200_prof_enable(0)
        lib.wire_b200_prof_reset()
        top = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
        gemm_flop = 8.0 * M * M * n  # one complex M x M GEMM over n coordinates (fwd, dgrad and wgrad alike), SURVEY 8(d)
        unit_b = 8.0 * M * n         # one complex fp32 activation tensor [n, M] in HBM
        # algorithmic HBM bytes per launch of THIS design, in units; a = activation y, g = gradient g_z, z = saved
        # pre-activation (FP16 on both tensor-core paths), real g_z0 = half of that.  The forward entry is the average over the H
        # launches (the last layer's y is never written: the final Linear is fused into its epilogue).
        a_u = 0.5 if mixed else 1.0
        g_u = 0.5 if mixed else 1.0
        z_u = 1.0 if args.precision == "fp32" else 0.5
        g0_u = 0.25 if mixed else 0.5   # real g_z0 [n, M]: BF16 on the mixed16 path, fp32 otherwise
        alg_units = {"first_fwd": a_u, "tc_rows_gabor_fwd": (H * (a_u + z_u) + (H - 1) * a_u) / H, "top_bwd": z_u + g_u,
                     "tc_wgrad": a_u + g_u, "tc_rows_dgrad_gabor_bwd": 2 * g_u + z_u, "tc_rows_dgrad_first_bwd": g_u + g0_u,
                     "first_wgrad": g0_u}
        alg_bytes = {k: v * unit_b for k, v in alg_units.items()}
        step_units = (alg_units["first_fwd"] + H * alg_units["tc_rows_gabor_fwd"] + alg_units["top_bwd"] + H * alg_units["tc_wgrad"]
                      + (H - 1) * alg_units["tc_rows_dgrad_gabor_bwd"] + alg_units["tc_rows_dgrad_first_bwd"] + alg_units["first_wgrad"])
        for k, v in kernels.items():
            if k in alg_bytes:
                v["hbm_gbs"] = alg_bytes[k] / (v["ms_avg"] * 1e-3) / 1e9
                v["hbm_frac"] = v["hbm_gbs"] / peaks["hbm_gbs"]
            if k.startswith("tc_"):
                v["tflops"] = gemm_flop / (v["ms_avg"] * 1e-3) / 1e12
                v["tensor_frac"] = v["tflops"] / tensor_peak
        if top is not None and "tflops" in kernels[top]:
            # the binding roofline of the dense complex contraction is the TENSOR pipe; HBM is the co-constraint
            ach = kernels[top]["tflops"]
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                        "frac": ach / tensor_peak,
                        "traffic": NCU_TRAFFIC.get((args.precision, top)) if size == 512 else None,
                        "peak_source": f"{peaks_src} MEASURED_PEAKS.json bf16_tflops (burst)" + ("" if mixed else " / 2 (kind::tf32 issues at half the BF16 rate)"),
                        "algorithmic_flop_per_launch": gemm_flop,
                        "share_of_step": kernels[top]["ms_total"] / total_ms if total_ms else None,
                        "hbm": {"achieved": kernels[top].get("hbm_gbs"), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": kernels[top].get("hbm_frac"), "algorithmic_bytes_per_launch": alg_bytes.get(top),
                                "peak_source": f"{peaks_src} MEASURED_PEAKS.json hbm_gbs"},
                        "step": {"algorithmic_flop": step_flop, "tflops": step_flop / (ms_step * 1e-3) / 1e12,
                                 "tensor_frac": step_flop / (ms_step * 1e-3) / 1e12 / tensor_peak,
                                 "design_bytes": step_units * unit_b,
                                 "hbm_frac": step_units * unit_b / (ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                 "fused_minimum_bytes": float(n * (4 * CFG["in_features"] + 8 * CFG["out_features"]))}}

    # ---------------- secondary: the 32-bit-operand (TF32) mode on the same workload ----------------
    trainer.close()
    del trainer
    tf32 = None
    if args.precision == "mixed16" and not args.no_extras:
        m32 = wire_b200.get_INR(**CFG, precision="tf32").to(dev)
        t32 = wire_b200.Trainer(m32, lr=LR, graph=not args.no_graph)
        for _ in range(3):
            t32.step(coords, target)
        ms32 = timed_steps(t32, coords, target, args.steps, barrier, world, dev, dist)
        tf32 = {"ms_per_step": ms32, "value": world * n / (ms32 * 1e-3), "unit": UNIT,
                "algorithmic_tflops_per_gpu": step_flop / (ms32 * 1e-3) / 1e12,
                "frac_of_nominal_tf32_peak": step_flop / (ms32 * 1e-3) / 1e12 / 1100.0,
                "frac_of_measured_bf16_burst_halved": step_flop / (ms32 * 1e-3) / 1e12 / (peaks["bf16_tflops"] / 2.0)}
        t32.close()
        del t32, m32
        torch.cuda.empty_cache()

    # ---------------- BASELINE configs [2] and [3] under the same clock: wire2d SISR 1024^2, occupancy 512^3 ----------------
    extras = {}
    if not args.no_extras:
        if world == 1:
            try:
                extras["sisr_1024"] = sisr_extra(dev, args.steps)
            except Exception as exc:
                extras["sisr_1024"] = {"error": repr(exc)[:300]}
        if world == 1:
            try:
                extras["width_sweep"] = width_sweep_extra(dev, max(5, args.steps // 2))
            except Exception as exc:
                extras["width_sweep"] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import occupancy_bench as OB
            vol = OB.synthetic_volume(args.occupancy_size, dev)
            occ = {}
            for scaling in (("weak", "strong") if world > 1 else ("weak",)):
                occ[scaling] = OB.run_occupancy(dev, world, rank, size=args.occupancy_size, chunk=200000, steps=args.occupancy_steps,
                                                warmup=5, scaling=scaling, precision=args.precision, iou=True, vol=vol)
            if world == 1:
                occ["strong"] = occ["weak"]    # one GPU: the same run
            del vol
            extras["occupancy_512cube"] = occ
        except Exception as exc:
            extras["occupancy_512cube"] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    if rank == 0:
        cpu = cpu_reference_run(size, 3, 1) if (world == 1 and not args.no_cpu_baseline) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"tf32": "tf32", "fp32": "f32", "mixed16": "f16 activations x bf16 gradients, f32 accumulate"}[args.precision],
                "data": "synthetic",
                "config": {"workload": f"WIRE image fit {size}x{size} RGB ({n} coords/GPU full-batch fwd+bwd+Adam), "
                                       f"wire_image_denoise.py defaults: hidden 300 -> M={M}, H=2, omega0=7, sigma0=6",
                           "width": M, "hidden_layers": H, "coords_per_gpu": n,
                           "api": "wire_b200.Trainer.step" + (" (CUDA graph)" if saved and (world == 1 or launches_per_step) else ""),
                           "parallelism": f"coord-sharded dp{world}" if world > 1 else "single GPU",
                           "exchange": (None if world == 1 else "flat fp32 gradients summed by the Adam kernel with P2P loads over NVLink "
                                        "(no NCCL call per step) when the ranks share a host with peer access, else one all-reduce"),
                           "precision": args.precision,
                           "l2": "per-step activation traffic (>3.7 GB) far exceeds the 126 MB L2; no explicit flush"},
                "algorithmic_tflops": world * step_flop / (ms_step * 1e-3) / 1e12,
                "frac_of_measured_bf16_burst": step_flop / (ms_step * 1e-3) / 1e12 / tensor_peak,   # per GPU, peak of the pipe used
                "frac_of_nominal_tf32_peak": step_flop / (ms_step * 1e-3) / 1e12 / 1100.0,  # per GPU (the north-star's denominator)
                "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roofline, "sustained": sustained, "module_api": module_api, "tf32": tf32,
                "extras": extras, "kernels": kernels}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["gpu_eager_baseline"] = gpu_eager_reference_run(size, dev)
            except Exception as exc:  # e.g. out of memory on a shared box: the headline numbers do not depend on it
                line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:200]}
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu["coords_per_s"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                                    "sample": cpu["sample"]}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def width_sweep_extra(dev, steps):
    """BASELINE config [4] (tensor-pipe utilisation against width / depth / batch), a few points of tools/sweep.py under the
    bench's clock: full training step through wire_b200.Trainer on random coordinates, mixed16."""
    import wire_b200
    peaks, _ = load_peaks()
    out = []
    for hidden, H, n in ((128, 2, 1 << 18), (512, 2, 1 << 18), (1024, 2, 1 << 18), (300, 5, 1 << 18), (300, 2, 1 << 16), (300, 2, 1 << 22)):
        torch.manual_seed(0)
        model = wire_b200.get_INR(nonlin="wire", in_features=2, hidden_features=hidden, hidden_layers=H, out_features=3,
                                  first_omega_0=7.0, hidden_omega_0=7.0, scale=6.0).to(dev)
        tr = wire_b200.Trainer(model, lr=LR)
        coords = torch.rand(1, n, 2, device=dev) * 2 - 1
        target = torch.rand(1, n, 3, device=dev)
        for _ in range(3):
            tr.step(coords, target)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.step(coords, target)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        M = model.width
        tf = flop_per_coord(M, H, 2, 3) * n / (ms * 1e-3) / 1e12
        out.append({"hidden_features": hidden, "M": M, "H": H, "coords": n, "ms_per_step": ms, "coords_per_s": n / (ms * 1e-3),
                    "algorithmic_tflops": tf, "frac_of_measured_bf16_burst": tf / peaks["bf16_tflops"]})
        tr.close()
        del tr, model, coords, target
        torch.cuda.empty_cache()
    return out


def sisr_extra(dev, steps):
    """BASELINE config [2]: wire2d 4x super-resolution of a synthetic 1024x1024 image (wire_SISR.py:154-177: grad forward over
    the 1 048 576 HR coordinates, AvgPool2d(4) + MSE against the 256x256 LR image, backward, Adam; the reference's second
    no_grad forward is the same numbers as the first and is served by it — Trainer.step_sisr)."""
    import wire_b200
    Hh = Ww = 1024
    scale = 4
    img, _ = synthetic_image(Hh, Ww, seed=5)
    gt_hr = torch.from_numpy(img.reshape(Hh * Ww, 3)).to(dev)
    gt_lr = torch.nn.functional.avg_pool2d(gt_hr.reshape(Hh, Ww, 3).permute(2, 0, 1)[None], scale)[0].permute(1, 2, 0).reshape(-1, 3).contiguous()
    coords = image_coords(Hh, Ww).to(dev)
    torch.manual_seed(0)
    model = wire_b200.get_INR(nonlin="wire2d", in_features=2, hidden_features=256, hidden_layers=2, out_features=3,
                              first_omega_0=8.0, hidden_omega_0=8.0, scale=9.0).to(dev)     # wire_SISR.py:50-56,111-120
    tr = wire_b200.Trainer(model, lr=5e-3)
    tr.set_loss_avgpool(Hh, Ww, scale)
    for _ in range(3):
        tr.step_sisr(coords, gt_lr, gt_hr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, rec_hr, mse_hr = tr.step_sisr(coords, gt_lr, gt_hr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    M, H = model.width, 2
    train_flop = (48 * H * M * M + 12 * M * 3 + 8 * 2 * M) * Hh * Ww
    peaks, _ = load_peaks()
    res = {"workload": "wire2d 4x SISR, 1024x1024 HR (1 048 576 coords) -> 256x256 LR, M=128, H=2, omega0=8, sigma0=9",
           "ms_per_iteration": ms, "coords_per_s": Hh * Ww / (ms * 1e-3), "steps": steps,
           "algorithmic_tflops": train_flop / (ms * 1e-3) / 1e12,
           "frac_of_measured_bf16_burst": train_flop / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
           "loss": float(loss), "mse_hr": float(mse_hr),
           "note": "one iteration = grad forward + pooled MSE + backward + Adam + HR metrics (MSE against the HR image on the "
                   "device); the reference's second no_grad forward is shared with the first (bit-identical inputs and weights)"}
    tr.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--precision", default="mixed16", choices=["mixed16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--e2e-first", action="store_true", help="measure the end-to-end region before the device-timed one")
    ap.add_argument("--no-extras", action="store_true", help="skip the tf32 / wire2d SISR / occupancy 512^3 secondary measurements")
    ap.add_argument("--sustained-steps", type=int, default=2000)
    ap.add_argument("--occupancy-size", type=int, default=512)
    ap.add_argument("--occupancy-steps", type=int, default=200)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
