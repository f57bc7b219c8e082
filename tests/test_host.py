"""Host-side logic that needs no GPU: the drop-in module API, the C-ABI library surface, loud failure without
a CUDA device, the reference patch hook, and the coordinate-sharded data-parallel helpers (gloo, world_size 2)."""
import os
import re
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
import wire_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    import wire_b200
    header = open(os.path.join(ROOT, "include", "wire_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(wire_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = wire_b200._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wire_b200.h but not exported"
    assert declared == set(wire_b200._lib.SIGNATURES), "ctypes SIGNATURES and the header disagree"
    assert lib.wire_b200_abi_version() == 2
    assert lib.wire_b200_prof_kinds() >= 8


def test_state_dict_matches_reference_layout():
    import wire_b200
    m = wire_b200.get_INR(nonlin="wire", in_features=2, out_features=3, hidden_features=300, hidden_layers=2,
                          first_omega_0=7.0, hidden_omega_0=7.0, scale=6.0, scale_tensor=[1.0], pos_encode=False, sidelength=256)
    ref = O.TorchOracle("wire", 2, 300, 2, 3, 7.0, 7.0, 6.0)
    a, b = m.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
    # 91 587 trainable parameters (complex counted once), omega_0 / scale_0 without grad
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 91587
    assert not m.net[0].omega_0.requires_grad and float(m.net[0].omega_0) == 7.0 and float(m.net[0].scale_0) == 6.0
    assert m.complex and m.wavelet == "gabor" and m.pos_encode is False and len(m.net) == 4
    m.load_state_dict(b, strict=True)
    m2 = wire_b200.get_INR("wire2d", 2, 256, None, 2, 3)
    ref2 = O.TorchOracle("wire2d", 2, 256, 2, 3)
    assert list(m2.state_dict().keys()) == list(ref2.state_dict().keys())
    assert m2.net[1].linear.weight.shape == (128, 128)


def test_get_inr_signature_superset_of_the_fork():
    """modules/models.py:27-30 positional order (with scaled_hidden_features) and the drivers' keyword calls."""
    import wire_b200
    pos = wire_b200.get_INR("wire", 3, 300, 0, 3, 1, True, 20.0, 20.0, 10.0, [], False, 512, None, True)
    assert pos.net[0].linear.weight.shape == (212, 3) and len(pos.net) == 5 and float(pos.net[1].omega_0) == 20.0
    kw = wire_b200.get_INR(nonlin="wire", in_features=3, out_features=1, hidden_features=300, hidden_layers=3,
                           first_omega_0=20.0, hidden_omega_0=20.0, scale=10.0, pos_encode=False, sidelength=512)
    assert [tuple(p.shape) for p in kw.parameters()] == [tuple(p.shape) for p in pos.parameters()]
    with pytest.raises(ValueError):
        wire_b200.get_INR("siren", 2, 256, None, 2, 3)
    with pytest.raises(TypeError):
        wire_b200.get_INR("wire", 2, 256)
    direct = wire_b200.wire2d.INR(2, 256, 2, 3, True, 8.0, 8.0, 9.0)
    assert float(direct.net[2].scale_0) == 9.0


def test_no_cpu_fallback():
    import wire_b200
    m = wire_b200.get_INR("wire", 2, 64, None, 1, 3)
    with pytest.raises(wire_b200.WireB200Error, match="CUDA"):
        m(torch.zeros(1, 8, 2))
    with pytest.raises(wire_b200.WireB200Error, match="CUDA"):
        m.net[0](torch.zeros(1, 8, 2))
    with pytest.raises(wire_b200.WireB200Error):
        wire_b200._lib.load(os.path.join(ROOT, "wire_b200", "lib", "does_not_exist.so"))
    # trainable omega_0 / scale_0 run through the layer-by-layer CUDA route: still no CPU path ...
    with pytest.raises(wire_b200.WireB200Error, match="CUDA"):
        wire_b200.wire.ComplexGaborLayer(2, 8, is_first=True, trainable=True)(torch.zeros(4, 2))
    # ... the fused Trainer takes them on the mixed16 path (CUDA only), and refuses them for the other precisions instead of
    # silently freezing the scalars
    m2 = wire_b200.get_INR("wire", 2, 64, None, 1, 3, precision="tf32")
    m2.net[1].omega_0.requires_grad_(True)
    with pytest.raises(wire_b200.WireB200Error, match="mixed16"):
        wire_b200.Trainer(m2)
    m3 = wire_b200.get_INR("wire", 2, 64, None, 1, 3)
    m3.net[1].omega_0.requires_grad_(True)
    with pytest.raises(wire_b200.WireB200Error, match="CUDA"):
        wire_b200.Trainer(m3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wire_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "wire_oracle" not in src and "oracle/" not in src and "/root/reference" not in src, f


def test_patch_reference_rebinds_the_hot_path():
    import wire_b200
    pkg = types.ModuleType("fake_modules")
    pkg.__path__ = []
    sub_wire, sub_w2, sub_models = (types.ModuleType(f"fake_modules.{n}") for n in ("wire", "wire2d", "models"))
    sub_wire.INR = sub_wire.ComplexGaborLayer = sub_w2.INR = sub_w2.ComplexGaborLayer2D = object
    calls = []
    sub_models.get_INR = lambda nonlin, *a, **k: calls.append(nonlin) or "reference-model"
    sub_models.model_dict = {"wire": sub_wire, "wire2d": sub_w2, "siren": object}
    for m in (pkg, sub_wire, sub_w2, sub_models):
        sys.modules[m.__name__] = m
    try:
        patched = wire_b200.patch_reference(pkg)
        assert "fake_modules.models.get_INR" in patched and sub_wire.INR is wire_b200.wire.INR
        assert sub_w2.ComplexGaborLayer2D is wire_b200.wire2d.ComplexGaborLayer2D
        model = sub_models.get_INR(nonlin="wire", in_features=2, out_features=3, hidden_features=64, hidden_layers=1)
        assert isinstance(model, wire_b200.wire.INR)
        assert sub_models.get_INR("siren", 2, 3) == "reference-model" and calls == ["siren"]
    finally:
        for m in (pkg, sub_wire, sub_w2, sub_models):
            sys.modules.pop(m.__name__, None)


@pytest.mark.skipif(not os.path.isdir("/root/reference/modules"), reason="reference checkout only exists in the build container")
def test_patch_real_reference_package():
    import wire_b200
    sys.path.insert(0, "/root/reference")
    try:
        import modules.wire as rw
        import modules.wire2d as rw2
        import modules
        orig = rw.INR, rw.ComplexGaborLayer, rw2.INR, rw2.ComplexGaborLayer2D
        patched = wire_b200.patch_reference(modules)
        assert "modules.wire.INR" in patched and rw.INR is wire_b200.wire.INR
        rw.INR, rw.ComplexGaborLayer, rw2.INR, rw2.ComplexGaborLayer2D = orig
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
            sys.modules.pop(k, None)


def test_shard_range_partitions_exactly():
    from wire_b200.parallel import shard_range
    for n in (0, 1, 7, 200000, 134217728):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wire_b200 import parallel
    torch.manual_seed(0)
    ref = O.TorchOracle("wire", 2, 24, 1, 3, 7.0, 7.0, 6.0)  # stands in for the CUDA module: same parameters/grads
    if rank != 0:
        for p in ref.parameters():
            p.data.zero_()
    parallel.broadcast_parameters(ref)
    coords = torch.rand(1, 101, 2, generator=torch.Generator().manual_seed(1)) * 2 - 1
    target = torch.rand(1, 101, 3, generator=torch.Generator().manual_seed(2))
    lo, hi = parallel.shard_range(101, rank, world)
    params = [p for p in ref.parameters() if p.requires_grad]
    loss = ((ref(coords[:, lo:hi]) - target[:, lo:hi]) ** 2).mean()
    loss.backward()
    parallel.weighted_allreduce_gradients(params, hi - lo)
    flat = parallel.flatten_grads(params)
    assert flat.dtype == torch.float32 and flat.numel() == sum(p.numel() * (2 if p.is_complex() else 1) for p in params)
    if rank == 0:
        ret["flat"] = flat.clone()
        ret["w"] = torch.view_as_real(ref.net[1].linear.weight.data).clone()
    dist.barrier()
    dist.destroy_process_group()


def test_coordinate_sharded_dp_matches_single_process_gloo():
    """Two ranks, unequal shards: the all-reduced gradient equals the full-batch gradient."""
    world, port = 2, 29500 + (os.getpid() % 500)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, ret), nprocs=world, join=True)
    torch.manual_seed(0)
    ref = O.TorchOracle("wire", 2, 24, 1, 3, 7.0, 7.0, 6.0)
    coords = torch.rand(1, 101, 2, generator=torch.Generator().manual_seed(1)) * 2 - 1
    target = torch.rand(1, 101, 3, generator=torch.Generator().manual_seed(2))
    ((ref(coords) - target) ** 2).mean().backward()
    from wire_b200 import parallel
    full = parallel.flatten_grads([p for p in ref.parameters() if p.requires_grad])
    assert util.rel_err(ret["flat"].numpy(), full.numpy()) < 1e-5
    assert torch.equal(ret["w"], torch.view_as_real(ref.net[1].linear.weight.data))  # broadcast made ranks identical


# ---- the epoch loop's sharding logic (wire_b200.data.run_epoch) on gloo, world size 2, with stand-ins for the CUDA pieces ----
class _StubBatcher:
    def __init__(self, total):
        self.total, self.device = total, torch.device("cpu")


class _StubTrainer:
    """Records what run_epoch asks each rank to do; 'loss' = this rank's part of the chunk mean of the index values."""

    def __init__(self, world, rank):
        self.world, self.group, self.rank, self.calls = world, None, rank, []

    def step_indexed(self, batcher, idx, n_global=None, rec=None):
        self.calls.append((idx.clone(), n_global))
        if rec is not None:
            rec[idx] = idx.float().reshape(-1, 1)          # "prediction" of an index = its value
        return idx.float().sum() / float(n_global if n_global is not None else max(idx.numel(), 1))


def _epoch_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wire_b200 import data
    N, maxpoints = 1003, 250                                 # 5 chunks, the last one ragged (3 indices: unequal shards)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(9))
    tr = _StubTrainer(world, rank)
    rec = torch.full((N, 1), -1.0)
    loss = data.run_epoch(tr, _StubBatcher(N), maxpoints, indices=perm, rec=rec)
    ret[rank] = ([(c.tolist(), g) for c, g in tr.calls], rec.reshape(-1).tolist(), float(loss))
    dist.destroy_process_group()


def test_run_epoch_shards_every_chunk_exactly_gloo():
    world, port = 2, 30100 + (os.getpid() % 500)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_epoch_worker, args=(world, port, ret), nprocs=world, join=True)
    N, maxpoints = 1003, 250
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(9))
    chunks = [perm[b:min(N, b + maxpoints)] for b in range(0, N, maxpoints)]
    for ci, chunk in enumerate(chunks):
        parts = [ret[r][0][ci] for r in range(world)]
        assert all(g == chunk.numel() for _, g in parts)                       # loss normalised by the GLOBAL chunk size
        assert sum((p for p, _ in parts), []) == chunk.tolist()                  # contiguous shards, nothing lost or repeated
        assert abs(len(parts[0][0]) - len(parts[1][0])) <= 1                     # balanced
    # rec: every index predicted by exactly one rank, then summed over ranks (identical on both)
    assert ret[0][1] == ret[1][1] == [float(i) for i in range(N)]
    # this rank's loss parts add up to the epoch's mean chunk loss
    want = float(np.mean([c.float().mean().item() for c in chunks]))
    assert abs(ret[0][2] + ret[1][2] - want) < 1e-3 * want


def test_data_module_has_no_cpu_path():
    import wire_b200
    with pytest.raises(wire_b200.WireB200Error):
        wire_b200.GridBatcher((4, 4), torch.zeros(16, 3))                        # CPU signal
    with pytest.raises(wire_b200.WireB200Error):
        wire_b200.GridBatcher((4, 4), device=torch.device("cpu"))
    with pytest.raises(wire_b200.WireB200Error):
        wire_b200.data.iou_counts(torch.zeros(8), torch.zeros(8), 0.5)
    with pytest.raises(wire_b200.WireB200Error):
        wire_b200.data.psnr(torch.zeros(8), torch.zeros(8))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path: the oracle port on the host cores) prints ONE JSON line with the
    contract's keys; non-zero ranks of a torchrun launch print nothing and exit 0."""
    import json
    import subprocess
    bench = os.path.join(ROOT, "bench.py")
    # (--size 128 keeps this test short; the arm's default is the full 512 x 512 batch of the GPU arm's config)
    res = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "1", "--size", "128"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert "128x128" in d["config"]["workload"] and d["config"]["coords_per_gpu"] == 128 * 128
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "coords/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    res = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_mixed16_is_not_used_where_fp16_activations_could_overflow():
    """|y| peaks at exp(omega_0^2 / (4 scale_0^2)): models whose hyper-parameters push that beyond FP16's range fall back to the
    32-bit-operand kernels (with a warning); every driver's pair stays on the default mixed16 path."""
    import warnings
    import wire_b200
    for w0, s0 in ((7.0, 6.0), (8.0, 9.0), (20.0, 10.0), (3.0, 4.0), (30.0, 10.0)):
        assert wire_b200.get_INR("wire", 2, 64, None, 2, 3, True, w0, w0, s0).precision == "mixed16"
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        m = wire_b200.get_INR("wire", 2, 64, None, 2, 3, True, 30.0, 30.0, 4.0)
    assert m.precision == "tf32" and all(layer.precision == "tf32" for layer in list(m.net)[:-1])
    assert any("FP16" in str(w.message) for w in rec)


def _peer_possible_worker(rank, world, port, ret, fake_hosts):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import socket
    from wire_b200 import parallel
    if fake_hosts:
        socket.gethostname = lambda: f"node{rank}"
    ok, why = parallel.peer_exchange_possible(torch.device("cuda", 0), None)
    ret[rank] = (ok, why)
    dist.destroy_process_group()


def test_peer_exchange_is_only_offered_on_one_host_gloo():
    """The NVLink peer-memory exchange maps cudaIpc handles, which do not cross hosts: every rank must reach the same verdict
    (ranks on two hosts -> fall back to the all-reduce), decided collectively before anything is mapped."""
    mgr = mp.Manager()
    for fake, want in ((True, False), (False, True)):
        ret = mgr.dict()
        mp.spawn(_peer_possible_worker, args=(2, 29600 + (os.getpid() % 300) + int(fake), ret, fake), nprocs=2, join=True)
        assert ret[0] == ret[1] and ret[0][0] is want, dict(ret)
        if fake:
            assert "several hosts" in ret[0][1]
