"""GPU parity of the on-device coordinate pipeline and metrics (wire_b200/csrc/data_kernels.cuh, through the C ABI)
against the reference's own outputs (tests/golden/data_pipeline.npz) and the oracle restatements — bit-exact: this is
index / integer / rounding-exact work — plus the epoch loop built on it."""
import numpy as np
import pytest
import torch

import util
import wire_oracle as O

pytestmark = pytest.mark.gpu


def test_grid_coordinates_bit_exact_vs_reference_fixtures():
    import wire_b200
    g = util.load_data_golden()
    dev = torch.device("cuda", 0)
    for key in g.files:
        if key.startswith("coords_np_"):
            H, W, T = (int(v) for v in key.split("_")[2:])
            b = wire_b200.GridBatcher((H, W, T) if T else (H, W), linspace="numpy", device=dev)
        elif key.startswith("coords_torch_"):
            H, W = (int(v) for v in key.split("_")[2:])
            b = wire_b200.GridBatcher((H, W), linspace="torch", device=dev)
        else:
            continue
        got = b.coords().cpu().numpy()
        assert np.array_equal(got, g[key]), key


@pytest.mark.parametrize("shape,kind", [((512, 512), "torch"), ((96, 128, 80), "numpy"), ((1024, 1024), "torch"), ((37, 1, 5), "numpy")])
def test_grid_batch_gather_scatter_vs_oracle_full_size(shape, kind):
    """Random index batches at the BASELINE sizes (512^2, 1024^2, a 3-D volume): coordinates == the oracle's table rows,
    targets == signal rows, scatter == index assignment; ranges (idx=None) and duplicates included."""
    import wire_b200
    dev = torch.device("cuda", 0)
    table = O.get_coords_np(*shape) if kind == "numpy" else O.image_coords_torch(*shape)
    total = table.shape[0]
    out_f = 3 if len(shape) == 2 else 1
    gen = torch.Generator().manual_seed(1)
    signal = torch.rand(total, out_f, generator=gen)
    b = wire_b200.GridBatcher(shape, signal.to(dev), linspace=kind)
    idx = torch.randint(0, total, (min(total, 50000),), generator=gen)
    idx[:3] = torch.tensor([0, total - 1, total // 2])
    coords, target = b.assemble(idx.to(dev))
    assert np.array_equal(coords.cpu().numpy(), table[idx.numpy()])
    assert torch.equal(target.cpu(), signal[idx])
    c2, t2 = b.assemble(start=total // 3, count=min(1000, total - total // 3))
    assert np.array_equal(c2.cpu().numpy(), table[total // 3: total // 3 + c2.shape[0]])
    assert torch.equal(t2.cpu(), signal[total // 3: total // 3 + c2.shape[0]])
    perm = torch.randperm(total, generator=gen)[: min(total, 40000)]
    src = torch.rand(perm.numel(), out_f, generator=gen)
    rec = torch.zeros(total, out_f, device=dev)
    b.scatter(rec, src.to(dev), perm.to(dev))
    want = torch.zeros(total, out_f)
    want[perm] = src
    assert torch.equal(rec.cpu(), want)
    b.check_indices()
    # out-of-range indices are reported, not dereferenced
    bad = torch.tensor([0, total], dtype=torch.int64, device=dev)
    b.assemble(bad)
    with pytest.raises(wire_b200.WireB200Error):
        b.check_indices()
    # empty batch
    c0, t0 = b.assemble(torch.empty(0, dtype=torch.int64, device=dev))
    assert c0.shape == (0, len(shape)) and t0.shape == (0, out_f)


def test_iou_and_psnr_vs_reference_fixtures_and_large():
    import wire_b200
    from wire_b200 import data
    g = util.load_data_golden()
    dev = torch.device("cuda", 0)
    gt = torch.from_numpy(g["iou_gt"]).to(dev)
    for thres, inter, union in g["iou_results"]:
        p = torch.from_numpy(g["iou_preds"].copy()).to(dev)
        th = None if np.isnan(thres) else float(thres)
        c = data.iou_counts(p, gt, th)
        assert (int(c[0]), int(c[1])) == (int(inter), int(union)), thres
        if th is not None:
            assert np.array_equal(p.cpu().numpy(), g[f"iou_binarized_{thres}"])       # in place, like the reference
            p2 = torch.from_numpy(g["iou_preds"].copy()).to(dev)
            data.iou_counts(p2, gt, th, in_place=False)
            assert np.array_equal(p2.cpu().numpy(), g["iou_preds"])
    iou = data.get_IoU(torch.from_numpy(g["iou_preds"].copy()).to(dev), gt, 0.5)
    assert abs(float(iou) - float(g["iou_value_0.5"])) < 1e-7
    x, xhat = torch.from_numpy(g["psnr_x"]).to(dev), torch.from_numpy(g["psnr_xhat"]).to(dev)
    assert abs(float(data.psnr(x, xhat)) - float(g["psnr_value"])) < 1e-6
    # a 256^3 volume (16.7 M voxels): counts exact against numpy
    rs = np.random.RandomState(0)
    pv = rs.uniform(-0.2, 1.2, size=256 ** 3).astype(np.float32)
    gv = (rs.uniform(size=256 ** 3) < 0.15).astype(np.float32)
    want = O.iou_counts_np(pv.copy(), gv, 0.5)
    c = data.iou_counts(torch.from_numpy(pv).to(dev), torch.from_numpy(gv).to(dev), 0.5)
    assert (int(c[0]), int(c[1])) == want
    xv = rs.uniform(size=1 << 22).astype(np.float32)
    yv = (xv + rs.normal(scale=0.1, size=xv.shape)).astype(np.float32)
    assert abs(float(data.psnr(torch.from_numpy(xv).to(dev), torch.from_numpy(yv).to(dev))) - O.psnr_np(xv.astype(np.float64), yv.astype(np.float64))) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "mixed16"])
def test_epoch_loop_on_device_matches_reference_style_loop(precision):
    """wire_occupancy.py:136-158 on a 24x20x16 volume, 3 epochs of 4 chunks (the last one ragged): the on-device pipeline
    (run_epoch: indices -> generated coords + gathered targets -> fused step -> scatter) against the reference-style loop
    (host permutation, coords table gather, module forward, MSELoss, torch.optim.Adam) on the same permutations."""
    import wire_b200
    dev = torch.device("cuda", 0)
    H, W, T = 24, 20, 16
    N = H * W * T
    maxpoints = 2000
    rs = np.random.RandomState(3)
    vol = (rs.uniform(size=(H, W, T)) < 0.3).astype(np.float32)
    imten = torch.from_numpy(vol).reshape(N, 1).to(dev)
    coords_tab = torch.from_numpy(O.get_coords_np(H, W, T))
    kw = dict(nonlin="wire", in_features=3, hidden_features=64, hidden_layers=2, out_features=1, first_omega_0=10.0,
              hidden_omega_0=10.0, scale=5.0, precision=precision)
    torch.manual_seed(0)
    m_a = wire_b200.get_INR(**kw).to(dev)
    m_b = wire_b200.get_INR(**kw).to(dev)
    m_b.load_state_dict(m_a.state_dict())
    perms = [torch.randperm(N, generator=torch.Generator().manual_seed(10 + e)) for e in range(3)]

    # reference-style loop on the module API
    opt = torch.optim.Adam(m_a.parameters(), lr=5e-3)
    est_a = torch.zeros(N, 1, device=dev)
    losses_a = []
    for e in range(3):
        indices = perms[e]
        tot, nch = 0.0, 0
        for b_idx in range(0, N, maxpoints):
            b_indices = indices[b_idx:min(N, b_idx + maxpoints)]
            b_coords = coords_tab[b_indices, ...].to(dev)
            b_indices = b_indices.to(dev)
            pix = m_a(b_coords[None, ...]).squeeze()[:, None]
            with torch.no_grad():
                est_a[b_indices, :] = pix
            loss = torch.nn.functional.mse_loss(pix, imten[b_indices, :])
            opt.zero_grad(); loss.backward(); opt.step()
            tot += float(loss); nch += 1
        losses_a.append(tot / nch)

    # on-device pipeline
    tr = wire_b200.Trainer(m_b, lr=5e-3)
    batcher = wire_b200.GridBatcher((H, W, T), imten, linspace="numpy")
    est_b = torch.zeros(N, 1, device=dev)
    losses_b = [float(wire_b200.run_epoch(tr, batcher, maxpoints, indices=perms[e], rec=est_b)) for e in range(3)]
    batcher.check_indices()
    assert tr.steps_done == 3 * 4
    tol = 1e-3 if precision == "fp32" else 2e-2
    for a, b in zip(losses_a, losses_b):
        assert abs(a - b) <= tol * max(abs(a), 1e-3), (losses_a, losses_b)
    if precision == "fp32":
        assert float((est_a - est_b).abs().max()) <= 5e-3
    else:  # 12 Adam steps amplify the 16-bit rounding differences between two runs on a few voxels: bound the RMS
        assert float((est_a - est_b).pow(2).mean().sqrt()) <= 0.08
    from wire_b200 import data
    iou_a = float(data.get_IoU(est_a.clone(), imten, 0.5))
    iou_b = float(data.get_IoU(est_b.clone(), imten, 0.5))
    assert abs(iou_a - iou_b) <= 0.005     # north_star's IoU bar


@pytest.mark.parametrize("H,W,scale", [(64, 48, 4), (50, 37, 4), (1024, 1024, 4), (33, 33, 16)])
def test_avgpool_mse_loss_grad_vs_torch_autograd(H, W, scale):
    """wire_avgpool_mse_loss_grad == autograd of the reference's SISR loss (wire_SISR.py:151-161), including image sizes
    that AvgPool2d truncates."""
    import ctypes
    from wire_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(H * W)
    rec_hr = torch.rand(1, H * W, 3, device=dev, requires_grad=True)
    H2, W2 = H // scale, W // scale
    gt_lr = torch.rand(1, H2 * W2, 3, device=dev)
    rec = torch.nn.AvgPool2d(scale)(rec_hr.reshape(H, W, 3).permute(2, 0, 1)[None, ...])
    loss = ((gt_lr - rec.reshape(1, 3, -1).permute(0, 2, 1)) ** 2).mean()
    loss.backward()
    grad = torch.full((H * W, 3), 7.0, device=dev)
    loss_dev = torch.zeros(1, device=dev)
    _lib.check(lib.wire_avgpool_mse_loss_grad(rec_hr.data_ptr(), gt_lr.data_ptr(), H, W, 3, scale, grad.data_ptr(), loss_dev.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "wire_avgpool_mse_loss_grad")
    assert abs(float(loss_dev) - float(loss)) <= 2e-6 * max(1.0, float(loss))
    assert torch.allclose(grad, rec_hr.grad[0], rtol=1e-5, atol=1e-10)


def test_trainer_sisr_loss_matches_reference_style_loop():
    """wire2d 4x super-resolution (BASELINE config [2], scaled down to 96x96): Trainer with the fused pooled loss against
    model(coords_hr) -> AvgPool2d -> MSE -> backward -> torch.optim.Adam (wire_SISR.py:154-177), FP32 kernels."""
    import wire_b200
    dev = torch.device("cuda", 0)
    H = W = 96
    scale = 4
    x = torch.linspace(-1, 1, W); y = torch.linspace(-1, 1, H)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    coords_hr = torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None, ...].to(dev)
    torch.manual_seed(2)
    gt_lr = torch.rand(1, (H // scale) * (W // scale), 3, device=dev)
    kw = dict(nonlin="wire2d", in_features=2, hidden_features=64, hidden_layers=2, out_features=3, first_omega_0=8.0,
              hidden_omega_0=8.0, scale=9.0, precision="fp32")
    a = wire_b200.get_INR(**kw).to(dev)
    b = wire_b200.get_INR(**kw).to(dev)
    b.load_state_dict(a.state_dict())
    opt = torch.optim.Adam(a.parameters(), lr=5e-3)
    pool = torch.nn.AvgPool2d(scale)
    ref = []
    for _ in range(10):
        rec_hr = a(coords_hr)
        rec = pool(rec_hr.reshape(H, W, 3).permute(2, 0, 1)[None, ...])
        loss = ((gt_lr - rec.reshape(1, 3, -1).permute(0, 2, 1)) ** 2).mean()
        opt.zero_grad(); loss.backward(); opt.step()
        ref.append(float(loss))
    tr = wire_b200.Trainer(b, lr=5e-3)
    tr.set_loss_avgpool(H, W, scale)
    got = [float(tr.step(coords_hr, gt_lr)) for _ in range(10)]
    assert util.rel_err(np.array(got), np.array(ref)) < 2e-3, (got, ref)
    with torch.no_grad():
        assert util.rel_err(b(coords_hr).cpu().numpy(), a(coords_hr).cpu().numpy()) < 2e-2


@pytest.mark.parametrize("precision", ["mixed16", "tf32", "fp32"])
def test_sisr_iteration_shares_one_forward(precision):
    """wire_SISR.py:154-177 runs the model twice per iteration on the same coordinates with the same weights (a grad forward for
    the loss, a ``no_grad`` forward for the metrics) — the second reproduces the first bit for bit, so ``Trainer.step_sisr``
    serves both from one forward.  Checked here against REAL second forwards: (a) the training forward and the no_grad forward
    of the CUDA module are bit-identical; (b) step_sisr's rec_hr / mse_hr equal what the reference-style iteration computes."""
    import wire_b200
    dev = torch.device("cuda", 0)
    H = W = 128
    scale = 4
    x = torch.linspace(-1, 1, W); y = torch.linspace(-1, 1, H)
    X, Y = torch.meshgrid(x, y, indexing="xy")
    coords_hr = torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None, ...].to(dev)
    torch.manual_seed(3)
    gt = torch.rand(1, H * W, 3, device=dev)
    gt_lr = torch.nn.AvgPool2d(scale)(gt.reshape(H, W, 3).permute(2, 0, 1)[None, ...]).reshape(1, 3, -1).permute(0, 2, 1).contiguous()
    kw = dict(nonlin="wire2d", in_features=2, hidden_features=256, hidden_layers=2, out_features=3, first_omega_0=8.0,
              hidden_omega_0=8.0, scale=9.0, precision=precision)
    a = wire_b200.get_INR(**kw).to(dev)
    b = wire_b200.get_INR(**kw).to(dev)
    b.load_state_dict(a.state_dict())
    # (a) grad forward == no_grad forward, bit for bit
    rec_train = a(coords_hr)
    with torch.no_grad():
        rec_eval = a(coords_hr)
    assert rec_train.requires_grad and not rec_eval.requires_grad
    assert torch.equal(rec_train.detach(), rec_eval)
    # (b) the reference-style iteration against step_sisr
    opt = torch.optim.Adam(a.parameters(), lr=5e-3)
    pool = torch.nn.AvgPool2d(scale)
    tr = wire_b200.Trainer(b, lr=5e-3)
    tr.set_loss_avgpool(H, W, scale)
    for it in range(4):
        rec_hr = a(coords_hr)
        rec = pool(rec_hr.reshape(H, W, 3).permute(2, 0, 1)[None, ...])
        loss = ((gt_lr - rec.reshape(1, 3, -1).permute(0, 2, 1)) ** 2).mean()
        with torch.no_grad():
            rec_hr2 = a(coords_hr)
            mse_ref = float(((gt - rec_hr2) ** 2).mean())
        opt.zero_grad(); loss.backward(); opt.step()
        loss_b, rec_b, mse_b = tr.step_sisr(coords_hr, gt_lr, gt)
        tol = 1e-4 if precision == "fp32" else 2e-2
        assert abs(float(loss_b) - float(loss)) <= tol * max(float(loss), 1e-6), (it, float(loss_b), float(loss))
        assert abs(float(mse_b) - mse_ref) <= tol * mse_ref, (it, float(mse_b), mse_ref)
        assert util.rel_err(rec_b.cpu().numpy(), rec_hr2.reshape(-1, 3).cpu().numpy()) < (1e-4 if precision == "fp32" else 3e-2)


def test_trainer_step_from_pinned_host_buffers_is_pipelined_and_correct():
    """Trainer.step with PINNED HOST inputs (copy stream + two staging buffers) follows the same trajectory as with device
    inputs, also when the host buffers change every step."""
    import wire_b200
    dev = torch.device("cuda", 0)
    kw = dict(nonlin="wire", in_features=2, hidden_features=100, hidden_layers=2, out_features=3, first_omega_0=7.0,
              hidden_omega_0=7.0, scale=6.0, precision="fp32")
    a = wire_b200.get_INR(**kw).to(dev)
    b = wire_b200.get_INR(**kw).to(dev)
    b.load_state_dict(a.state_dict())
    gen = torch.Generator().manual_seed(0)
    batches = [((torch.rand(1, 4096, 2, generator=gen) * 2 - 1), torch.rand(1, 4096, 3, generator=gen)) for _ in range(6)]
    ta, tb = wire_b200.Trainer(a, lr=5e-3), wire_b200.Trainer(b, lr=5e-3)
    la = [float(ta.step(c.to(dev), t.to(dev))) for c, t in batches]
    pinned = [(c.pin_memory(), t.pin_memory()) for c, t in batches]
    lb = [tb.step(c, t).clone() for c, t in pinned]       # no host sync between steps
    lb = [float(v) for v in lb]
    assert util.rel_err(np.array(lb), np.array(la)) < 1e-3, (la, lb)


def test_empty_shard_step_keeps_the_trainer_consistent():
    """A rank whose shard of a chunk is empty (chunk smaller than the world) still takes an optimiser step with zero
    gradients; the loss ring and the per-size states stay consistent for the steps that follow."""
    import wire_b200
    dev = torch.device("cuda", 0)
    H, W = 32, 24
    gen = torch.Generator().manual_seed(4)
    signal = torch.rand(H * W, 3, generator=gen).to(dev)
    batcher = wire_b200.GridBatcher((H, W), signal, linspace="torch")
    model = wire_b200.get_INR(nonlin="wire", in_features=2, hidden_features=64, hidden_layers=2, out_features=3, first_omega_0=7.0,
                              hidden_omega_0=7.0, scale=6.0, precision="fp32").to(dev)
    tr = wire_b200.Trainer(model, lr=5e-3, graph=False)
    idx = torch.randperm(H * W, generator=gen).to(dev)
    l0 = float(tr.step_indexed(batcher, idx[:300], n_global=600))          # this rank holds half of a 600-coordinate chunk
    before = tr.flat.clone()
    le = float(tr.step_indexed(batcher, idx[:0], n_global=5))              # empty shard
    assert le == 0.0 and tr.steps_done == 2
    assert not torch.equal(tr.flat, before)                                 # Adam still moved the parameters (momentum)
    l1 = float(tr.step_indexed(batcher, idx[:300], n_global=600))          # same key as the first step: state must be intact
    assert np.isfinite(l1) and 0.2 * l0 < l1 < 1.5 * l0
    assert tr._n_global == 600 and tr.steps_done == 3
