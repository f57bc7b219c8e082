"""The oracle is only trusted after it reproduces the reference's own outputs (tests/golden/*.npz were
produced by oracle/make_golden.py running /root/reference/modules/{wire,wire2d}.py in c64 and c128)."""
import os
import numpy as np
import pytest
import torch

import util
import wire_oracle as O


@pytest.mark.parametrize("name", util.golden_cases())
def test_torch_oracle_matches_reference_c64(name):
    c = util.load_golden(name)
    g = c["g"]
    m = util.oracle_model(c)
    out, grads, gc = util.run_oracle(m, torch.from_numpy(g["coords"]), torch.from_numpy(g["grad_out"]))
    # same op sequence, same torch build: agree to float32 round-off
    assert util.rel_err(out.numpy(), g["out_c64"]) < 1e-6
    assert util.rel_err(gc.numpy(), g["gcoords_c64"]) < 1e-5
    for k, v in grads.items():
        a, b = util.golden_grad(c, "c64", k, v.numpy())
        assert util.rel_err(a, b) < 1e-5, k


@pytest.mark.parametrize("name", util.golden_cases())
def test_torch_oracle_matches_reference_c128(name):
    c = util.load_golden(name)
    g = c["g"]
    m = util.oracle_model(c, torch.complex128)
    out, grads, gc = util.run_oracle(m, torch.from_numpy(g["coords"]).double(), torch.from_numpy(g["grad_out"]).double())
    assert util.rel_err(out.numpy(), g["out_c128"]) < 1e-13
    for k, v in grads.items():
        a, b = util.golden_grad(c, "c128", k, v.numpy())
        assert util.rel_err(a, b) < 1e-12, k


@pytest.mark.parametrize("name", util.golden_cases())
def test_closed_form_matches_reference_c128(name):
    """Autograd-free NumPy closed form (Wirtinger convention, appendix A.2) vs the reference's autograd."""
    c = util.load_golden(name)
    g = c["g"]
    m = util.oracle_model(c)
    state = {k: v.numpy().astype(np.complex128 if v.is_complex() else np.float64) for k, v in m.state_dict().items()}
    out = O.forward_np(state, g["coords"])
    assert util.rel_err(out, g["out_c128"]) < 1e-12
    grads = O.backward_np(state, g["coords"], g["grad_out"])
    assert util.rel_err(grads["coords"], g["gcoords_c128"]) < 1e-11
    for k in (k for k in g.files if k.startswith("grad_c128.")):
        key = k[len("grad_c128."):]
        a, b = util.golden_grad(c, "c128", key, grads[key])
        assert util.rel_err(a, b) < 1e-11, key
    # final-layer bias gradient is exactly real
    last = max(int(k.split(".")[1]) for k in state)
    assert np.all(grads[f"net.{last}.bias"].imag == 0)


@pytest.mark.parametrize("name", util.golden_cases())
def test_single_stage_closed_forms_compose_to_the_reference(name):
    """layer_forward_np / gabor_backward_np / linear_backward_np (what the per-kernel GPU parity tests check each CUDA kernel
    against) chained over the stack reproduce the reference's complex128 outputs and gradients."""
    c = util.load_golden(name)
    g = c["g"]
    m = util.oracle_model(c)
    state = {k: v.numpy().astype(np.complex128 if v.is_complex() else np.float64) for k, v in m.state_dict().items()}
    layers, final = O._layers_from_state(state)
    x = g["coords"].reshape(-1, c["in_f"]).astype(np.float64)
    saved = []
    for i, L in enumerate(layers):
        z, w, y = O.layer_forward_np(L, x)
        if f"layer{i}_c128" in g.files:   # the big fixtures keep only outputs and (sub-sampled) gradients
            assert util.rel_err(y, g[f"layer{i}_c128"].reshape(y.shape)) < 1e-12, i
        saved.append((x, z, w))
        x = y
    out = (x @ final["W"].T + final["b"]).real
    assert util.rel_err(out, g["out_c128"].reshape(out.shape)) < 1e-12
    g_o = g["grad_out"].reshape(out.shape).astype(np.complex128)
    g_y, g_Wf, g_bf = O.linear_backward_np(final["W"], x, g_o)
    idx = len(layers)
    got = {f"net.{idx}.weight": g_Wf, f"net.{idx}.bias": g_bf}
    for i in range(len(layers) - 1, -1, -1):
        L = layers[i]
        xin, z, w = saved[i]
        g_z, g_w = O.gabor_backward_np(L, z, w, g_y)
        g_y, got[f"net.{i}.linear.weight"], got[f"net.{i}.linear.bias"] = O.linear_backward_np(L["W"], xin, g_z)
        if g_w is not None:
            gx2, got[f"net.{i}.scale_orth.weight"], got[f"net.{i}.scale_orth.bias"] = O.linear_backward_np(L["W2"], xin, g_w)
            g_y = g_y + gx2
    assert util.rel_err(np.real(g_y), g["gcoords_c128"].reshape(g_y.shape)) < 1e-11
    for k in (k for k in g.files if k.startswith("grad_c128.")):
        key = k[len("grad_c128."):]
        a, b = util.golden_grad(c, "c128", key, got[key])
        assert util.rel_err(a, b) < 1e-11, key


def test_layer_outputs_match_reference():
    c = util.load_golden("wire_small")
    g = c["g"]
    m = util.oracle_model(c)
    x = torch.from_numpy(g["coords"])
    for i, layer in enumerate(m.net):
        x = layer(x)
        assert util.rel_err(x.detach().numpy(), g[f"layer{i}_c64"]) < 1e-6


def test_stored_params_equal_deterministic_state():
    c = util.load_golden("wire2d_small")
    m = util.oracle_model(c)
    for k, v in m.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), c["g"][f"param.{k}"])


def test_metrics():
    x = np.linspace(0, 1, 64).reshape(8, 8)
    assert abs(O.psnr(x, x + 0.1) - 10 * np.log10(1.0 / 0.01)) < 1e-9
    a = np.zeros((4, 4)); b = np.zeros((4, 4)); a[:2] = 1; b[1:3] = 1
    assert abs(O.iou(a, b) - 4 / 12) < 1e-12
    assert tuple(O.image_coords(4, 6).shape) == (1, 24, 2)
    assert tuple(O.volume_coords(2, 3, 4).shape) == (24, 3)


@pytest.mark.parametrize("name", ["wire_small", "wire_occ_small", "wire2d_small", "wire_odd_width"])
def test_c_oracle_matches_reference_c128(name):
    """oracle/wire_oracle.c (plain C, double) vs the reference's complex128 run."""
    import c_oracle
    if not c_oracle.available():
        pytest.skip("oracle/_build/libwire_oracle_c.so not built (run __graft_entry__.build() or make -C oracle)")
    c = util.load_golden(name)
    g = c["g"]
    m = util.oracle_model(c)
    state = {k: v.numpy() for k, v in m.state_dict().items()}
    out, grads, gc = c_oracle.run(state, c["kind"] == "wire2d", g["coords"], g["grad_out"])
    assert util.rel_err(out.reshape(g["out_c128"].shape), g["out_c128"]) < 1e-12
    assert util.rel_err(gc.reshape(g["gcoords_c128"].shape), g["gcoords_c128"]) < 1e-11
    for k in (k for k in g.files if k.startswith("grad_c128.")):
        key = k[len("grad_c128."):]
        assert util.rel_err(grads[key], g[k]) < 1e-11, key


# ---- data pipeline restatements (oracle/wire_oracle.py) vs the reference's own outputs (tests/golden/data_pipeline.npz) ----
def test_data_pipeline_oracle_matches_reference_fixtures():
    g = util.load_data_golden()
    n_coords = 0
    for key in g.files:
        if key.startswith("coords_np_"):
            H, W, T = (int(v) for v in key.split("_")[2:])
            got = O.get_coords_np(H, W, T or None)
        elif key.startswith("coords_torch_"):
            H, W = (int(v) for v in key.split("_")[2:])
            got = O.image_coords_torch(H, W)
        else:
            continue
        assert got.dtype == np.float32 and np.array_equal(got, g[key]), key   # bit-exact
        n_coords += 1
    assert n_coords >= 9
    for thres, inter, union in g["iou_results"]:
        p = g["iou_preds"].copy()
        i, u = O.iou_counts_np(p, g["iou_gt"], None if np.isnan(thres) else float(thres))
        assert (i, u) == (int(inter), int(union))
        if not np.isnan(thres):
            assert np.array_equal(p, g[f"iou_binarized_{thres}"])             # in-place thresholding, as the reference
    assert abs(O.psnr_np(g["psnr_x"], g["psnr_xhat"]) - float(g["psnr_value"])) < 1e-12


def test_trainable_scalar_grads_closed_form_matches_reference_autograd():
    g = np.load(os.path.join(util.GOLDEN_DIR, "trainable_scalars.npz"))
    for tag in ("wire_first", "wire_hidden", "wire2d_first", "wire2d_hidden"):
        w0, s0 = (float(v) for v in g[f"{tag}.hyper"])
        two_d = tag.startswith("wire2d")
        y, g_om, g_s0 = O.gabor_scalar_grads_np(g[f"{tag}.x"], g[f"{tag}.param.linear.weight"], g[f"{tag}.param.linear.bias"],
                                                 g[f"{tag}.param.scale_orth.weight"] if two_d else None,
                                                 g[f"{tag}.param.scale_orth.bias"] if two_d else None, w0, s0, g[f"{tag}.gy"])
        assert util.rel_err(y, g[f"{tag}.y_c128"]) < 1e-12
        assert abs(g_om - float(g[f"{tag}.g_omega_c128"][0])) <= 1e-10 * max(1.0, abs(g_om)), tag
        assert abs(g_s0 - float(g[f"{tag}.g_scale_c128"][0])) <= 1e-10 * max(1.0, abs(g_s0)), tag


def test_real_gabor_restatement_matches_reference_fixture():
    g = np.load(os.path.join(util.GOLDEN_DIR, "real_gabor.npz"))
    w0, s0 = (float(v) for v in g["hyper"])
    y = O.real_gabor_np(g["x"], g["param.freqs.weight"], g["param.freqs.bias"], g["param.scale.weight"], g["param.scale.bias"], w0, s0)
    assert util.rel_err(y, g["y_f64"]) < 1e-6      # (the fixture's parameters are float32 values)


def test_radon_restatement_is_self_consistent():
    """The Radon oracle (parity unpinned: kornia absent, see oracle/wire_oracle.py): affine_grid / grid_sample form == direct
    float64 loops; angle 0 is the plain column sum, 180 degrees the mirrored one, 90 degrees of a square image the row sums."""
    rs = np.random.RandomState(0)
    im = rs.uniform(size=(9, 12))
    ang = [0.0, 17.0, 90.0, 133.5, 180.0, -45.0]
    a = O.radon_torch(torch.from_numpy(im)[None, None], torch.tensor(ang, dtype=torch.float64)).numpy()
    b = O.radon_np(im, ang)
    assert a.shape == (len(ang), 12) and np.abs(a - b).max() < 1e-12
    assert np.abs(a[0] - im.sum(0)).max() < 1e-12
    assert np.abs(a[4] - im.sum(0)[::-1]).max() < 1e-9
    sq = rs.uniform(size=(8, 8))
    r = O.radon_np(sq, [90.0])[0]
    assert np.abs(r - sq.sum(1)).max() < 1e-9 or np.abs(r - sq.sum(1)[::-1]).max() < 1e-9
    # is_3d: (1, nimg, H, W) -> (nimg, nangles, W)
    vol = torch.from_numpy(rs.uniform(size=(1, 3, 8, 8)))
    s3 = O.radon_torch(vol, torch.tensor([0.0, 30.0], dtype=torch.float64), is_3d=True)
    assert tuple(s3.shape) == (3, 2, 8)
    assert np.abs(s3[1, 0].numpy() - vol[0, 1].numpy().sum(0)).max() < 1e-12
