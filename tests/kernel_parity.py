"""Per-kernel parity of the fused whole-network path (test infrastructure).

Every CUDA kernel of a training step is checked FROM ITS OWN INPUTS: the tensors the step left in its workspace
(C ABI ``wire_net_workspace_read``: FP16 activations / pre-activations and BF16 gradients under ``mixed16``) are fed, stage
by stage, through the oracle's complex128 closed forms (``oracle/wire_oracle.py``: ``layer_forward_np`` =
modules/wire.py:88-93 / modules/wire2d.py:56-67, ``gabor_backward_np`` / ``linear_backward_np`` = their autograd,
SURVEY.md appendix A.2, pinned to the reference by tests/test_oracle.py), and each kernel's output is compared with what the
oracle makes of the same input.  This is the north-star's "per-layer outputs and gradients from identical inputs" bar for
the kernels that produce the benchmarked number (tc_rows16 / first_fwd16 / top_bwd16 / OP16 tc_wgrad / first_wgrad16), and
the same harness runs the tf32 and fp32 kernels.
"""
import numpy as np
import torch

import util
import wire_oracle as O


def np_state(ref):
    return {k: v.detach().cpu().numpy().astype(np.complex128 if v.is_complex() else np.float64) for k, v in ref.state_dict().items()}


def kernel_errors(model, ref, coords, grad_out):
    """Runs one forward + backward of `model` (wire_b200 INR on the GPU) and returns {stage: relative RMS error} of every
    kernel output against the oracle evaluated on that kernel's own inputs.  `ref`: TorchOracle with the same weights."""
    from wire_b200 import functional as F
    state = np_state(ref)
    layers, final = O._layers_from_state(state)
    H = len(layers) - 1
    two_d = layers[0]["W2"] is not None
    cg = coords.cuda().requires_grad_(True)
    out = model(cg)
    ctx = out.grad_fn
    (out * grad_out.cuda()).sum().backward()
    torch.cuda.synchronize()
    desc, ws = ctx.desc, ctx.lease.buf
    n = coords.numel() // coords.shape[-1]

    def rd(which, index=0):
        return F.workspace_read(desc, n, ws, which, index).cpu().numpy().astype(np.complex128 if which not in ("gz0", "gw0") else np.float64)

    err = {}
    c64 = coords.reshape(n, -1).numpy().astype(np.float64)
    go = grad_out.reshape(n, -1).numpy().astype(np.float64)
    out_np = out.detach().cpu().numpy().reshape(n, -1)
    # ---------------- forward ----------------
    z0, w0, y0_ref = O.layer_forward_np(layers[0], c64)
    y_prev = rd("y", 0)
    err["fwd0.y (first_fwd)"] = util.rel_err(y_prev, y0_ref)
    zs, ws_ = {}, {}
    for l in range(1, H + 1):
        z_ref, w_ref, y_ref = O.layer_forward_np(layers[l], y_prev)
        zs[l] = rd("z", l)
        err[f"fwd{l}.z (rows_gabor_fwd)"] = util.rel_err(zs[l], z_ref)
        if two_d:
            ws_[l] = rd("w", l)
            err[f"fwd{l}.w (rows_gabor_fwd)"] = util.rel_err(ws_[l], w_ref)
        if l < H:
            y_prev = rd("y", l)
            err[f"fwd{l}.y (rows_gabor_fwd)"] = util.rel_err(y_prev, y_ref)
        else:
            o_ref = (y_ref @ final["W"].T + final["b"]).real
            err[f"fwd{l}.out (fused final Linear)"] = util.rel_err(out_np, o_ref)
    # ---------------- backward ----------------
    grads = {k: (torch.view_as_real(p.grad).cpu().numpy() if p.grad.is_complex() else p.grad.cpu().numpy())
             for k, p in model.named_parameters() if p.grad is not None}

    def cplx(a):
        return a[..., 0].astype(np.complex128) + 1j * a[..., 1]

    def slot(l):  # g_z of hidden layer l after a complete backward pass (the two buffers alternate)
        return (H - l) & 1 if l <= 2 else None

    # top of the backward pass: inputs g_out, saved z_H (w_H), W_f
    y_H = O.gabor_np(zs[H], ws_.get(H), layers[H]["omega"], layers[H]["scale"])
    g_y, g_Wf, g_bf = O.linear_backward_np(final["W"], y_H, go.astype(np.complex128))
    idx = H + 1
    err["top.g_Wf (top_bwd)"] = util.rel_err(cplx(grads[f"net.{idx}.weight"]), g_Wf)
    err["top.g_bf (top_bwd)"] = util.rel_err(cplx(grads[f"net.{idx}.bias"]), g_bf)
    gz_ref, gw_ref = O.gabor_backward_np(layers[H], zs[H], ws_.get(H), g_y)
    gz = {}
    gw = {}
    if slot(H) is not None:
        gz[H] = rd("gz", slot(H))
        err["top.g_z (top_bwd)"] = util.rel_err(gz[H], gz_ref)
        if two_d:
            gw[H] = rd("gw", slot(H))
            err["top.g_w (top_bwd)"] = util.rel_err(gw[H], gw_ref)
    gz0 = rd("gz0")
    gw0 = rd("gw0") if two_d else None
    for l in range(H, 0, -1):
        if l not in gz:
            if slot(l) is None:
                continue
            gz[l] = rd("gz", slot(l))
            if two_d:
                gw[l] = rd("gw", slot(l))
        x_in = rd("y", l - 1)
        # weight gradient GEMM: inputs y_{l-1}, g_z(l)
        _, gW_ref, gb_ref = O.linear_backward_np(layers[l]["W"], x_in, gz[l])
        err[f"wgrad{l}.g_W (tc_wgrad)"] = util.rel_err(cplx(grads[f"net.{l}.linear.weight"]), gW_ref)
        err[f"wgrad{l}.g_b (tc_wgrad)"] = util.rel_err(cplx(grads[f"net.{l}.linear.bias"]), gb_ref)
        g_x = gz[l] @ np.conj(layers[l]["W"])
        if two_d:
            _, gW2_ref, gb2_ref = O.linear_backward_np(layers[l]["W2"], x_in, gw[l])
            err[f"wgrad{l}.g_W2 (tc_wgrad)"] = util.rel_err(cplx(grads[f"net.{l}.scale_orth.weight"]), gW2_ref)
            err[f"wgrad{l}.g_b2 (tc_wgrad)"] = util.rel_err(cplx(grads[f"net.{l}.scale_orth.bias"]), gb2_ref)
            g_x = g_x + gw[l] @ np.conj(layers[l]["W2"])
        # dgrad fused with the Gabor backward of the layer below: inputs g_z(l), W_l, saved z_{l-1} (coords for the first layer)
        if l - 1 >= 1:
            if slot(l - 1) is None:
                continue
            gzb_ref, gwb_ref = O.gabor_backward_np(layers[l - 1], zs[l - 1], ws_.get(l - 1), g_x)
            gz[l - 1] = rd("gz", slot(l - 1))
            err[f"dgrad{l}.g_z{l - 1} (rows_dgrad_gabor_bwd)"] = util.rel_err(gz[l - 1], gzb_ref)
            if two_d:
                gw[l - 1] = rd("gw", slot(l - 1))
                err[f"dgrad{l}.g_w{l - 1} (rows_dgrad_gabor_bwd)"] = util.rel_err(gw[l - 1], gwb_ref)
        else:
            gz0_ref, gw0_ref = O.gabor_backward_np(layers[0], z0, w0, g_x)
            err["dgrad1.g_z0 (rows_dgrad_first_bwd)"] = util.rel_err(gz0, gz0_ref)
            if two_d:
                err["dgrad1.g_w0 (rows_dgrad_first_bwd)"] = util.rel_err(gw0, gw0_ref)
    # first-layer weight gradient and coordinate gradient: inputs g_z0 (g_w0), coords, W0
    gc_ref, gW0_ref, gb0_ref = O.linear_backward_np(layers[0]["W"], c64, gz0)
    err["wgrad0.g_W (first_wgrad)"] = util.rel_err(grads["net.0.linear.weight"], gW0_ref)
    err["wgrad0.g_b (first_wgrad)"] = util.rel_err(grads["net.0.linear.bias"], gb0_ref)
    if two_d:
        gc2, gW0b_ref, gb0b_ref = O.linear_backward_np(layers[0]["W2"], c64, gw0)
        err["wgrad0.g_W2 (first_wgrad)"] = util.rel_err(grads["net.0.scale_orth.weight"], gW0b_ref)
        err["wgrad0.g_b2 (first_wgrad)"] = util.rel_err(grads["net.0.scale_orth.bias"], gb0b_ref)
        gc_ref = gc_ref + gc2
    err["grad_coords"] = util.rel_err(cg.grad.cpu().numpy().reshape(n, -1), gc_ref)
    return err


# the three network shapes of BASELINE.json configs [1] (denoise), [2] (wire2d SISR) and [3] (occupancy; H = 2 keeps the top
# kernel's g_z readable, the H = 3 variant covers the three-hidden-layer chain without it)
CASES = {
    "denoise": dict(kind="wire", in_f=2, hidden=300, H=2, out_f=3, w0=7.0, w0h=7.0, s0=6.0),
    "sisr2d": dict(kind="wire2d", in_f=2, hidden=256, H=2, out_f=3, w0=8.0, w0h=8.0, s0=9.0),
    "occupancy_h2": dict(kind="wire", in_f=3, hidden=300, H=2, out_f=1, w0=20.0, w0h=20.0, s0=10.0),
    "occupancy": dict(kind="wire", in_f=3, hidden=300, H=3, out_f=1, w0=20.0, w0h=20.0, s0=10.0),
}


def build_case(name, precision, seed=7):
    import wire_b200
    c = CASES[name]
    ref = O.TorchOracle(c["kind"], c["in_f"], c["hidden"], c["H"], c["out_f"], c["w0"], c["w0h"], c["s0"])
    ref.load_state_dict(O.deterministic_state(ref, seed), strict=True)
    m = wire_b200.get_INR(nonlin=c["kind"], in_features=c["in_f"], hidden_features=c["hidden"], hidden_layers=c["H"],
                          out_features=c["out_f"], first_omega_0=c["w0"], hidden_omega_0=c["w0h"], scale=c["s0"],
                          precision=precision)
    m.load_state_dict(ref.state_dict(), strict=True)
    return m.cuda(), ref, c
