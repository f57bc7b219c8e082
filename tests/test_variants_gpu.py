"""Kernel variants must agree with each other.

Several stages of the mixed16 step exist in two implementations selected by shape, with an environment switch that forces the
older one: the bias gradient of `tc_wgrad` from the "ones" column or summed by the converter warps
(+ both g tensors of a wire2d layer in one work item), the first-layer weight gradient streamed through a shared-memory ring or
register-pipelined, stored tensors ending at column 2M or on the next 32-byte sector boundary.  Each pair computes the same sums
from the same 16-bit tensors, so their gradients may differ only by the order of FP32 additions: asserted here to 2e-5 relative
(measured <= 2.1e-6, tools/ab_grad_check.py), at row counts that are ragged against every tile size and large enough for the
streamed kernel (>= 65 536 rows).  The absolute accuracy of either variant is the business of test_kernel_parity_gpu.py.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # (switch, nonlin, hidden_features, in_features, out_features, hidden_layers)
    ("WIRE_B200_SECTOR_ALIGN", "wire", 300, 2, 3, 2),     # M = 212: 2M = 424 columns -> stored width 432
    ("WIRE_B200_SECTOR_ALIGN", "wire", 300, 3, 1, 3),     # the occupancy network
    ("WIRE_B200_FWGRAD_STREAM", "wire", 300, 2, 3, 2),    # full-width g_z0 rows (32 lanes per row)
    ("WIRE_B200_FWGRAD_STREAM", "wire2d", 256, 2, 3, 2),  # M = 128: 256-byte rows, two to a warp, 256-row chunks
    ("WIRE_B200_FWGRAD_STREAM", "wire", 90, 3, 1, 2),     # M = 63: 8 lanes per row, ragged last feature quad
    ("WIRE_B200_FWGRAD_BULK1D", "wire", 300, 2, 3, 2),
    ("WIRE_B200_BIAS_SUM", "wire2d", 256, 2, 3, 2),       # 2K = 256: pair + dual work items against three single-CTA x tiles
    ("WIRE_B200_BIAS_SUM", "wire2d", 128, 2, 3, 2),       # 2K = 128: single CTAs
    ("WIRE_B200_WGRAD_DUAL", "wire2d", 256, 2, 3, 2),
]


@pytest.mark.parametrize("switch,nonlin,hidden,in_f,out_f,layers", CASES)
def test_variants_agree(switch, nonlin, hidden, in_f, out_f, layers):
    import wire_b200
    n = 70001
    torch.manual_seed(0)
    m = wire_b200.get_INR(nonlin=nonlin, in_features=in_f, hidden_features=hidden, hidden_layers=layers, out_features=out_f,
                          first_omega_0=8.0, hidden_omega_0=8.0, scale=9.0, precision="mixed16").cuda()
    c = torch.rand(1, n, in_f, device="cuda") * 2 - 1
    g = torch.randn(1, n, out_f, device="cuda") / n
    saved = os.environ.get(switch)
    res = {}
    try:
        for flag in ("0", "1"):
            os.environ[switch] = flag
            m.zero_grad()
            out = m(c)
            (out * g).sum().backward()
            res[flag] = ({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, out.detach().clone())
    finally:
        if saved is None:
            os.environ.pop(switch, None)
        else:
            os.environ[switch] = saved
    (ga, oa), (gb, ob) = res["0"], res["1"]
    assert torch.equal(oa, ob) or float((oa - ob).norm() / oa.norm()) <= 1e-6   # the forward pass computes the same values either way
    assert ga.keys() == gb.keys() and len(ga) > 0
    for k in ga:
        a, b = ga[k], gb[k]
        if a.is_complex():
            a, b = torch.view_as_real(a), torch.view_as_real(b)
        assert torch.isfinite(b).all(), k
        err = float((a - b).norm() / a.norm().clamp_min(1e-30))
        assert err <= 2e-5, (switch, k, err)
