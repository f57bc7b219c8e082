"""The CT driver's forward operator on the device (wire_b200.lin_inverse.radon, C ABI wire_radon_forward / _backward) against
the oracle restatement of modules/lin_inverse.py:19-40 (oracle/wire_oracle.py radon_torch; parity against kornia itself is
unpinned — kornia is not in the image), and the CT iteration of wire_ct.py:126-138 built on it."""
import numpy as np
import pytest
import torch

import util
import wire_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,nimg,nangles", [(256, 256, 1, 100), (96, 130, 1, 33), (64, 64, 3, 20), (17, 9, 1, 7)])
def test_radon_forward_and_adjoint_vs_oracle(H, W, nimg, nangles):
    from wire_b200 import lin_inverse
    rs = np.random.RandomState(H + W)
    im = torch.from_numpy(rs.uniform(size=(1, nimg, H, W)).astype(np.float32))
    angles = torch.from_numpy(np.linspace(0, 180, nangles, endpoint=False).astype(np.float32))
    gy_shape = (nangles, W) if nimg == 1 else (nimg, nangles, W)
    gy = torch.from_numpy(rs.normal(size=gy_shape).astype(np.float32))
    # oracle in float64 on the CPU
    im_r = im.double().requires_grad_(True)
    ref = O.radon_torch(im_r, angles.double(), is_3d=nimg > 1)
    (ref * gy.double()).sum().backward()
    im_c = im.cuda().requires_grad_(True)
    got = lin_inverse.radon(im_c, angles.cuda(), is_3d=nimg > 1)
    assert tuple(got.shape) == tuple(ref.shape)
    (got * gy.cuda()).sum().backward()
    torch.cuda.synchronize()
    e_f = util.rel_err(got.detach().cpu().numpy(), ref.detach().numpy())
    e_b = util.rel_err(im_c.grad.cpu().numpy(), im_r.grad.numpy())
    util.record("radon", f"{H}x{W}x{nimg}x{nangles}", {"forward": e_f, "adjoint": e_b})
    # float32 bilinear weights near pixel boundaries: a coordinate error of ~1e-5 pixel moves weight between two taps
    assert e_f < 2e-5 and e_b < 2e-5, (e_f, e_b)


def test_ct_iteration_on_device_matches_the_oracle_operator():
    """wire_ct.py:126-138 on a 64x64 phantom with 40 angles, 12 iterations: model(coords) -> radon -> MSE against the measured
    sinogram -> backward -> Adam, once with the CUDA operator and once with the oracle's operator (torch grid_sample on the
    GPU) behind the same FP32 CUDA model: the loss trajectories must coincide."""
    import wire_b200
    from wire_b200 import lin_inverse
    H = W = 64
    nmeas = 40
    yy, xx = np.meshgrid(np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
    img = ((xx ** 2 + (yy * 1.3) ** 2 < 0.6).astype(np.float32) * 0.6 + ((xx - 0.2) ** 2 + (yy + 0.1) ** 2 < 0.05) * 0.4).astype(np.float32)
    imten = torch.from_numpy(img)[None, None].cuda()
    thetas = torch.tensor(np.linspace(0, 180, nmeas, dtype=np.float32)).cuda()
    coords = O.image_coords(H, W).cuda()
    kw = dict(nonlin="wire", in_features=2, out_features=1, hidden_features=128, hidden_layers=2, first_omega_0=3.0,
              hidden_omega_0=3.0, scale=4.0, precision="fp32")
    torch.manual_seed(0)
    init = wire_b200.get_INR(**kw).state_dict()
    with torch.no_grad():
        sino_a = lin_inverse.radon(imten, thetas)
        sino_b = O.radon_torch(imten, thetas)
    assert util.rel_err(sino_a.cpu().numpy(), sino_b.cpu().numpy()) < 2e-5
    losses = {}
    for name, op in (("cuda", lin_inverse.radon), ("oracle", O.radon_torch)):
        m = wire_b200.get_INR(**kw)
        m.load_state_dict(init)
        m.cuda()
        opt = torch.optim.Adam(lr=5e-3, params=m.parameters())
        cur = []
        for _ in range(12):
            img_estim = m(coords).reshape(-1, H, W)[None, ...]
            sino_estim = op(img_estim, thetas)
            loss = ((sino_b - sino_estim) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            cur.append(float(loss))
        losses[name] = cur
    assert util.rel_err(np.array(losses["cuda"]), np.array(losses["oracle"])) < 1e-3, losses
    assert losses["cuda"][-1] < losses["cuda"][0]
