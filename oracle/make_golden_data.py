"""Golden vectors for the data pipeline, produced by RUNNING THE REFERENCE'S OWN FUNCTIONS (build container only).

``modules/utils.py`` and ``modules/volutils.py`` import plotting / meshing packages that are not installed here
(matplotlib, mcubes, open3d, skimage, cv2 may be missing): those imports are stubbed with empty modules — none of the
functions exercised below touches them — and ``get_coords`` / ``psnr`` / ``get_I_and_U`` are then called unmodified.
The image drivers build their coordinates inline (wire_image_denoise.py:63-66); that snippet is executed verbatim.

    python oracle/make_golden_data.py       # rewrites tests/golden/data_pipeline.npz
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("WIRE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "data_pipeline.npz")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def stub_missing(names):
    for n in names:
        try:
            __import__(n)
        except Exception:
            parts = n.split(".")
            for i in range(1, len(parts) + 1):
                sys.modules.setdefault(".".join(parts[:i]), _Stub(".".join(parts[:i])))


def main():
    stub_missing(["matplotlib", "matplotlib.pyplot", "mcubes", "open3d", "skimage", "skimage.metrics", "cv2", "pandas",
                  "scipy.io", "scipy.linalg", "scipy.interpolate"])
    sys.path.insert(0, REF)
    from modules import utils as ref_utils        # the reference, unmodified
    from modules import volutils as ref_volutils  # the reference, unmodified

    blob = {}
    for H, W, T in [(5, 7, 3), (4, 6, None), (1, 9, 2), (33, 17, None), (8, 8, 8)]:
        key = f"coords_np_{H}_{W}_{T or 0}"
        blob[key] = ref_utils.get_coords(H, W, T).numpy()
    for H, W in [(6, 9), (16, 16), (7, 1000), (513, 4)]:
        x = torch.linspace(-1, 1, W)                                    # wire_image_denoise.py:63-66, verbatim
        y = torch.linspace(-1, 1, H)
        X, Y = torch.meshgrid(x, y, indexing="xy")
        coords = torch.hstack((X.reshape(-1, 1), Y.reshape(-1, 1)))[None, ...]
        blob[f"coords_torch_{H}_{W}"] = coords[0].numpy()

    rs = np.random.RandomState(5)
    preds = rs.uniform(-0.3, 1.3, size=4099).astype(np.float32)
    preds[::97] = 0.5                                                    # values exactly at the threshold
    gt = (rs.uniform(size=4099) < 0.2).astype(np.float32)
    blob["iou_preds"], blob["iou_gt"] = preds.copy(), gt
    res = []
    for thres in (0.5, 0.0, -1.0, 2.0):
        p = preds.copy()
        i, u = ref_volutils.get_I_and_U(p, gt, thres)
        res.append((thres, int(i), int(u)))
        blob[f"iou_binarized_{thres}"] = p
    i, u = ref_volutils.get_I_and_U(preds.copy(), gt, None)
    res.append((np.nan, int(i), int(u)))
    blob["iou_results"] = np.array(res, dtype=np.float64)
    blob["iou_value_0.5"] = np.array(ref_volutils.get_IoU(preds.copy(), gt, 0.5), dtype=np.float64)

    x = rs.uniform(0, 1, size=(32, 32, 3)).astype(np.float32)
    xhat = (x + rs.normal(scale=0.05, size=x.shape)).astype(np.float32)
    blob["psnr_x"], blob["psnr_xhat"] = x, xhat
    blob["psnr_value"] = np.array(ref_utils.psnr(x, xhat), dtype=np.float64)

    np.savez_compressed(OUT, **blob)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1024:.1f} KiB; iou {res}; psnr {float(blob['psnr_value']):.6f}")


if __name__ == "__main__":
    main()


# ---- trainable omega_0 / scale_0 (modules/wire.py:66,80-81; modules/wire2d.py:27,42-43): the reference's own layers ----
def trainable_fixture():
    """ComplexGaborLayer / ComplexGaborLayer2D with trainable=True, first and hidden, forward + backward in complex64 and
    complex128: gradients of omega_0, scale_0 (and of the Linear's parameters and the input, for completeness)."""
    sys.path.insert(0, REF)
    from modules import wire as ref_wire, wire2d as ref_wire2d
    blob = {}
    rs = np.random.RandomState(11)
    for tag, cls, is_first, K, M, n, w0, s0 in [("wire_first", ref_wire.ComplexGaborLayer, True, 2, 24, 40, 7.0, 6.0),
                                                 ("wire_hidden", ref_wire.ComplexGaborLayer, False, 24, 24, 40, 7.0, 6.0),
                                                 ("wire2d_first", ref_wire2d.ComplexGaborLayer2D, True, 3, 16, 36, 8.0, 9.0),
                                                 ("wire2d_hidden", ref_wire2d.ComplexGaborLayer2D, False, 16, 16, 36, 8.0, 9.0)]:
        if is_first:
            x = rs.uniform(-1, 1, size=(1, n, K))
        else:
            x = (rs.normal(size=(1, n, K)) + 1j * rs.normal(size=(1, n, K))) * 0.3
        gy = rs.normal(size=(1, n, M)) + 1j * rs.normal(size=(1, n, M))
        torch.manual_seed(5)
        layer = cls(K, M, is_first=is_first, omega0=w0, sigma0=s0, trainable=True)
        state = {k: v.clone() for k, v in layer.state_dict().items()}
        for k, v in state.items():
            a = v.numpy()
            blob[f"{tag}.param.{k}"] = a
        for prec, double in (("c64", False), ("c128", True)):
            layer = cls(K, M, is_first=is_first, omega0=w0, sigma0=s0, trainable=True)
            layer.load_state_dict(state)
            if double:
                for p in layer.parameters():
                    p.data = p.data.to(torch.complex128 if p.is_complex() else torch.float64)
            xt = torch.from_numpy(x.astype((np.float64 if double else np.float32) if is_first else (np.complex128 if double else np.complex64)))
            xt.requires_grad_(True)
            gt = torch.from_numpy(gy.astype(np.complex128 if double else np.complex64))
            y = layer(xt)
            torch.view_as_real(y).mul(torch.view_as_real(gt)).sum().backward()     # sum Re(conj(g) y): upstream gradient g
            blob[f"{tag}.y_{prec}"] = y.detach().numpy()
            blob[f"{tag}.g_omega_{prec}"] = layer.omega_0.grad.numpy()
            blob[f"{tag}.g_scale_{prec}"] = layer.scale_0.grad.numpy()
            blob[f"{tag}.g_x_{prec}"] = xt.grad.numpy()
            blob[f"{tag}.g_weight_{prec}"] = layer.linear.weight.grad.numpy()
        blob[f"{tag}.x"] = x
        blob[f"{tag}.gy"] = gy
        blob[f"{tag}.meta"] = np.array([int(is_first), K, M, n], dtype=np.int64)
        blob[f"{tag}.hyper"] = np.array([w0, s0], dtype=np.float64)
    out = os.path.join(HERE, "..", "tests", "golden", "trainable_scalars.npz")
    np.savez_compressed(out, **blob)
    print(f"wrote {out}: {os.path.getsize(out) / 1024:.1f} KiB; g_omega(wire_hidden) = {blob['wire_hidden.g_omega_c128']}, "
          f"g_scale = {blob['wire_hidden.g_scale_c128']}")


if __name__ == "__main__":
    trainable_fixture()


def real_gabor_fixture():
    """RealGaborLayer (modules/wire.py:6-42), the reference's own class: forward + autograd in float32 and float64."""
    sys.path.insert(0, REF)
    from modules import wire as ref_wire
    blob = {}
    rs = np.random.RandomState(21)
    K, M, n, w0, s0 = 3, 20, 50, 10.0, 10.0
    x = rs.uniform(-1, 1, size=(1, n, K))
    gy = rs.normal(size=(1, n, M))
    torch.manual_seed(3)
    layer = ref_wire.RealGaborLayer(K, M, omega0=w0, sigma0=s0)
    state = {k: v.clone() for k, v in layer.state_dict().items()}
    for k, v in state.items():
        blob[f"param.{k}"] = v.numpy()
    for prec, dt in (("f32", torch.float32), ("f64", torch.float64)):
        layer = ref_wire.RealGaborLayer(K, M, omega0=w0, sigma0=s0)
        layer.load_state_dict(state)
        layer = layer.to(dt)
        xt = torch.from_numpy(x).to(dt).requires_grad_(True)
        y = layer(xt)
        (y * torch.from_numpy(gy).to(dt)).sum().backward()
        blob[f"y_{prec}"] = y.detach().numpy()
        blob[f"g_x_{prec}"] = xt.grad.numpy()
        blob[f"g_freqs_w_{prec}"] = layer.freqs.weight.grad.numpy()
        blob[f"g_scale_w_{prec}"] = layer.scale.weight.grad.numpy()
        blob[f"g_scale_b_{prec}"] = layer.scale.bias.grad.numpy()
    blob["x"], blob["gy"] = x, gy
    blob["meta"] = np.array([K, M, n], dtype=np.int64)
    blob["hyper"] = np.array([w0, s0])
    out = os.path.join(HERE, "..", "tests", "golden", "real_gabor.npz")
    np.savez_compressed(out, **blob)
    print(f"wrote {out}: {os.path.getsize(out) / 1024:.1f} KiB")


if __name__ == "__main__":
    real_gabor_fixture()
