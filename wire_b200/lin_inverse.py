"""CUDA-backed counterpart of the part of the reference's ``modules/lin_inverse.py`` that sits next to the WIRE hot path:
the Radon forward operator of the CT driver (``radon``, modules/lin_inverse.py:19-40; called on the network's output every
iteration, wire_ct.py:126-128).  Same signature; autograd through the C ABI's ``wire_radon_forward`` / ``wire_radon_backward``.

The reference rotates with ``kornia.geometry.rotate`` (bilinear, zero padding, align_corners=True, about the image centre).
kornia is not installed in the build image, so the convention is restated from its published source
("Radon"; the CPU restatement used by the tests lives with the other checkers); parity against kornia itself is unpinned.
"""
from __future__ import annotations

import torch

from . import _lib
from . import functional as F
from ._lib import WireB200Error, check


class _RadonFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, imten: torch.Tensor, angles: torch.Tensor):
        lib = _lib.load()
        if imten.dim() != 4 or imten.shape[0] != 1:
            raise WireB200Error("radon expects a (1, nimg, H, W) image tensor")
        im = F._require_cuda(imten, "image", torch.float32)
        ang = F._require_cuda(angles.to(torch.float32), "angles", torch.float32).reshape(-1)
        _, nimg, H, W = im.shape
        sino = torch.empty((ang.numel(), nimg, W), dtype=torch.float32, device=im.device)
        with torch.cuda.device(im.device):
            check(lib.wire_radon_forward(im.data_ptr(), nimg, H, W, ang.data_ptr(), ang.numel(), sino.data_ptr(), F._stream()),
                  "wire_radon_forward")
        ctx.save_for_backward(ang)
        ctx.shape = (nimg, H, W)
        return sino

    @staticmethod
    def backward(ctx, g_sino):
        lib = _lib.load()
        (ang,) = ctx.saved_tensors
        nimg, H, W = ctx.shape
        g = F._require_cuda(g_sino, "grad sinogram", torch.float32)
        g_im = torch.empty((1, nimg, H, W), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.wire_radon_backward(g.data_ptr(), nimg, H, W, ang.data_ptr(), ang.numel(), g_im.data_ptr(), F._stream()),
                  "wire_radon_backward")
        return g_im, None


def radon(imten: torch.Tensor, angles: torch.Tensor, is_3d: bool = False) -> torch.Tensor:
    """``modules/lin_inverse.py:19-40``.  imten: (1, nimg, H, W) CUDA tensor; angles: (nangles,) degrees on the same device.
    Returns what the reference returns: ``imten_rot.sum(2).squeeze()`` — (nangles, W) for one image — or, with ``is_3d``,
    ``.permute(1, 0, 2)`` of it: (nimg, nangles, W)."""
    sino = _RadonFn.apply(imten, angles)          # (nangles, nimg, W) == imten_rot.sum(2)
    if is_3d:
        return sino.squeeze().permute(1, 0, 2)
    return sino.squeeze()
