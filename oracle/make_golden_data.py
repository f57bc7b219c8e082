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
