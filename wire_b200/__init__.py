"""wire_b200 — B200-native (sm_100a) implementation of the WIRE hot path.

The forward and backward pass through the complex Gabor-wavelet MLP of Annatk26/wire
(`modules/wire.py`, `modules/wire2d.py`), behind the reference's own ``models.get_INR`` / ``nn.Module``
API.  All arithmetic runs in hand-written CUDA kernels (tcgen05/TMEM/TMA) reached through the C ABI in
``include/wire_b200.h``; there is no CPU or eager-PyTorch fallback.
"""
from . import _lib
from ._lib import WireB200Error
from .modules import models, wire, wire2d
from .modules.models import get_INR
from .patch import patch_reference
from .train import Trainer
from . import data
from . import lin_inverse
from .data import GridBatcher, run_epoch

__all__ = ["models", "wire", "wire2d", "get_INR", "patch_reference", "Trainer", "GridBatcher", "run_epoch", "data", "lin_inverse", "WireB200Error", "_lib"]
__version__ = "0.1.0"
