"""ctypes binding of the C ABI in ``include/wire_b200.h`` (the drop-in boundary).

The shared library is built in-tree by ``python -m wire_b200.build`` (or ``__graft_entry__.build()``)
into ``wire_b200/lib/libwire_b200.so``.  There is no fallback: if the library is missing, or the
device is not an sm_100 part, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

MAX_LAYERS = 16
PRECISION_TF32 = 0
PRECISION_FP32 = 1
PRECISION_MIXED16 = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwire_b200.so")


class NetDesc(Structure):
    _fields_ = [("two_d", c_int32), ("in_features", c_int32), ("width", c_int32),
                ("hidden_layers", c_int32), ("out_features", c_int32), ("precision", c_int32)]


class LayerParams(Structure):
    _fields_ = [("weight", c_void_p), ("bias", c_void_p), ("weight2", c_void_p), ("bias2", c_void_p),
                ("omega0", c_void_p), ("scale0", c_void_p)]


class NetParams(Structure):
    _fields_ = [("layer", LayerParams * MAX_LAYERS), ("final_weight", c_void_p), ("final_bias", c_void_p)]


class LayerGrads(Structure):
    _fields_ = [("weight", c_void_p), ("bias", c_void_p), ("weight2", c_void_p), ("bias2", c_void_p),
                ("omega0", c_void_p), ("scale0", c_void_p)]


GRADS_CLEAR_SLOTS, GRADS_CLEAR_FLAT, GRADS_PREZEROED = 0, 1, 2


class NetGrads(Structure):
    _fields_ = [("layer", LayerGrads * MAX_LAYERS), ("final_weight", c_void_p), ("final_bias", c_void_p),
                ("clear_mode", c_int32), ("flat_base", c_void_p), ("flat_floats", c_size_t)]


# name -> (restype, argtypes); mirrors include/wire_b200.h one to one
SIGNATURES = {
    "wire_b200_abi_version": (c_int32, []),
    "wire_b200_last_error": (c_char_p, []),
    "wire_b200_device_ok": (c_int32, []),
    "wire_b200_sm_count": (c_int32, []),
    "wire_b200_infer_chunk_rows": (c_int64, []),
    "wire_b200_prof_enable": (c_int32, [c_int32]),
    "wire_b200_prof_reset": (c_int32, []),
    "wire_b200_prof_kinds": (c_int32, []),
    "wire_b200_prof_name": (c_char_p, [c_int32]),
    "wire_b200_prof_get": (c_int32, [c_int32, POINTER(ctypes.c_uint64), POINTER(ctypes.c_double)]),
    "wire_net_workspace_bytes": (c_size_t, [POINTER(NetDesc), c_int64, c_int32]),
    "wire_net_workspace_init": (c_int32, [POINTER(NetDesc), c_int64, c_int32, c_void_p, c_size_t, c_void_p]),
    "wire_net_forward": (c_int32, [POINTER(NetDesc), POINTER(NetParams), c_void_p, c_int64, c_void_p, c_void_p,
                                   c_size_t, c_int32, c_void_p]),
    "wire_net_backward": (c_int32, [POINTER(NetDesc), POINTER(NetParams), c_void_p, c_int64, c_void_p, c_void_p,
                                    c_size_t, POINTER(NetGrads), c_void_p, c_void_p]),
    "wire_net_workspace_read": (c_int32, [POINTER(NetDesc), c_int64, c_void_p, c_size_t, c_int32, c_int32, c_void_p, c_void_p]),
    "wire_net_backward_mse": (c_int32, [POINTER(NetDesc), POINTER(NetParams), c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                        c_int32, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(NetGrads), c_void_p, c_void_p]),
    "wire_gabor_layer_workspace_bytes": (c_size_t, [POINTER(NetDesc), c_int32, c_int32, c_int64]),
    "wire_gabor_layer_forward": (c_int32, [POINTER(NetDesc), c_int32, c_int32, POINTER(LayerParams), c_void_p,
                                           c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "wire_gabor_layer_backward": (c_int32, [POINTER(NetDesc), c_int32, c_int32, POINTER(LayerParams), c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_int64, c_void_p, POINTER(LayerGrads),
                                            c_void_p, c_size_t, c_void_p]),
    "wire_final_linear_forward": (c_int32, [POINTER(NetDesc), c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                            c_void_p]),
    "wire_final_linear_backward": (c_int32, [POINTER(NetDesc), c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                             c_void_p, c_void_p, c_void_p]),
    "wire_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float,
                                 c_float, c_float, c_int64, c_float, c_void_p]),
    "wire_adam_step_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float,
                                     c_float, c_void_p, c_float, c_void_p, c_int32, c_void_p]),
    "wire_mse_loss_grad": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "wire_mse_loss_grad_n": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "wire_mse_loss_grad_ring": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "wire_peer_header_bytes": (c_size_t, []),
    "wire_peer_alloc": (c_int32, [c_size_t, POINTER(c_void_p), c_void_p]),
    "wire_peer_open": (c_int32, [c_void_p, POINTER(c_void_p)]),
    "wire_peer_close": (c_int32, [c_void_p]),
    "wire_peer_free": (c_int32, [c_void_p]),
    "wire_adam_step_peer": (c_int32, [c_void_p, POINTER(c_void_p), c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_float, c_float, c_float, c_float, c_void_p, c_float, c_void_p, c_void_p]),
    "wire_peer_wait_done": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_void_p, c_void_p]),
    "wire_grid_batch": (c_int32, [POINTER(c_int32), c_int32, c_int32, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "wire_scatter_rows": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "wire_iou_counts": (c_int32, [c_void_p, c_void_p, c_int64, c_float, c_int32, c_int32, c_void_p, c_void_p]),
    "wire_sq_err_stats": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "wire_radon_forward": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p]),
    "wire_radon_backward": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p]),
    "wire_gabor_scalar_grads": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "wire_real_gabor_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "wire_real_gabor_layer_forward": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "wire_real_gabor_layer_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_float,
                                                 c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "wire_real_gabor_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "wire_avgpool_mse_loss_grad": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class WireB200Error(RuntimeError):
    pass


def load(path: str = LIB_PATH) -> ctypes.CDLL:
    """Load the shared library and bind every symbol the header declares (no compute happens)."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise WireB200Error(
            f"{path} not found: build it with `python -m wire_b200.build` (nvcc, sm_100a). "
            "wire_b200 has no CPU or PyTorch fallback for the WIRE hot path.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().wire_b200_last_error()
        raise WireB200Error(f"{what} failed: {msg.decode() if msg else 'unknown error'}")
