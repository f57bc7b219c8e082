"""ctypes wrapper of oracle/wire_oracle.c (TEST INFRASTRUCTURE). Build with `make -C oracle`."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libwire_oracle_c.so")


def available() -> bool:
    return os.path.exists(_PATH)


def _lib():
    lib = ctypes.CDLL(_PATH)
    lib.wire_oracle_param_doubles.restype = ctypes.c_size_t
    lib.wire_oracle_param_doubles.argtypes = [ctypes.c_int] * 5
    lib.wire_oracle_run.restype = ctypes.c_int
    lib.wire_oracle_run.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 8
    return lib


def _keys(two_d, H):
    ks = []
    for l in range(H + 1):
        ks += [f"net.{l}.linear.weight", f"net.{l}.linear.bias"]
        if two_d:
            ks += [f"net.{l}.scale_orth.weight", f"net.{l}.scale_orth.bias"]
    return ks + [f"net.{H + 1}.weight", f"net.{H + 1}.bias"]


def _flat(a):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        a = np.stack([a.real, a.imag], -1)
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def run(state, two_d, coords, grad_out=None):
    """state: name -> numpy array (reference state_dict names). Returns (out, grads dict or None, grad_coords)."""
    lib = _lib()
    H = max(int(k.split(".")[1]) for k in state) - 1
    M, in_f = state["net.0.linear.weight"].shape
    out_f = state[f"net.{H + 1}.weight"].shape[0]
    keys = _keys(two_d, H)
    params = np.concatenate([_flat(state[k]) for k in keys])
    assert params.size == lib.wire_oracle_param_doubles(int(two_d), in_f, M, H, out_f)
    omega = np.array([float(np.asarray(state[f"net.{l}.omega_0"]).reshape(-1)[0]) for l in range(H + 1)])
    scale = np.array([float(np.asarray(state[f"net.{l}.scale_0"]).reshape(-1)[0]) for l in range(H + 1)])
    c = np.ascontiguousarray(np.asarray(coords, dtype=np.float64).reshape(-1, in_f))
    n = c.shape[0]
    out = np.zeros((n, out_f))
    gp = gc = go = None
    if grad_out is not None:
        go = np.ascontiguousarray(np.asarray(grad_out, dtype=np.float64).reshape(n, out_f))
        gp = np.zeros_like(params)
        gc = np.zeros_like(c)
    ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.wire_oracle_run(int(two_d), n, in_f, M, H, out_f, ptr(c), ptr(params), ptr(omega), ptr(scale), ptr(go),
                             ptr(out), ptr(gp), ptr(gc))
    assert rc == 0
    grads = None
    if gp is not None:
        grads, off = {}, 0
        for k in keys:
            a = np.asarray(state[k])
            cnt = a.size * (2 if np.iscomplexobj(a) else 1)
            g = gp[off:off + cnt]
            off += cnt
            grads[k] = (g.reshape(*a.shape, 2)[..., 0] + 1j * g.reshape(*a.shape, 2)[..., 1]) if np.iscomplexobj(a) else g.reshape(a.shape)
    return out, grads, gc
