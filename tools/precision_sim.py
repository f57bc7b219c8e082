"""Emulate operand roundings of the tensor-core paths on CPU (float64 math + explicit rounding) and report
relative errors vs the float64 closed form.  Decides which 16-bit formats are safe for which operand.
usage: python tools/precision_sim.py [wire|wire2d] [H]"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import wire_oracle as O

def rnd(a, fmt):
    if fmt is None: return a
    if np.iscomplexobj(a): return rnd(a.real, fmt) + 1j * rnd(a.imag, fmt)
    t = torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32)
    if fmt == "f16": r = t.to(torch.float16).to(torch.float64)
    elif fmt == "bf16": r = t.to(torch.bfloat16).to(torch.float64)
    elif fmt == "tf32":
        i = t.view(torch.int32); i = (i + 0x1000) & ~0x1FFF; r = i.view(torch.float32).to(torch.float64)
    elif fmt == "f32": r = t.to(torch.float64)
    else: raise ValueError(fmt)
    return r.numpy()

def run(state, coords, g_out, fy, fw, fg, fwb, fz, gscale=1.0, fx=None):
    """fy: format of stored y (GEMM A operand); fw: fwd weight format; fg: stored g_z; fwb: dgrad weight fmt; fz: saved z"""
    layers, final = O._layers_from_state(state)
    x = coords.astype(np.float64); saved = []
    for li, L in enumerate(layers):
        first = li == 0
        W = L["W"].astype(np.complex128) if not first else L["W"].astype(np.float64)
        z = (x @ (W if first else rnd(W, fw)).T) + L["b"]
        z = rnd(z, "f32")
        y = np.exp(1j * L["omega"] * z - L["scale"] ** 2 * np.abs(z) ** 2)
        zs = rnd(z, fz) if not first else z
        ys = rnd(y, fy)
        saved.append(dict(x=rnd(x, fx) if not first else x, z=zs, y=y)); x = ys
    # final layer on unrounded y (fused in epilogue)
    h = saved[-1]["y"]
    out = (h @ final["W"].astype(np.complex128).T + final["b"]).real
    grads = {}
    n = len(layers)
    grads[f"net.{n}.weight"] = g_out.T.astype(np.complex128) @ np.conj(h)
    g_y = g_out.astype(np.complex128) @ np.conj(final["W"].astype(np.complex128))
    for i in range(n - 1, -1, -1):
        L, S = layers[i], saved[i]; first = i == 0
        zz = S["z"]
        yy = np.exp(1j * L["omega"] * zz - L["scale"] ** 2 * np.abs(zz) ** 2)  # recomputed from saved z
        p = np.conj(yy) * g_y; s2 = L["scale"] ** 2
        g_z = (L["omega"] * p.imag - 2 * s2 * zz * p.real) if first else (-1j * L["omega"] * p - 2 * s2 * zz * p.real)
        if not first: g_z = rnd(g_z * gscale, fg) / gscale
        grads[f"net.{i}.linear.weight"] = g_z.T @ np.conj(S["x"])
        grads[f"net.{i}.linear.bias"] = g_z.sum(0)
        if not first: g_y = g_z @ np.conj(rnd(L["W"].astype(np.complex128), fwb))
    return out, grads

def rel(a, b): return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

if __name__ == "__main__":
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = dict(D=("wire", 2, 300, H, 3, 7.0, 7.0, 6.0), O=("wire", 3, 300, 3, 1, 20.0, 20.0, 10.0))
    for name, c in cfg.items():
        m = O.TorchOracle(*c); st = O.deterministic_state(m, 5)
        state = {k: v.numpy() for k, v in st.items()}
        rs = np.random.RandomState(1); N = 4096
        coords = rs.uniform(-1, 1, (N, c[1])); g_out = rs.randn(N, c[4]) / (N * c[4])
        ref_out, ref_g = run(state, coords, g_out, None, None, None, None, None)
        for label, f in dict(tf32=("tf32", "tf32", "tf32", "tf32", "f16"),
                             mixed16=("f16", "f16", "bf16", "bf16", "f16"),
                             mixed16_wtf=("f16", "f16", "bf16", "f16", "f16"),
                             allbf16=("bf16", "bf16", "bf16", "bf16", "f16"),
                             mixed16_xbf=("f16", "f16", "bf16", "bf16", "f16", "bf16"),
                             f16_scaled=("f16", "f16", "f16", "f16", "f16")).items():
            gs = 1.0
            if label == "f16_scaled": gs = 2.0 ** 20
            fx = f[5] if len(f) > 5 else None
            o, g = run(state, coords, g_out, *f[:5], gscale=gs, fx=fx)
            print(f"{name} {label:12s} out {rel(o, ref_out):.2e} " + " ".join(f"{k.split('net.')[1][:8]}:{rel(g[k], ref_g[k]):.1e}" for k in sorted(ref_g)))
