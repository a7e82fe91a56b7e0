"""Standalone op runner for ncu / quick timing on the GPU box (not a pytest).
usage: python profiles/tools/prof_ops.py attn|gno|graph [reps]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gaot_3d_b200 import ops, _lib  # noqa: E402
from tests import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "attn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)


def timeit(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if what == "attn":
    B, S, H, d = 1, 16384, 8, 32
    q, k, v = (torch.randn(B, S, H * d, device=dev, requires_grad=True) for _ in range(3))
    fr = (1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))).to(dev)
    go = torch.randn(B, S, H * d, device=dev)
    lib = _lib.load()
    lib.gaot_profile_enable(1)

    def f():
        o = ops.attention(q, k, v, H, H, rope_freqs=fr)
        o.backward(go)
    ms = timeit(f, reps)
    import ctypes
    buf = ctypes.create_string_buffer(4096)
    lib.gaot_profile_summary(buf, 4096)
    print("fwd+bwd region ms", ms)
    for line in buf.value.decode().splitlines():
        n, c, t, w = line.split()
        print(n, "avg ms", float(t) / int(c), "TFLOP/s", float(w) / float(t) / 1e9)
elif what == "gno":
    N, G = 500000, (64, 64, 32)
    phys, lat = torch.from_numpy(synth.surface_cloud(N)).to(dev), torch.from_numpy(synth.latent_grid(G)).to(dev)
    for name, strat, dec in (("enc knn", "knn", False), ("dec radius", "radius", True)):
        from gaot_3d_b200.graph import get_neighbor_strategy
        ei = get_neighbor_strategy(strat, phys, None, lat, None, 0.033, 1, dec)
        ypos, xpos = (lat, phys) if dec else (phys, lat)
        layers = [6, 64, 64, 32] if dec else [6, 64, 64, 64, 32]
        ws = [(torch.randn(layers[i + 1], layers[i], device=dev) / np.sqrt(layers[i])).requires_grad_(True) for i in range(len(layers) - 1)]
        bs = [torch.zeros(layers[i + 1], device=dev, requires_grad=True) for i in range(len(layers) - 1)]
        f_y = torch.randn(ypos.shape[0], 32, device=dev, requires_grad=True)
        csr = ops.csr_of(ei, ypos.shape[0], xpos.shape[0])
        for prec in ("fp32", "bf16"):
            try:
                fw = timeit(lambda: ops.gno(ypos, xpos, f_y, csr, ws, bs, precision=prec), reps)
                out = ops.gno(ypos, xpos, f_y, csr, ws, bs, precision=prec)
                g = torch.randn_like(out)
                fb = timeit(lambda: ops.gno(ypos, xpos, f_y, csr, ws, bs, precision=prec).backward(g), reps)
                E = ei.shape[1]
                print(f"{name} {prec}: E={E} fwd {fw:.3f} ms ({E / fw / 1e6:.1f} G edges/s... {E / fw * 1e3 / 1e9:.2f} Gedge/s) fwd+bwd {fb:.3f} ms ({E / fb * 1e3 / 1e9:.2f} Gedge/s)")
            except NotImplementedError as e:
                print(name, prec, "unsupported:", e)
elif what == "dense":
    import ctypes
    lib = _lib.load()
    M = 16384
    for N, K in ((256, 256), (768, 256), (1024, 256), (256, 1024), (256, 512)):
        x = torch.randn(M, K, device=dev, requires_grad=True)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).requires_grad_(True)
        go = torch.randn(M, N, device=dev)
        lib.gaot_profile_enable(1)
        tf = timeit(lambda: ops.linear(x, w), reps)
        tfb = timeit(lambda: ops.linear(x, w).backward(go), reps)
        buf = ctypes.create_string_buffer(4096)
        lib.gaot_profile_summary(buf, 4096)
        lib.gaot_profile_enable(0)
        tt = timeit(lambda: torch.nn.functional.linear(x, w), reps)
        ttb = timeit(lambda: torch.nn.functional.linear(x, w).backward(go), reps)
        parts = {l.split()[0]: float(l.split()[2]) / int(l.split()[1]) for l in buf.value.decode().splitlines()}
        io = lambda *n: sum(n) * 4 / 1e6
        print(f"M={M} N={N} K={K}: ours fwd {tf*1e3:.1f} us fwd+bwd {tfb*1e3:.1f} us | torch fp32 fwd {tt*1e3:.1f} us fwd+bwd {ttb*1e3:.1f} us | "
              f"kernels us: " + " ".join(f"{k}={v*1e3:.1f}" for k, v in parts.items()) +
              f" | fwd HBM floor {io(M*K, N*K, M*N)/6.5e3*1e3:.1f} us")
elif what == "graph":
    from gaot_3d_b200.graph import get_neighbor_strategy
    for N in (100000, 500000, 1000000, 3160000, 10000000):
        G = (64, 64, 32)
        phys, lat = torch.from_numpy(synth.surface_cloud(N)).to(dev), torch.from_numpy(synth.latent_grid(G)).to(dev)
        for strat, dec in (("knn", False), ("radius", False), ("radius", True), ("bidirectional", False), ("bidirectional", True)):
            t = timeit(lambda: get_neighbor_strategy(strat, phys, None, lat, None, 0.033, 1, dec), reps)
            E = get_neighbor_strategy(strat, phys, None, lat, None, 0.033, 1, dec).shape[1]
            print(f"N={N} {strat} dec={dec}: {t:.3f} ms  E={E}  {N / t * 1e3 / 1e6:.1f} Mpoints/s  {E / t * 1e3 / 1e6:.1f} Medges/s", flush=True)
