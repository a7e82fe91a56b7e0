"""Dev helper (GPU box): time the attention forward/backward kernels under a list of GAOT_ATTN_DEBUG values.
usage: python profiles/tools/prof_attn_dbg.py 0 1 2 ...   (debug bits skip parts of the kernels: results are then WRONG, timing only)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from gaot_3d_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, H, d = 1, 16384, 8, 32
q, k, v = (torch.randn(B, S, H * d, device=dev, requires_grad=True) for _ in range(3))
fr = (1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))).to(dev)
go = torch.randn(B, S, H * d, device=dev)
lib = _lib.load()
for dbg in sys.argv[1:]:
    os.environ["GAOT_ATTN_DEBUG"] = dbg
    for rep in range(2):
        lib.gaot_profile_enable(1)
        for _ in range(5):
            o = ops.attention(q, k, v, H, H, rope_freqs=fr)
            o.backward(go)
        torch.cuda.synchronize()
        buf = ctypes.create_string_buffer(4096)
        lib.gaot_profile_summary(buf, 4096)
        lib.gaot_profile_enable(0)
    out = {l.split()[0]: float(l.split()[2]) / int(l.split()[1]) for l in buf.value.decode().splitlines()}
    print(f"debug={dbg:>4}: " + "  ".join(f"{n} {t:.3f} ms" for n, t in out.items()), flush=True)
