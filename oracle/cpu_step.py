"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU path for one fwd+bwd sample, timed on the
host cores by bench.py (`cpu_baseline` leg and `--impl reference`).

It is the oracle port of the reference's hot path (graph build: oracle.graph = restated
torch_cluster KD-tree search; GNO: oracle.gno = reference integral_transform.py + scatter_native;
transformer block: reference attn.py:205-230 with F.scaled_dot_product_attention exactly as attn.py:126)
with autograd.  A full 10-layer S=16384 transformer fwd+bwd does not fit a few-minute budget on CPU, so a
*bounded sample* is timed: the complete graph build, the complete GNO encoder and decoder fwd+bwd, and
`layers_timed` of the `num_layers` transformer blocks (their cost is identical per block) scaled up.
"""
import math
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import gno as ogno
from . import graph as og
from .rope import RotaryEmbedding


def _rms(x, w, eps=1e-6):
    return (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)) * w


class _Block(torch.nn.Module):
    def __init__(self, hs, ffn, heads, skip):
        super().__init__()
        self.q, self.k, self.v, self.o = (torch.nn.Linear(hs, hs, bias=False) for _ in range(4))
        self.w1, self.w3 = torch.nn.Linear(hs, ffn, bias=False), torch.nn.Linear(hs, ffn, bias=False)
        self.w2 = torch.nn.Linear(ffn, hs, bias=False)
        self.n1, self.n2 = torch.nn.Parameter(torch.ones(hs)), torch.nn.Parameter(torch.ones(hs))
        self.skip = torch.nn.Linear(2 * hs, hs) if skip else None
        self.heads = heads
        self.rope = RotaryEmbedding(hs // heads)

    def forward(self, x, skip=None):
        if self.skip is not None and skip is not None:
            x = self.skip(torch.cat([x, skip], -1))
        B, S, H = x.shape
        h = _rms(x, self.n1)
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).transpose(1, 2)
        q, k, v = sp(self.q(h)), sp(self.k(h)), sp(self.v(h))
        q, k = self.rope.rotate_queries_or_keys(q), self.rope.rotate_queries_or_keys(k)
        a = F.scaled_dot_product_attention(q, k, v)
        h = x + self.o(a.transpose(1, 2).reshape(B, S, H))
        h = _rms(h, self.n2)
        return h + self.w2(F.silu(self.w1(h)) * self.w3(h))


def cpu_step_seconds(pos, normals, target, lat, cfg, layers_timed=1, search_workers=-1, seed=0):
    """Returns dict(seconds per full fwd+bwd sample (extrapolated), parts=...).  cfg: dict(k, C, hidden, heads,
    ffn, num_layers, patch, latent_tokens, enc_mlp, dec_mlp)."""
    torch.manual_seed(seed)
    C, k = cfg["C"], cfg["k"]
    parts = {}
    t0 = time.perf_counter()
    enc_e = torch.from_numpy(og.knn_np(lat, pos, k, workers=search_workers))                 # [phys, latent]
    dec_e = enc_e.flip(0).contiguous()                                                       # decoder knn = flipped search
    # the reference runs the search twice (encoder and decoder call get_neighbor_strategy independently)
    _ = og.knn_np(lat, pos, k, workers=search_workers)
    parts["graph"] = time.perf_counter() - t0

    P, L = torch.from_numpy(pos), torch.from_numpy(lat)
    feat = torch.cat([P, torch.from_numpy(normals)], -1)
    lift = torch.nn.Linear(6, C)
    mk = lambda dims: ([torch.nn.Parameter(torch.randn(dims[i + 1], dims[i]) / math.sqrt(dims[i])) for i in range(len(dims) - 1)],
                       [torch.nn.Parameter(torch.zeros(dims[i + 1])) for i in range(len(dims) - 1)])
    we, be = mk(cfg["enc_mlp"])
    wd, bd = mk(cfg["dec_mlp"])
    proj = torch.nn.Sequential(torch.nn.Linear(C, 256), torch.nn.GELU(), torch.nn.Linear(256, target.shape[1]))

    t0 = time.perf_counter()
    lat_feat = ogno.integral_transform(P, L, enc_e, lift(feat), we, be)
    g_lat = torch.randn_like(lat_feat)
    lat_feat.backward(g_lat)
    parts["encoder"] = time.perf_counter() - t0

    Pz = cfg["patch"]
    D, H, W = cfg["latent_tokens"]
    S = (D // Pz) * (H // Pz) * (W // Pz)
    hs = cfg["hidden"]
    x = torch.randn(1, S, hs, requires_grad=True)
    blk = _Block(hs, cfg["ffn"], cfg["heads"], skip=False)
    t0 = time.perf_counter()
    for _ in range(layers_timed):
        y = blk(x)
        y.backward(torch.ones_like(y))
    parts["transformer_per_layer"] = (time.perf_counter() - t0) / layers_timed
    parts["transformer"] = parts["transformer_per_layer"] * cfg["num_layers"]

    rn = torch.randn(L.shape[0], C, requires_grad=True)
    t0 = time.perf_counter()
    out = proj(ogno.integral_transform(L, P, dec_e, rn, wd, bd))
    loss = F.mse_loss(out, torch.from_numpy(target))
    loss.backward()
    parts["decoder"] = time.perf_counter() - t0
    total = parts["graph"] + parts["encoder"] + parts["transformer"] + parts["decoder"]
    return {"seconds": total, "parts": parts}
