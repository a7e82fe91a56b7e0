"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU path for one fwd+bwd sample, timed on the
host cores by bench.py (`cpu_baseline` leg and `--impl reference`).

It is the oracle port of the reference's hot path (graph build: oracle.graph = restated
torch_cluster KD-tree search; GNO: oracle.gno = reference integral_transform.py + scatter_native;
transformer block: reference attn.py:205-230 with F.scaled_dot_product_attention exactly as attn.py:126)
with autograd.  One call = one complete fwd+bwd sample: the complete graph build, the complete GNO encoder and decoder
fwd+bwd and ALL `num_layers` transformer blocks chained (U-shaped: the second half takes the long-range skip through
Linear(2H->H), attn.py:210-212).  `layers_timed < num_layers` times only that many blocks and scales the rest -- the
result then carries `extrapolated: True`.  The search is timed with all host threads (`search_workers=-1`, scipy's
cKDTree) AND single-threaded, as torch_cluster's CPU search is (the reference never passes num_workers).
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


def cpu_step_seconds(pos, normals, target, lat, cfg, layers_timed=None, search_workers=-1, seed=0, single_thread_search=True):
    """Returns dict(seconds per full fwd+bwd sample, parts=..., extrapolated=bool).  cfg: dict(k, C, hidden, heads,
    ffn, num_layers, patch, latent_tokens, enc_mlp, dec_mlp[, strategy, radius])."""
    torch.manual_seed(seed)
    C, k = cfg["C"], cfg["k"]
    parts = {}
    strat = cfg.get("strategy", "knn")
    s_enc, s_dec = (strat, strat) if isinstance(strat, str) else strat
    radius = cfg.get("radius", 0.033)
    t0 = time.perf_counter()
    # the reference runs the search once per side (encoder and decoder call get_neighbor_strategy independently, magno.py:520,:770)
    enc_e = torch.from_numpy(og.get_neighbor_strategy_np(s_enc, pos, None, lat, None, radius, k, False, workers=search_workers))    # [phys, latent]
    dec_e = torch.from_numpy(og.get_neighbor_strategy_np(s_dec, pos, None, lat, None, radius, k, True, workers=search_workers))     # [latent, phys]
    parts["graph"] = time.perf_counter() - t0
    parts["edges_enc_dec"] = [int(enc_e.shape[1]), int(dec_e.shape[1])]
    if single_thread_search:                     # what the reference's torch_cluster CPU search is: one thread (reported, not summed)
        t0 = time.perf_counter()
        og.get_neighbor_strategy_np(s_enc, pos, None, lat, None, radius, k, False, workers=1)
        og.get_neighbor_strategy_np(s_dec, pos, None, lat, None, radius, k, True, workers=1)
        parts["graph_single_thread"] = time.perf_counter() - t0

    P, L = torch.from_numpy(pos), torch.from_numpy(lat)
    feat = torch.cat([P, torch.from_numpy(normals)], -1)
    lift = torch.nn.Linear(feat.shape[1], C)
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
    nl = cfg["num_layers"]
    n_run = nl if layers_timed is None else max(1, min(int(layers_timed), nl))
    n_enc = nl // 2
    blocks = [_Block(hs, cfg["ffn"], cfg["heads"], skip=(i >= nl - n_enc)) for i in range(nl)][:n_run]   # attn.py:267-295: nl//2 encoder, optional middle, nl//2 decoder (skips)
    t0 = time.perf_counter()
    skips, y = [], x
    for i, blk in enumerate(blocks):
        y = blk(y, skips.pop() if (blk.skip is not None and skips) else None)
        if i < n_enc:
            skips.append(y)
    y.backward(torch.ones_like(y))
    parts["transformer_per_layer"] = (time.perf_counter() - t0) / n_run
    parts["transformer"] = parts["transformer_per_layer"] * nl
    parts["transformer_layers_run"] = n_run

    rn = torch.randn(L.shape[0], C, requires_grad=True)
    t0 = time.perf_counter()
    out = proj(ogno.integral_transform(L, P, dec_e, rn, wd, bd))
    loss = F.mse_loss(out, torch.from_numpy(target))
    loss.backward()
    parts["decoder"] = time.perf_counter() - t0
    ps = float(cfg.get("point_scale", 1.0))          # > 1: `pos` is a 1/ps subsample of the cloud, point-proportional parts scaled up
    parts["point_scale"] = ps
    total = ps * (parts["graph"] + parts["encoder"] + parts["decoder"]) + parts["transformer"]
    return {"seconds": total, "parts": parts, "extrapolated": n_run < nl or ps != 1.0}
