"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the whole GAOT3D forward from a state_dict.

Follows reference src/model/gaot_3d.py:248-332 (forward), :166-222 (process: patchify, patch_linear,
positional embedding, Transformer, un-patchify), src/model/layers/magno.py:468-600 / :691-798
(encoder / decoder orchestration, mlp_type='linear') and src/model/layers/attn.py:205-230, :298-325
(block and U-Net skip wiring).  Pinned against the reference's own GAOT3D in
tests/test_oracle_vs_reference.py and against tests/golden/model_*.pt.
"""
import math
import torch
import torch.nn.functional as F

from . import gno as ognno
from . import graph as ograph
from .attn import attention_core


def _lin(sd, name, x, bias=True):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"] if bias and (name + ".bias") in sd else None)


def _mlp(sd, prefix, x):
    n = 0
    while f"{prefix}.fcs.{n}.weight" in sd:
        n += 1
    for i in range(n):
        x = _lin(sd, f"{prefix}.fcs.{i}", x)
        if i < n - 1:
            x = F.gelu(x)
    return x


def _mlp_wb(sd, prefix):
    w, b, n = [], [], 0
    while f"{prefix}.fcs.{n}.weight" in sd:
        w.append(sd[f"{prefix}.fcs.{n}.weight"]); b.append(sd[f"{prefix}.fcs.{n}.bias"]); n += 1
    return w, b


def _rms(x, w, eps):
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * w


def _block(sd, p, x, cfg, rope, skip=None):
    if skip is not None:                                                   # attn.py:222-224
        x = _lin(sd, p + ".skip_proj", torch.cat([x, skip], dim=-1))
    h = _rms(x, sd[p + ".attn_norm.weight"], cfg["norm_eps"])              # :226
    q, k, v = (_lin(sd, f"{p}.attn.{n}_proj", h, bias=False) for n in "qkv")
    a = attention_core(q, k, v, cfg["num_heads"], cfg["num_kv_heads"], rope)
    h = x + _lin(sd, p + ".attn.o_proj", a, bias=False)                    # :227
    h = _rms(h, sd[p + ".ffn_norm.weight"], cfg["norm_eps"])               # :228 (residual after the norm)
    f = _lin(sd, p + ".ffn.w2", F.silu(_lin(sd, p + ".ffn.w1", h, bias=False)) * _lin(sd, p + ".ffn.w3", h, bias=False), bias=False)
    return h + f                                                           # :229


def _abs_pe(positions, embed_dim):
    half = embed_dim // 2
    freq = 1 / 10000 ** (2 * torch.arange(0, half, dtype=torch.float32) / embed_dim)
    ang = positions[:, :, None] * freq[None, None, :]
    pe = torch.zeros(positions.shape[0], embed_dim)
    pe[:, 0::2] = torch.sin(ang).sum(1)
    pe[:, 1::2] = torch.cos(ang).sum(1)
    return pe


def gaot3d_forward(sd, cfg, pos, feats, latent_pos=None, enc_edges=None, dec_edges=None, keep_graph=False):
    """cfg: dict(latent_tokens, patch_size, lifting_channels, radius, k, enc_strategy, dec_strategy,
    use_geoembed=(enc,dec), num_layers, num_heads, num_kv_heads, norm_eps, positional_embedding, scales).
    Single example (batch of one), mlp_type='linear', transform 'linear'."""
    sd = {k: (v if keep_graph else v.detach()).float().cpu() for k, v in sd.items()}
    pos = pos.float().cpu()
    lat = (sd["latent_tokens"] if latent_pos is None else latent_pos).float().cpu()
    C = cfg["lifting_channels"]
    # ---- encoder (magno.py:468-600)
    phys_feat = torch.cat([f.float().cpu() for f in feats], dim=-1)
    enc_scales = []
    for si, scale in enumerate(cfg.get("scales", [1.0])):
        if enc_edges is None:
            ei = torch.from_numpy(ograph.get_neighbor_strategy_np(cfg["enc_strategy"], pos.numpy(), None, lat.numpy(), None,
                                                                  cfg["radius"] * scale, cfg["k"], False))
        else:
            ei = enc_edges[si]
        lifted = _mlp(sd, "encoder.lifting", phys_feat)
        w, b = _mlp_wb(sd, "encoder.gno.channel_mlp")
        enc = ognno.integral_transform(pos, lat, ei, lifted, w, b)
        if cfg["use_geoembed"][0]:
            geo = ognno.geo_embedding(pos, lat, ei, sd["encoder.geoembed.mlp.0.weight"], sd["encoder.geoembed.mlp.0.bias"],
                                      sd["encoder.geoembed.mlp.2.weight"], sd["encoder.geoembed.mlp.2.bias"])
            enc = _mlp(sd, "encoder.recovery", torch.cat([enc, geo], dim=-1))
        enc_scales.append(enc)
    rn = enc_scales[0] if len(enc_scales) == 1 else torch.stack(enc_scales).sum(0)
    # ---- process (gaot_3d.py:166-222)
    D, H, W = cfg["latent_tokens"]
    P = cfg["patch_size"]
    nd, nh, nw = D // P, H // P, W // P
    x = rn.view(1, nd, P, nh, P, nw, P, C).permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(1, nd * nh * nw, P ** 3 * C)
    x = _lin(sd, "patch_linear", x)
    rope = cfg["positional_embedding"] == "rope"
    if not rope:
        ax = [torch.arange(n, dtype=torch.float32) for n in (nd, nh, nw)]
        ppos = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
        x = x + _abs_pe(ppos, P ** 3 * C)
    if "processor.input_proj.weight" in sd:
        x = _lin(sd, "processor.input_proj", x)
    nl = cfg["num_layers"]
    skips = []
    for i in range(nl // 2):
        x = _block(sd, f"processor.encoder_layers.{i}", x, cfg, rope)
        skips.append(x)
    if nl % 2 == 1:
        x = _block(sd, "processor.middle_layer", x, cfg, rope)
    for i in range(nl // 2):
        x = _block(sd, f"processor.decoder_layers.{i}", x, cfg, rope, skip=skips.pop())
    if "processor.output_proj.weight" in sd:
        x = _lin(sd, "processor.output_proj", x)
    x = x.view(1, nd, nh, nw, P, P, P, C).permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(D * H * W, C)
    # ---- decoder (magno.py:691-798)
    dec_scales = []
    for si, scale in enumerate(cfg.get("scales", [1.0])):
        if dec_edges is None:
            ei = torch.from_numpy(ograph.get_neighbor_strategy_np(cfg["dec_strategy"], pos.numpy(), None, lat.numpy(), None,
                                                                  cfg["radius"] * scale, cfg["k"], True))
        else:
            ei = dec_edges[si]
        w, b = _mlp_wb(sd, "decoder.gno.channel_mlp")
        dec = ognno.integral_transform(lat, pos, ei, x, w, b)
        if cfg["use_geoembed"][1]:
            geo = ognno.geo_embedding(lat, pos, ei, sd["decoder.geoembed.mlp.0.weight"], sd["decoder.geoembed.mlp.0.bias"],
                                      sd["decoder.geoembed.mlp.2.weight"], sd["decoder.geoembed.mlp.2.bias"])
            dec = _mlp(sd, "decoder.recovery", torch.cat([dec, geo], dim=-1))
        dec_scales.append(dec)
    out = dec_scales[0] if len(dec_scales) == 1 else torch.stack(dec_scales).sum(0)
    return _mlp(sd, "decoder.projection", out)
