"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the whole GAOT3D forward from a state_dict.

Follows reference src/model/gaot_3d.py:248-332 (forward), :166-222 (process: patchify, patch_linear,
positional embedding, Transformer, un-patchify), src/model/layers/magno.py:468-600 / :691-798
(encoder / decoder orchestration incl. the per-scale loop :502/:711, the scale aggregation :586-596 /
:780-790 and both mlp_type flavours -- a kernel-size-1 Conv1d weight [out,in,1] is the Linear weight
[out,in]) and src/model/layers/attn.py:205-230, :298-325 (block and U-Net skip wiring).  Pinned against
the reference's own GAOT3D in tests/test_oracle_vs_reference.py and against tests/golden/model_*.pt.

Two evaluation modes besides the plain fp32 one:
  dtype=torch.float64         -- the "truth" the tolerance tiers are measured against;
  emulate_bf16=True           -- the YARDSTICK of the mixed-precision tier: the same math in fp64 with every
                                 tensor-core operand of the transformer (weights, GEMM inputs, q/k/v after RoPE,
                                 probabilities, and the gradients that feed a GEMM) rounded to bf16 at the point
                                 where an ideal bf16-operand / fp32-accumulate implementation rounds it.  A kernel
                                 that is as good as the arithmetic it is allowed to use agrees with this to well
                                 below rtol 2e-2 even where bf16 itself moves the fp32 answer by more (q/k
                                 projection gradients: dS = P * (dP - D) is a difference of nearly equal terms).
"""
import math
import torch
import torch.nn.functional as F

from . import gno as ognno
from . import graph as ograph
from .attn import attention_core
from .rope import RotaryEmbedding


class _RoundBoth(torch.autograd.Function):
    """bf16 rounding of a tensor-core operand: the value on the way forward, its gradient on the way back
    (the backward GEMMs read the bf16 copy of the incoming gradient too)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBothWays(torch.autograd.Function):
    """operand rounded forward AND its own gradient rounded backward (dX GEMMs that write a bf16 result)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundGrad(torch.autograd.Function):
    """identity forward; the gradient flowing back is rounded to bf16 (it is about to be a GEMM operand)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _rb(x, on):
    return _RoundBoth.apply(x) if on else x


def _rg(x, on):
    return _RoundGrad.apply(x) if on else x


def _w2(w):
    return w.reshape(w.shape[0], -1)            # Conv1d(k=1) [out,in,1] == Linear [out,in]


def _lin(sd, name, x, bias=True, emu=False, bf16_out=False, bf16_dx=False):
    """nn.Linear.  emu: operands (x, W) rounded to bf16, the output gradient rounded before it feeds dX / dW;
    bf16_out / bf16_dx: the layer's output / input gradient is itself stored in bf16 (tblock.py hand-offs)."""
    w = _w2(sd[name + ".weight"])
    b = sd[name + ".bias"] if bias and (name + ".bias") in sd else None
    xin = x if not emu else (_RoundBothWays.apply(x) if bf16_dx else _RoundBoth.apply(x))
    y = F.linear(xin, _rb(w, emu), b)
    if emu and bf16_out:
        y = _RoundBoth.apply(y)
    return _rg(y, emu)


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
        w.append(_w2(sd[f"{prefix}.fcs.{n}.weight"])); b.append(sd[f"{prefix}.fcs.{n}.bias"]); n += 1
    return w, b


def _rms(x, w, eps):
    return (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)) * w


def _attention(q, k, v, H, Hkv, rope, emu):
    if not emu:
        return attention_core(q, k, v, H, Hkv, rope, dtype=q.dtype)
    # the kernels' operand roundings: q*scale*log2e, k, v after RoPE -> bf16; P -> bf16; dO, dS -> bf16
    B, S, HD = q.shape
    d = HD // H
    q = q.view(B, S, H, d).transpose(1, 2)
    k = k.view(B, S, Hkv, d).transpose(1, 2)
    v = v.view(B, S, Hkv, d).transpose(1, 2)
    rep = H // Hkv
    if rep != 1:
        k, v = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
    if rope:
        re = RotaryEmbedding(d).to(device=q.device, dtype=q.dtype)
        q, k = re.rotate_queries_or_keys(q), re.rotate_queries_or_keys(k)
    q, k, v = _rb(q / math.sqrt(d), True), _rb(k, True), _rb(v, True)
    s = _rg(q @ k.transpose(-1, -2), True)                          # dS is a bf16 operand of dQ / dK
    p = _rb(torch.softmax(s, dim=-1), True)
    o = _rg(p @ v, True)                                            # dO is a bf16 operand of dV / dP
    return o.transpose(1, 2).contiguous().view(B, S, HD)


def _block(sd, p, x, cfg, rope, skip=None, emu=False):
    if skip is not None:                                                   # attn.py:222-224
        x = _lin(sd, p + ".skip_proj", torch.cat([x, skip], dim=-1), emu=emu)
    h = _rms(x, sd[p + ".attn_norm.weight"], cfg["norm_eps"])              # :226
    q, k, v = (_lin(sd, f"{p}.attn.{n}_proj", h, bias=False, emu=emu, bf16_out=True) for n in "qkv")
    a = _attention(q, k, v, cfg["num_heads"], cfg["num_kv_heads"], rope, emu)
    h = x + _lin(sd, p + ".attn.o_proj", a, bias=False, emu=emu, bf16_dx=True)   # :227
    h = _rms(h, sd[p + ".ffn_norm.weight"], cfg["norm_eps"])               # :228 (residual after the norm)
    g = F.silu(_lin(sd, p + ".ffn.w1", h, bias=False, emu=emu, bf16_out=True)) * \
        _lin(sd, p + ".ffn.w3", h, bias=False, emu=emu, bf16_out=True)
    return h + _lin(sd, p + ".ffn.w2", g, bias=False, emu=emu, bf16_dx=True)     # :229


def _abs_pe(positions, embed_dim):
    half = embed_dim // 2
    freq = 1 / 10000 ** (2 * torch.arange(0, half, dtype=torch.float32) / embed_dim)
    ang = positions[:, :, None] * freq[None, None, :]
    pe = torch.zeros(positions.shape[0], embed_dim)
    pe[:, 0::2] = torch.sin(ang).sum(1)
    pe[:, 1::2] = torch.cos(ang).sum(1)
    return pe


def _scale_aggregate(sd, prefix, per_scale, pos, use_w):
    """magno.py:586-596 / :780-790."""
    if len(per_scale) == 1:
        return per_scale[0]
    stack = torch.stack(per_scale, dim=0)
    if use_w:
        w = F.linear(F.relu(F.linear(pos, sd[f"{prefix}.scale_weighting.0.weight"], sd[f"{prefix}.scale_weighting.0.bias"])),
                     sd[f"{prefix}.scale_weighting.2.weight"], sd[f"{prefix}.scale_weighting.2.bias"])
        w = torch.softmax(w, dim=-1)
        return (stack * w.permute(1, 0).unsqueeze(-1)).sum(0)
    return stack.sum(0)


def gaot3d_forward(sd, cfg, pos, feats, latent_pos=None, enc_edges=None, dec_edges=None, keep_graph=False,
                   batch_idx=None, num_graphs=1, dtype=torch.float32, emulate_bf16=False):
    """cfg: dict(latent_tokens, patch_size, lifting_channels, radius, k, enc_strategy, dec_strategy,
    use_geoembed=(enc,dec), num_layers, num_heads, num_kv_heads, norm_eps, positional_embedding[, scales,
    use_scale_weights]).  `enc_edges` / `dec_edges`: one explicit [2,E] edge_index per scale (global indices; the
    precomputed-edge path magno.py:506-516 or a masked graph) instead of the online build.  `batch_idx` [N] sorted
    example index of every point with `num_graphs` examples; the latent grid is repeated per example
    (gaot_3d.py:278-290).  transform 'linear'; both mlp_type flavours (same matrices)."""
    sd = {k: (v if keep_graph else v.detach()).cpu() for k, v in sd.items()}
    sd = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    pos32 = pos.float().cpu()
    lat32 = (sd["latent_tokens"] if latent_pos is None else latent_pos).detach().float().cpu()
    B = int(num_graphs)
    M1 = lat32.shape[0]
    lat32 = lat32.repeat(B, 1)
    bphys = None if batch_idx is None else batch_idx.cpu().numpy()
    blat = None if B == 1 else torch.arange(B).repeat_interleave(M1).numpy()
    pos_d, lat_d = pos32.to(dtype), lat32.to(dtype)
    C = cfg["lifting_channels"]
    scales = cfg.get("scales", [1.0])
    use_w = bool(cfg.get("use_scale_weights", False))
    # ---- encoder (magno.py:468-600)
    phys_feat = torch.cat([f.float().cpu().to(dtype) for f in feats], dim=-1)
    enc_scales = []
    for si, scale in enumerate(scales):
        if enc_edges is None:
            ei = torch.from_numpy(ograph.get_neighbor_strategy_np(cfg["enc_strategy"], pos32.numpy(), bphys, lat32.numpy(), blat,
                                                                  cfg["radius"] * scale, cfg["k"], False))
        else:
            ei = enc_edges[si].long().cpu()
        lifted = _mlp(sd, "encoder.lifting", phys_feat)
        w, b = _mlp_wb(sd, "encoder.gno.channel_mlp")
        enc = ognno.integral_transform(pos_d, lat_d, ei, lifted, w, b)
        if cfg["use_geoembed"][0]:
            geo = ognno.geo_embedding(pos_d, lat_d, ei, sd["encoder.geoembed.mlp.0.weight"], sd["encoder.geoembed.mlp.0.bias"],
                                      sd["encoder.geoembed.mlp.2.weight"], sd["encoder.geoembed.mlp.2.bias"])
            enc = _mlp(sd, "encoder.recovery", torch.cat([enc, geo], dim=-1))
        enc_scales.append(enc)
    rn = _scale_aggregate(sd, "encoder", enc_scales, lat_d, use_w)
    # ---- process (gaot_3d.py:166-222)
    D, H, W = cfg["latent_tokens"]
    P = cfg["patch_size"]
    nd, nh, nw = D // P, H // P, W // P
    x = rn.view(B, nd, P, nh, P, nw, P, C).permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(B, nd * nh * nw, P ** 3 * C)
    emu = bool(emulate_bf16)
    x = _lin(sd, "patch_linear", x, emu=emu)
    rope = cfg["positional_embedding"] == "rope"
    if not rope:
        ax = [torch.arange(n, dtype=torch.float32) for n in (nd, nh, nw)]
        ppos = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
        x = x + _abs_pe(ppos, P ** 3 * C).to(dtype)
    if "processor.input_proj.weight" in sd:
        x = _lin(sd, "processor.input_proj", x, emu=emu)
    nl = cfg["num_layers"]
    skips = []
    for i in range(nl // 2):
        x = _block(sd, f"processor.encoder_layers.{i}", x, cfg, rope, emu=emu)
        skips.append(x)
    if nl % 2 == 1:
        x = _block(sd, "processor.middle_layer", x, cfg, rope, emu=emu)
    for i in range(nl // 2):
        x = _block(sd, f"processor.decoder_layers.{i}", x, cfg, rope, skip=skips.pop(), emu=emu)
    if "processor.output_proj.weight" in sd:
        x = _lin(sd, "processor.output_proj", x, emu=emu)
    x = x.view(B, nd, nh, nw, P, P, P, C).permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B * D * H * W, C)
    # ---- decoder (magno.py:691-798)
    dec_scales = []
    for si, scale in enumerate(scales):
        if dec_edges is None:
            ei = torch.from_numpy(ograph.get_neighbor_strategy_np(cfg["dec_strategy"], pos32.numpy(), bphys, lat32.numpy(), blat,
                                                                  cfg["radius"] * scale, cfg["k"], True))
        else:
            ei = dec_edges[si].long().cpu()
        w, b = _mlp_wb(sd, "decoder.gno.channel_mlp")
        dec = ognno.integral_transform(lat_d, pos_d, ei, x, w, b)
        if cfg["use_geoembed"][1]:
            geo = ognno.geo_embedding(lat_d, pos_d, ei, sd["decoder.geoembed.mlp.0.weight"], sd["decoder.geoembed.mlp.0.bias"],
                                      sd["decoder.geoembed.mlp.2.weight"], sd["decoder.geoembed.mlp.2.bias"])
            dec = _mlp(sd, "decoder.recovery", torch.cat([dec, geo], dim=-1))
        dec_scales.append(dec)
    out = _scale_aggregate(sd, "decoder", dec_scales, pos_d, use_w)
    return _mlp(sd, "decoder.projection", out)
