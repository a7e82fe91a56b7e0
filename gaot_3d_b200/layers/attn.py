"""Latent-token transformer -- drop-in for reference src/model/layers/attn.py.

Config dataclasses keep the reference's field names/defaults; modules keep parameter names
(`attn.{q,k,v,o}_proj.weight`, `attn.rotary_emb.freqs`, `ffn.{w1,w2,w3}.weight`,
`attn_norm.weight`, `ffn_norm.weight`, `skip_proj.*`).  The attention core (RoPE + softmax(QK^T)V,
reference :110-128) is the tcgen05/TMEM flash kernel of this package; the projections, the SwiGLU
GEMMs and skip_proj run on the tcgen05 dense kernel (csrc/dense.cu) with the residual adds fused
into the GEMM epilogues.
"""
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops, tblock

# True: a TransformerBlock is ONE autograd node running only this library's kernels with bf16 hand-offs (tblock.py);
# False: module-by-module path (tcgen05 GEMMs and attention, torch RMSNorm / SiLU / adds in fp32)
FUSED_BLOCK = True


@dataclass
class AttentionConfig:
    hidden_size: int = 256
    num_heads: int = 8
    num_kv_heads: int = 8
    use_conditional_norm: bool = False
    cond_norm_hidden_size: int = 4
    atten_dropout: float = 0.1
    positional_embedding: str = "absolute"
    H: Optional[int] = None
    W: Optional[int] = None


@dataclass
class FFNConfig:
    hidden_size: int = 1024
    use_conditional_norm: bool = False
    cond_norm_hidden_size: int = 4


@dataclass
class TransformerConfig:
    patch_size: int = 8
    hidden_size: int = 256
    use_attn_norm: bool = True
    use_ffn_norm: bool = True
    norm_eps: float = 1e-6
    num_layers: int = 3
    positional_embedding: str = "absolute"
    use_long_range_skip: bool = True
    attn_config: AttentionConfig = field(default_factory=AttentionConfig)
    ffn_config: FFNConfig = field(default_factory=FFNConfig)


def _cfg_dict(obj) -> dict:
    if is_dataclass(obj):
        return {f.name: getattr(obj, f.name) for f in fields(obj)}
    return dict(obj)


def _lin(mod: nn.Linear, x, residual=None, x2=None):
    """nn.Linear forward on the tcgen05 dense kernel (parameters stay in the nn.Linear, so state_dict keys match)."""
    if x.is_cuda and ops.transformer_precision() == "fp32":      # strict tier: the reference's own fp32 library GEMM
        y = F.linear(x if x2 is None else torch.cat([x, x2], dim=-1), mod.weight, mod.bias)
        return y if residual is None else y + residual
    if x.is_cuda and ops.linear_supported(mod.in_features, mod.out_features) and (x2 is None or x.shape[-1] % 64 == 0):
        return ops.linear(x, mod.weight, mod.bias, residual=residual, x2=x2)
    if not x.is_cuda:
        raise RuntimeError("gaot_3d_b200: the B200 hot path has no CPU fallback; got a CPU tensor")
    # shapes outside the kernel envelope (rows not 16-byte aligned): node-level torch path, SURVEY 8f
    y = mod(x if x2 is None else torch.cat([x, x2], dim=-1))
    return y if residual is None else y + residual


class RotaryEmbedding(nn.Module):
    """Carrier of the `freqs` state_dict entry of rotary_embedding_torch.RotaryEmbedding(dim)
    (theta 10000, non-trainable nn.Parameter); the rotation itself is fused into the attention
    kernels' Q/K load."""

    def __init__(self, dim, theta=10000.0):
        super().__init__()
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)), requires_grad=False)


def _rope_fp32(t: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """1-D rotary embedding over the sequence axis of [B, h, S, d] (rotary_embedding_torch semantics as the reference uses
    them, attn.py:119-120: positions 0..S-1, angle f_i on the adjacent pair (2i, 2i+1), out = t cos + rotate_half(t) sin)."""
    S = t.shape[-2]
    ang = torch.arange(S, device=t.device, dtype=torch.float32)[:, None] * freqs.to(torch.float32)[None, :]     # [S, d/2]
    cos, sin = ang.cos().repeat_interleave(2, dim=-1), ang.sin().repeat_interleave(2, dim=-1)
    pair = t.reshape(*t.shape[:-1], t.shape[-1] // 2, 2)
    rot = torch.stack((-pair[..., 1], pair[..., 0]), dim=-1).reshape(t.shape)
    return t * cos + rot * sin


def _sdpa_fp32(q, k, v, num_heads, num_kv_heads, freqs, dropout_p):
    """Strict-FP32 tier of the attention core: the reference's own sequence of library calls (attn.py:110-128)."""
    B, S, _ = q.shape
    d = q.shape[-1] // num_heads
    q = q.view(B, S, num_heads, d).transpose(1, 2)
    k = k.view(B, S, num_kv_heads, d).transpose(1, 2)
    v = v.view(B, S, num_kv_heads, d).transpose(1, 2)
    if freqs is not None:
        q, k = _rope_fp32(q, freqs), _rope_fp32(k, freqs)
    if num_kv_heads != num_heads:
        rep = num_heads // num_kv_heads
        k, v = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=dropout_p)
    return o.transpose(1, 2).reshape(B, S, num_heads * d)


class GroupQueryFlashAttention(nn.Module):
    def __init__(self, input_size: int, output_size: int, hidden_size: int = 128, num_heads: int = 8,
                 num_kv_heads: int = 4, use_conditional_norm: bool = False, cond_norm_hidden_size: int = 4,
                 atten_dropout: float = 0.0, H: int = 64, W: int = 64, positional_embedding: str = "absolute"):
        super().__init__()
        assert hidden_size % num_heads == 0, f"hidden_size {hidden_size} must be divisible by num_heads {num_heads}"
        assert num_heads % num_kv_heads == 0, f"num_heads {num_heads} must be divisible by num_kv_heads {num_kv_heads}"
        self.num_heads, self.num_kv_heads = num_heads, num_kv_heads
        self.num_repeat = num_heads // num_kv_heads
        self.head_dim = hidden_size // num_heads
        self.atten_dropout = atten_dropout
        kv = self.head_dim * num_kv_heads
        self.q_proj = nn.Linear(input_size, hidden_size, bias=False)
        self.k_proj = nn.Linear(input_size, kv, bias=False)
        self.v_proj = nn.Linear(input_size, kv, bias=False)
        self.o_proj = nn.Linear(hidden_size, output_size, bias=False)
        if use_conditional_norm:
            raise NotImplementedError("time-conditional norm is unused by the 3-D static path (default_set.py:59)")
        self.correction = None
        if positional_embedding == "rope":
            self.rotary_emb = RotaryEmbedding(dim=self.head_dim)

    def forward(self, x, condition: Optional[float] = None, relative_positions: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None):
        """`residual` (extension): added to the o_proj output inside the GEMM epilogue (x + attn(x))."""
        lead = x.shape[:-2]
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        q, k, v = _lin(self.q_proj, x3), _lin(self.k_proj, x3), _lin(self.v_proj, x3)
        dp = self.atten_dropout if self.training else 0.0          # reference attn.py:122-126
        freqs = self.rotary_emb.freqs if relative_positions is not None else None
        if ops.transformer_precision() == "fp32":
            o = _sdpa_fp32(q, k, v, self.num_heads, self.num_kv_heads, freqs, dp)
            res3 = None if residual is None else residual.reshape(x3.shape[0], x3.shape[1], -1)
            return _lin(self.o_proj, o, residual=res3).reshape(*lead, x.shape[-2], -1)
        o = ops.attention(q, k, v, self.num_heads, self.num_kv_heads, rope_freqs=freqs, dropout_p=dp)
        res3 = None if residual is None else residual.reshape(x3.shape[0], x3.shape[1], -1)
        return _lin(self.o_proj, o, residual=res3).reshape(*lead, x.shape[-2], -1)

    @classmethod
    def from_config(cls, input_size: int, output_size: int, config: AttentionConfig):
        kw = _cfg_dict(config)
        for extra in ("D",):
            kw.pop(extra, None)
        return cls(input_size, output_size, **kw)


class FFN(nn.Module):
    def __init__(self, input_size: int, output_size: int, hidden_size: int = 256, use_conditional_norm: bool = False,
                 cond_norm_hidden_size: int = 4):
        super().__init__()
        self.w1 = nn.Linear(input_size, hidden_size, bias=False)
        self.w2 = nn.Linear(hidden_size, output_size, bias=False)
        self.w3 = nn.Linear(input_size, hidden_size, bias=False)
        if use_conditional_norm:
            raise NotImplementedError("time-conditional norm is unused by the 3-D static path")
        self.correction = None

    def forward(self, x, condition: Optional[float] = None, residual: Optional[torch.Tensor] = None):
        return _lin(self.w2, F.silu(_lin(self.w1, x)) * _lin(self.w3, x), residual=residual)

    @classmethod
    def from_config(cls, input_size: int, output_size: int, config: FFNConfig):
        return cls(input_size, output_size, **_cfg_dict(config))


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x) * self.weight


class TransformerBlock(nn.Module):
    def __init__(self, input_size: int, output_size: int, use_attn_norm: bool = True, use_ffn_norm: bool = True,
                 norm_eps: float = 1e-6, attn_config: AttentionConfig = None, ffn_config: FFNConfig = None,
                 skip_connection: bool = False):
        super().__init__()
        attn_config = attn_config or AttentionConfig()
        ffn_config = ffn_config or FFNConfig()
        self.attn = GroupQueryFlashAttention.from_config(input_size, attn_config.hidden_size, config=attn_config)
        self.ffn = FFN.from_config(attn_config.hidden_size, output_size, config=ffn_config)
        self.attn_norm = RMSNorm(input_size, eps=norm_eps) if use_attn_norm else None
        self.ffn_norm = RMSNorm(attn_config.hidden_size, eps=norm_eps) if use_ffn_norm else None
        self.skip_connection = skip_connection
        if skip_connection:
            self.skip_proj = nn.Linear(input_size + output_size, input_size)

    def forward(self, x, condition=None, relative_positions=None, skip=None):
        if self._fused_ok(x, condition):
            use_skip = self.skip_connection and skip is not None
            a, f = self.attn, self.ffn
            return tblock.transformer_block(
                x, skip if use_skip else None, num_heads=a.num_heads, num_kv_heads=a.num_kv_heads, eps=self.attn_norm.eps,
                rope_freqs=a.rotary_emb.freqs if (relative_positions is not None and hasattr(a, "rotary_emb")) else None,
                dropout_p=a.atten_dropout if self.training else 0.0,
                skip_w=self.skip_proj.weight if use_skip else None, skip_b=self.skip_proj.bias if use_skip else None,
                attn_norm_w=self.attn_norm.weight, wq=a.q_proj.weight, wk=a.k_proj.weight, wv=a.v_proj.weight,
                wo=a.o_proj.weight, ffn_norm_w=self.ffn_norm.weight, w1=f.w1.weight, w2=f.w2.weight, w3=f.w3.weight)
        if self.skip_connection and skip is not None:
            x = _lin(self.skip_proj, x, x2=skip)                  # Linear over cat[x, skip], read in place
        h = x if self.attn_norm is None else self.attn_norm(x)
        h = self.attn(h, condition=condition, relative_positions=relative_positions, residual=x)   # x + attn(h)
        h = h if self.ffn_norm is None else self.ffn_norm(h)      # reference quirk: residual taken after the norm
        return self.ffn(h, condition=condition, residual=h)       # h + ffn(h)

    def _fused_ok(self, x, condition) -> bool:
        """One-autograd-node path (tblock.py): both norms present, square block, shapes inside the kernel envelope."""
        a, f = self.attn, self.ffn
        hid = a.q_proj.in_features
        return (FUSED_BLOCK and ops.transformer_precision() == "bf16" and x.is_cuda and x.dtype == torch.float32 and self.attn_norm is not None and self.ffn_norm is not None
                and a.q_proj.out_features == hid and a.o_proj.out_features == hid and f.w2.out_features == hid
                and f.w1.in_features == hid and self.attn_norm.eps == self.ffn_norm.eps
                and (not self.skip_connection or self.skip_proj.in_features == 2 * hid)
                and tblock.block_supported(hid, f.w1.out_features, a.num_heads, a.num_kv_heads))

    @classmethod
    def from_config(cls, input_size: int, output_size: int, skip_connection: bool = False,
                    config: TransformerConfig = None):
        config = config or TransformerConfig()
        config.attn_config.positional_embedding = config.positional_embedding
        kw = _cfg_dict(config)
        for k in ("num_layers", "hidden_size", "positional_embedding", "use_long_range_skip", "patch_size"):
            kw.pop(k)
        return cls(input_size, output_size, skip_connection=skip_connection, **kw)


class Transformer(nn.Module):
    """U-ViT style stack: num_layers//2 encoder blocks, optional middle, num_layers//2 decoder
    blocks consuming the encoder outputs LIFO through `skip_proj` (reference attn.py:246-325)."""

    def __init__(self, input_size: int, output_size: int, config: TransformerConfig = None):
        super().__init__()
        config = config or TransformerConfig()
        hs, nl = config.hidden_size, config.num_layers
        self.use_long_range_skip = config.use_long_range_skip
        self.input_proj = nn.Linear(input_size, hs) if input_size != hs else nn.Identity()
        self.output_proj = nn.Linear(hs, output_size) if hs != output_size else nn.Identity()
        mk = lambda skip: TransformerBlock.from_config(input_size=hs, output_size=hs, skip_connection=skip, config=config)
        self.encoder_layers = nn.ModuleList(mk(False) for _ in range(nl // 2))
        self.middle_layer = mk(False) if nl % 2 == 1 else None
        self.decoder_layers = nn.ModuleList(mk(True) for _ in range(nl // 2))

    def forward(self, x, condition=None, relative_positions=None):
        x = x if isinstance(self.input_proj, nn.Identity) else _lin(self.input_proj, x)
        skips = []
        for layer in self.encoder_layers:
            x = layer(x, condition=condition, relative_positions=relative_positions)
            skips.append(x)
        if self.middle_layer is not None:
            x = self.middle_layer(x, condition=condition, relative_positions=relative_positions)
        for layer in self.decoder_layers:
            skip = skips.pop() if self.use_long_range_skip else None
            x = layer(x, condition=condition, relative_positions=relative_positions, skip=skip)
        return x if isinstance(self.output_proj, nn.Identity) else _lin(self.output_proj, x)
