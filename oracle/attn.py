"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the latent attention core.

Follows reference src/model/layers/attn.py:104-129 (GroupQueryFlashAttention.forward):
head split, GQA repeat_interleave, optional RoPE on q,k (1-D over the flattened
patch index, App. A6), softmax(QK^T/sqrt(d))V non-causal, no mask, merge heads.
Pinned against the reference module in tests/test_oracle_vs_reference.py.
"""
import math
import torch

from .rope import RotaryEmbedding


def _lowbias32(x):
    import numpy as np
    x = x.astype(np.uint64)
    M = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    return x


def dropout_keep(B, H, S, p, seed):
    """Restatement of the kernels' counter-based dropout mask (csrc/attn.cu: DropCfg / drop_keep):
    keep[b,h,q,k] = lowbias32(rowkey(b,h,q) ^ k*0x85EBCA6B) >= p*2^32.  Returns bool [B,H,S,S]."""
    import numpy as np
    M = np.uint64(0xFFFFFFFF)
    thresh = np.uint64(min(int(p * 4294967296.0), 0xFFFFFFFF))
    rows = np.arange(B * H * S, dtype=np.uint64)
    rowkey = (_lowbias32(np.uint64(seed & 0xFFFFFFFF) ^ ((rows * np.uint64(0x9E3779B1)) & M)) + np.uint64(seed >> 32)) & M
    kterm = (np.arange(S, dtype=np.uint64) * np.uint64(0x85EBCA6B)) & M
    keep = _lowbias32(rowkey[:, None] ^ kterm[None, :]) >= thresh
    return torch.from_numpy(keep.reshape(B, H, S, S))


def attention_core(q, k, v, num_heads, num_kv_heads, rope: bool, dtype=torch.float32, dropout_p=0.0, seed=0):
    """q:[B,S,H*d] k,v:[B,S,Hkv*d] (outputs of the bias-free projections) -> [B,S,H*d]."""
    B, S, HD = q.shape
    d = HD // num_heads
    q = q.view(B, S, num_heads, d).transpose(1, 2).to(dtype)
    k = k.view(B, S, num_kv_heads, d).transpose(1, 2).to(dtype)
    v = v.view(B, S, num_kv_heads, d).transpose(1, 2).to(dtype)
    rep = num_heads // num_kv_heads
    if rep != 1:                                                    # attn.py:114-116
        k = k.repeat_interleave(rep, dim=1)
        v = v.repeat_interleave(rep, dim=1)
    if rope:                                                        # attn.py:118-120
        re = RotaryEmbedding(d).to(device=q.device, dtype=dtype)
        q = re.rotate_queries_or_keys(q)
        k = re.rotate_queries_or_keys(k)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)                    # SDPA default scale
    p = torch.softmax(s, dim=-1)
    if dropout_p > 0.0:                                             # attn.py:122-126 (dropout inside SDPA)
        p = p * dropout_keep(B, num_heads, S, dropout_p, seed).to(device=p.device, dtype=p.dtype) / (1.0 - dropout_p)
    o = p @ v
    return o.transpose(1, 2).contiguous().view(B, S, HD)            # attn.py:128
