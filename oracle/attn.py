"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the latent attention core.

Follows reference src/model/layers/attn.py:104-129 (GroupQueryFlashAttention.forward):
head split, GQA repeat_interleave, optional RoPE on q,k (1-D over the flattened
patch index, App. A6), softmax(QK^T/sqrt(d))V non-causal, no mask, merge heads.
Pinned against the reference module in tests/test_oracle_vs_reference.py.
"""
import math
import torch

from .rope import RotaryEmbedding


def attention_core(q, k, v, num_heads, num_kv_heads, rope: bool, dtype=torch.float32):
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
        re = RotaryEmbedding(d).to(dtype)
        q = re.rotate_queries_or_keys(q)
        k = re.rotate_queries_or_keys(k)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)                    # SDPA default scale
    p = torch.softmax(s, dim=-1)
    o = p @ v
    return o.transpose(1, 2).contiguous().view(B, S, HD)            # attn.py:128
