"""TEST INFRASTRUCTURE ONLY -- restatement of ``rotary_embedding_torch.RotaryEmbedding``
(un-vendored, unpinned third-party package; reference call sites
src/model/layers/attn.py:7,87,119-120).  Published behaviour restated per
SURVEY.md Appendix A6: theta=10000, freqs_for='lang', learned_freq=False,
``freqs = 1/theta**(arange(0,dim,2)/dim)`` kept as a non-trainable
``nn.Parameter`` named ``freqs`` (state_dict key), positions 0..S-1 along dim -2,
angles repeated interleaved (f0,f0,f1,f1,...), ``out = t*cos + rotate_half(t)*sin``
with rotate_half on adjacent pairs (x0,x1)->(-x1,x0); all in fp32.
"""
import torch
import torch.nn as nn


def rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def apply_rotary_emb(freqs, t, start_index=0, scale=1.0, seq_dim=-2):
    rot_dim = freqs.shape[-1]
    end = start_index + rot_dim
    tl, tm, tr = t[..., :start_index], t[..., start_index:end], t[..., end:]
    tm = (tm * freqs.cos() * scale) + (rotate_half(tm) * freqs.sin() * scale)
    return torch.cat((tl, tm, tr), dim=-1).to(t.dtype)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000.0):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def forward(self, seq):
        f = torch.einsum("..., f -> ... f", seq.to(self.freqs.dtype), self.freqs)
        return f.repeat_interleave(2, dim=-1)

    def rotate_queries_or_keys(self, t, seq_dim=-2, offset=0):
        seq_len = t.shape[seq_dim]
        seq = torch.arange(seq_len, device=t.device, dtype=torch.float32) + offset
        return apply_rotary_emb(self.forward(seq), t, seq_dim=seq_dim)
