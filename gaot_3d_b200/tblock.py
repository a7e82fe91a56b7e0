"""One autograd node per latent TransformerBlock (reference src/model/layers/attn.py:205-230).

    [x = skip_proj(cat[x, skip])]; h = x + o_proj(attn(RMSNorm(x))); h = RMSNorm(h); out = h + w2(silu(w1 h) * w3 h)

The forward is 11 launches of this library's kernels, the backward 19; no torch op runs in between.  The residual
stream (x, h, normalised h, out and their gradients) stays fp32; everything that only feeds a GEMM or the
attention core is handed over in bf16, written by the producing kernel in the layout the consumer reads:

  rmsnorm_fwd -> h1 bf16 -> [Wq;Wk;Wv] GEMM -> qkv bf16 -> pack (+RoPE) -> tcgen05 attention -> o bf16
  -> o_proj GEMM (+x in the epilogue) -> h fp32 -> rmsnorm_fwd -> h2 fp32 + bf16 -> [W1;W3] GEMM -> gu bf16
  -> swiglu -> a bf16 -> w2 GEMM (+h2 in the epilogue) -> out fp32

Weights are cast to bf16 once per call (they change every optimizer step); q/k/v and w1/w3 are cast into one
concatenated operand so that each pair of projections is one GEMM.  Weight gradients come out of split-K GEMMs
with a fixed-order reduction, so the whole block is bit-reproducible except for the attention dQ reductions.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import ops
from .ops import _lib_, _p, _stream, _ws, check, _linear_fwd_raw, _linear_bwd_x_raw, _linear_bwd_w_raw

BF16 = torch.bfloat16

# Head-parallel attention for the intra-sample sharded mode (shard.py): the transformer is replicated on every rank, but its
# attention core -- two thirds of its time at S = 16384 -- splits by heads with no communication before it (q/k/v of a head
# come from the replicated activations): rank r runs heads [r H/R, (r+1) H/R), the head outputs are all-gathered (8 MB bf16
# per layer) before the replicated o_proj, and in the backward each rank differentiates its heads and dqkv is all-gathered.
#
# Sequence-parallel mode ("sp", round 2): the whole block runs on this rank's S/R token slab -- norms, every GEMM, SwiGLU --
# and only the attention core needs the full sequence: the bf16 [q|k|v] projections go through ONE all-to-all into the
# head-sharded layout (rank r receives all S tokens of its H/R heads; the projection weight rows are grouped by destination
# rank when they are cast, so the send buffer is a plain [R, S/R, W] transpose of the GEMM output), the tcgen05 attention
# kernels run unchanged on [S, H/R heads], and a second all-to-all brings the head outputs back to token slabs.  The backward
# mirrors it (dO forth, d[q|k|v] back).  Nothing is replicated; weight gradients are partial sums over the local tokens and
# are all-reduced once per step by the caller (shard.allreduce_partial_grads).
_HEAD_PARALLEL = {"group": None, "on": False, "mode": "hp"}


def set_head_parallel(on: bool, group=None, mode: str = "hp") -> None:
    _HEAD_PARALLEL["on"], _HEAD_PARALLEL["group"], _HEAD_PARALLEL["mode"] = bool(on), group, mode


def _head_shard(num_heads: int, num_kv_heads: int):
    if not _HEAD_PARALLEL["on"]:
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    grp = _HEAD_PARALLEL["group"]
    R = dist.get_world_size(grp)
    if R == 1 or num_heads % R or num_kv_heads % R:
        if _HEAD_PARALLEL["mode"] == "sp" and R > 1:
            raise NotImplementedError("sequence-parallel transformer: head counts must be divisible by the number of ranks")
        return None
    return dist.get_rank(grp), R, grp, _HEAD_PARALLEL["mode"]


def _all_to_all(send: torch.Tensor, grp) -> torch.Tensor:
    """[R, ...] -> [R, ...]: block j goes to rank j, block i of the result came from rank i.  Over NVLink peer memory (p2p.py:
    one store kernel + the symmetric allocation's barrier) when the node allows it, else ncclSend/ncclRecv groups.  The peer
    path returns a view of a symmetric buffer that the second next exchange overwrites: every caller below consumes or copies
    it before then."""
    import torch.distributed as dist
    from . import p2p
    out = p2p.all_to_all(send, grp)
    if out is not None:
        return out
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=grp)
    return recv


def a2a_backend(grp=None) -> str:
    from . import p2p
    return p2p.backend(grp)


def head_slices(qkv: torch.Tensor, r: int, R: int, H: int, Hkv: int, d: int) -> torch.Tensor:
    """[M, (H + 2 Hkv) d] = [q | k | v] of all heads -> the contiguous [q_l | k_l | v_l] block of rank r's heads
    (heads [r H/R, (r+1) H/R) and their kv heads [r Hkv/R, (r+1) Hkv/R): contiguous blocks keep the GQA grouping)."""
    nq, nkv, Ha, Hkva = H * d, Hkv * d, H // R, Hkv // R
    return torch.cat([qkv[:, r * Ha * d:(r + 1) * Ha * d], qkv[:, nq + r * Hkva * d: nq + (r + 1) * Hkva * d],
                      qkv[:, nq + nkv + r * Hkva * d: nq + nkv + (r + 1) * Hkva * d]], dim=1)


def merge_head_slices(parts, H: int, Hkv: int, d: int) -> torch.Tensor:
    """Inverse of head_slices over all ranks: per-rank [q_l | k_l | v_l] blocks -> [q | k | v] of all heads."""
    R = len(parts)
    qa, ka = (H // R) * d, (Hkv // R) * d
    return torch.cat([p_[:, :qa] for p_ in parts] + [p_[:, qa:qa + ka] for p_ in parts] + [p_[:, qa + ka:] for p_ in parts], dim=1)


def _gather_cols(local: torch.Tensor, R: int, grp) -> list:
    import torch.distributed as dist
    parts = [torch.empty_like(local) for _ in range(R)]
    dist.all_gather(parts, local.contiguous(), group=grp)
    return parts


def _cast_into(src: torch.Tensor, dst: torch.Tensor) -> None:
    """fp32 parameter -> bf16 operand (dst is a contiguous row slice of the concatenated operand)."""
    lib = _lib_()
    s = src.detach()
    if s.dtype != torch.float32 or not s.is_contiguous():
        s = s.to(torch.float32).contiguous()
    check(lib.gaot_cast_bf16(_p(s), _p(dst), s.numel(), _stream(s.device)), "cast_bf16")


def _cast_many(pairs) -> None:
    """[(fp32 parameter, bf16 destination view)] -> one launch per <= 8 casts (the weights of a block are cast once per call)."""
    lib = _lib_()
    srcs = []
    for s, d in pairs:
        s = s.detach()
        if s.dtype != torch.float32 or not s.is_contiguous():
            s = s.to(torch.float32).contiguous()
        assert d.is_contiguous() and d.numel() == s.numel()
        srcs.append((s, d))
    for i in range(0, len(srcs), 8):
        chunk = srcs[i:i + 8]
        n = len(chunk)
        a_src = (ctypes.c_void_p * n)(*[s.data_ptr() for s, _ in chunk])
        a_dst = (ctypes.c_void_p * n)(*[d.data_ptr() for _, d in chunk])
        a_n = (ctypes.c_int64 * n)(*[s.numel() for s, _ in chunk])
        check(lib.gaot_cast_bf16_batch(a_src, a_dst, a_n, n, _stream(chunk[0][0].device)), "cast_bf16_batch")


def _cast(src: torch.Tensor) -> torch.Tensor:
    dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    _cast_into(src, dst)
    return dst


def _rmsnorm_fwd(x, w, eps, want_f32: bool):
    lib = _lib_()
    M, H = x.shape
    dev = x.device
    yb = torch.empty(M, H, dtype=BF16, device=dev)
    yf = torch.empty(M, H, dtype=torch.float32, device=dev) if want_f32 else None
    rstd = torch.empty(M, dtype=torch.float32, device=dev)
    check(lib.gaot_rmsnorm_forward(_p(x), _p(w), M, H, float(eps), _p(yb), _p(yf), _p(rstd), _stream(dev)), "rmsnorm_forward")
    return yb, yf, rstd


def _rmsnorm_bwd(dy, x, rstd, w, dres):
    lib = _lib_()
    M, H = x.shape
    dev = x.device
    dx = torch.empty(M, H, dtype=torch.float32, device=dev)
    dw = torch.empty(H, dtype=torch.float32, device=dev)
    wsb = lib.gaot_rmsnorm_backward_workspace_bytes(H)
    ws = _ws(wsb, dev)
    check(lib.gaot_rmsnorm_backward(_p(dy), _p(x), _p(rstd), _p(w), _p(dres), M, H, _p(dx), _p(dw), _p(ws), wsb,
                                    _stream(dev)), "rmsnorm_backward")
    return dx, dw


def _colsum(x):
    lib = _lib_()
    M, N = x.shape
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    wsb = lib.gaot_colsum_workspace_bytes(M, N)
    ws = _ws(wsb, x.device)
    check(lib.gaot_colsum(_p(x), M, N, _p(out), _p(ws), wsb, _stream(x.device)), "colsum")
    return out


def block_supported(hidden: int, ffn_hidden: int, num_heads: int, num_kv_heads: int) -> bool:
    d = hidden // max(num_heads, 1)
    return (hidden % 128 == 0 and 128 <= hidden <= 1024 and d in (32, 64) and num_heads * d == hidden
            and num_heads % num_kv_heads == 0 and ffn_hidden % 64 == 0)


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, skip, cfg, skip_w, skip_b, n1_w, wq, wk, wv, wo, n2_w, w1, w2, w3):
        lib = _lib_()
        H, Hkv, eps, freqs, p_drop, seed, hp = cfg
        shape = x.shape
        S, Hd = shape[-2], shape[-1]
        B = x.numel() // (S * Hd)
        M = B * S
        d = Hd // H
        dev = x.device
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            x2d = f32(x).view(M, Hd)
            # every weight of the block -> bf16 operands in ONE launch (q/k/v and w1/w3 land in concatenated operands)
            nq, nkv, F = H * d, Hkv * d, w1.shape[0]
            wqkv = torch.empty(nq + 2 * nkv, Hd, dtype=BF16, device=dev)
            w13 = torch.empty(2 * F, Hd, dtype=BF16, device=dev)
            wob = torch.empty(wo.shape, dtype=BF16, device=dev)
            w2b = torch.empty(w2.shape, dtype=BF16, device=dev)
            wsk = torch.empty(skip_w.shape, dtype=BF16, device=dev) if skip is not None else None
            sp = hp is not None and hp[3] == "sp"
            if sp:      # projection rows grouped by destination rank: [q heads of r | k heads of r | v heads of r] for r = 0..R-1
                assert B == 1, "sequence-parallel block: one example per step"
                R_ = hp[1]
                qa, ka = nq // R_, nkv // R_
                Wd = qa + 2 * ka
                qkv_casts = []
                for j in range(R_):
                    qkv_casts += [(wq[j * qa:(j + 1) * qa], wqkv[j * Wd:j * Wd + qa]), (wk[j * ka:(j + 1) * ka], wqkv[j * Wd + qa:j * Wd + qa + ka]),
                                  (wv[j * ka:(j + 1) * ka], wqkv[j * Wd + qa + ka:(j + 1) * Wd])]
            else:
                qkv_casts = [(wq, wqkv[:nq]), (wk, wqkv[nq:nq + nkv]), (wv, wqkv[nq + nkv:])]
            _cast_many(qkv_casts + [(wo, wob), (w1, w13[:F]), (w3, w13[F:]), (w2, w2b)] + ([(skip_w, wsk)] if skip is not None else []))
            if skip is not None:
                s2d = f32(skip).view(M, -1)
                x_in = _linear_fwd_raw(x2d, s2d, wsk, f32(skip_b) if skip_b is not None else None, None)
            else:
                s2d = None
                x_in = x2d
            n1, n2 = f32(n1_w), f32(n2_w)
            h1, _, rstd1 = _rmsnorm_fwd(x_in, n1, eps, False)
            qkv = _linear_fwd_raw(h1, None, wqkv, None, None, BF16)
            Ha, Hkva = H, Hkv                                                   # heads this rank runs through the attention core
            Sa, Ma = S, M                                                       # sequence length / rows the attention core sees
            if sp:
                r, R, grp, _ = hp
                Ha, Hkva = H // R, Hkv // R
                Sa = Ma = R * M                                                 # every rank holds S/R tokens; the core sees all S
                qkv = _all_to_all(qkv.view(M, R, Wd).transpose(0, 1).contiguous(), grp).view(Ma, Wd)
            elif hp is not None:
                r, R, grp, _ = hp
                Ha, Hkva = H // R, Hkv // R
                qkv = head_slices(qkv, r, R, H, Hkv, d)
            packed = torch.empty(lib.gaot_attn_packed_bytes(B, Sa, Ha, Hkva, d), dtype=torch.uint8, device=dev)
            o = torch.empty(Ma, Ha * d, dtype=BF16, device=dev)
            o32 = torch.empty(Ma, Ha * d, dtype=torch.float32, device=dev)   # unrounded copy: the backward's D = rowsum(dO * O)
            lse = torch.empty(B, Ha, Sa, dtype=torch.float32, device=dev)
            fr = None if freqs is None else f32(freqs)
            with ops._timed("attn_fwd", dev):
                check(lib.gaot_attn_fused_forward(_p(qkv), qkv.stride(0), B, Sa, Ha, Hkva, d, _p(fr), float(p_drop), int(seed),
                                                  _p(packed), _p(o), _p(o32), _p(lse), _stream(dev)), "attn_fused_forward")
            del qkv
            if sp:                                                              # head outputs back to token slabs: [M, H*d]
                o = _all_to_all(o.view(R, M, Ha * d), grp).transpose(0, 1).reshape(M, H * d)
            elif hp is not None:
                o = torch.cat(_gather_cols(o, hp[1], hp[2]), dim=1)            # [M, H*d]: every rank's heads, in head order
            h = _linear_fwd_raw(o, None, wob, None, x_in)                       # x + attn(norm(x))
            h2b, h2, rstd2 = _rmsnorm_fwd(h, n2, eps, True)
            gu = _linear_fwd_raw(h2b, None, w13, None, None, BF16)
            a = torch.empty(M, F, dtype=BF16, device=dev)
            check(lib.gaot_swiglu_forward(_p(gu), M, F, _p(a), _stream(dev)), "swiglu_forward")
            out = _linear_fwd_raw(a, None, w2b, None, h2)                       # h2 + ffn(h2)
        ctx.save_for_backward(x2d, s2d, x_in, rstd1, h1, packed, o, o32, lse, h, rstd2, h2b, gu, a,
                              wsk, wqkv, wob, w13, w2b, n1, n2, fr)
        ctx.meta = (shape, None if skip is None else skip.shape, B, S, Hd, H, Hkv, d, F, float(p_drop), int(seed),
                    skip_b is not None, hp)
        return out.view(shape)

    @staticmethod
    def backward(ctx, dout):
        (x2d, s2d, x_in, rstd1, h1, packed, o, o32, lse, h, rstd2, h2b, gu, a, wsk, wqkv, wob, w13, w2b, n1, n2, fr) = ctx.saved_tensors
        shape, skip_shape, B, S, Hd, H, Hkv, d, F, p_drop, seed, has_skip_bias, hp = ctx.meta
        lib = _lib_()
        M = B * S
        dev = dout.device
        f32 = torch.float32
        with torch.cuda.device(dev):
            dout = dout.to(f32).contiguous().view(M, Hd)
            # ---- FFN: out = h2 + w2(silu(w1 h2) * w3 h2)
            # gradients of the fp32 residual stream feed two GEMMs each as an OPERAND: one bf16 copy (the rounding the
            # tensor core applies anyway) lets both run on the TMA-fed kernel
            dout_b = _cast(dout)
            da = _linear_bwd_x_raw(dout_b, w2b, BF16)
            dw2 = torch.empty(Hd, F, dtype=f32, device=dev)
            _linear_bwd_w_raw(dout_b, a, dw2)
            del dout_b
            dgu = torch.empty(M, 2 * F, dtype=BF16, device=dev)
            check(lib.gaot_swiglu_backward(_p(da), _p(gu), M, F, _p(dgu), _stream(dev)), "swiglu_backward")
            del da
            dh2 = _linear_bwd_x_raw(dgu, w13, f32, residual=dout)
            dw13 = torch.empty(2 * F, Hd, dtype=f32, device=dev)
            _linear_bwd_w_raw(dgu, h2b, dw13)
            del dgu
            dh, dn2 = _rmsnorm_bwd(dh2, h, rstd2, n2, None)
            del dh2
            # ---- attention: h = x_in + o_proj(attn(norm(x_in)))
            dh_b = _cast(dh)
            do = _linear_bwd_x_raw(dh_b, wob, BF16)
            dwo = torch.empty(Hd, H * d, dtype=f32, device=dev)
            _linear_bwd_w_raw(dh_b, o, dwo)
            del dh_b
            nq, nkv = H * d, Hkv * d
            Ha, Hkva = H, Hkv
            Sa, Ma = S, M
            sp = hp is not None and hp[3] == "sp"
            if sp:
                r, R, grp, _ = hp
                Ha, Hkva = H // R, Hkv // R
                Sa = Ma = R * M
                do = _all_to_all(do.view(M, R, Ha * d).transpose(0, 1).contiguous(), grp).view(Ma, Ha * d)
            elif hp is not None:
                r, R, grp, _ = hp
                Ha, Hkva = H // R, Hkv // R
                do = do[:, r * Ha * d:(r + 1) * Ha * d].contiguous()
            Wd = (Ha + 2 * Hkva) * d
            dqkv = torch.empty(Ma, Wd, dtype=BF16, device=dev)
            wsb = lib.gaot_attn_fused_backward_workspace_bytes(B, Sa, Ha, d)
            ws = _ws(wsb, dev)
            with ops._timed("attn_bwd", dev):
                check(lib.gaot_attn_fused_backward(_p(packed), _p(o32), _p(do), _p(lse), B, Sa, Ha, Hkva, d, _p(fr), p_drop, seed,
                                                   _p(ws), wsb, _p(dqkv), dqkv.stride(0), _stream(dev)), "attn_fused_backward")
            del do, ws
            if sp:                                                              # d[q|k|v] of my heads, all tokens -> my tokens, all heads
                dqkv = _all_to_all(dqkv.view(R, M, Wd), grp).transpose(0, 1).reshape(M, R * Wd)
            elif hp is not None:                                                # [dq_l | dk_l | dv_l] of every rank -> [dq | dk | dv]
                dqkv = merge_head_slices(_gather_cols(dqkv, hp[1], hp[2]), H, Hkv, d)
            dh1 = _linear_bwd_x_raw(dqkv, wqkv, f32)
            dwqkv = torch.empty(nq + 2 * nkv, Hd, dtype=f32, device=dev)
            _linear_bwd_w_raw(dqkv, h1, dwqkv)
            del dqkv
            if sp:                                                              # rank-grouped rows -> [dWq; dWk; dWv]
                g3 = dwqkv.view(R, Wd, Hd)
                qa, ka = Ha * d, Hkva * d
                dwqkv = torch.cat([g3[:, :qa].reshape(nq, Hd), g3[:, qa:qa + ka].reshape(nkv, Hd), g3[:, qa + ka:].reshape(nkv, Hd)])
            dx_in, dn1 = _rmsnorm_bwd(dh1, x_in, rstd1, n1, dh)
            del dh1, dh
            dskip = dwsk = dbsk = None
            if s2d is not None:
                dcat = _linear_bwd_x_raw(dx_in, wsk, f32)
                k1 = x2d.shape[1]
                dx = dcat[:, :k1].reshape(shape)
                dskip = dcat[:, k1:].reshape(skip_shape)
                dwsk = torch.empty(wsk.shape, dtype=f32, device=dev)
                _linear_bwd_w_raw(dx_in, x2d, dwsk[:, :k1])
                _linear_bwd_w_raw(dx_in, s2d, dwsk[:, k1:])
                if has_skip_bias:
                    dbsk = _colsum(dx_in)
            else:
                dx = dx_in.view(shape)
        return (dx, dskip, None, dwsk, dbsk, dn1, dwqkv[:nq], dwqkv[nq:nq + nkv], dwqkv[nq + nkv:], dwo, dn2,
                dw13[:F], dw2, dw13[F:])


def transformer_block(x, skip, *, num_heads: int, num_kv_heads: int, eps: float, rope_freqs: Optional[torch.Tensor],
                      dropout_p: float, skip_w, skip_b, attn_norm_w, wq, wk, wv, wo, ffn_norm_w, w1, w2, w3,
                      seed: Optional[int] = None):
    """Fused TransformerBlock forward (see module docstring).  x [..., S, hidden] fp32; skip like x or None."""
    ops._need_cuda(x, skip)
    if dropout_p > 0.0 and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    cfg = (int(num_heads), int(num_kv_heads), float(eps), rope_freqs, float(dropout_p), int(seed or 0),
           _head_shard(int(num_heads), int(num_kv_heads)))
    return _BlockFn.apply(x, skip, cfg, skip_w, skip_b, attn_norm_w, wq, wk, wv, wo, ffn_norm_w, w1, w2, w3)
