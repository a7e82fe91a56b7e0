"""Intra-sample sharding of one large mesh across the GPUs of a node (new -- the
reference only has sample-level DDP, src/trainer/stat.py:431-436).

Partitioning: rank r owns the contiguous physical-point index range [r*N/R, (r+1)*N/R); latent tokens,
all weights and the transformer are replicated.
  encoder : local edges (local phys shard x all latents) -> GNO partial SUMS [M,C] + counts [M]
            -> NCCL all-reduce(SUM) -> mean.   The radius cap (first 32 per latent by ascending phys
            index, globally) is restored with one all-gather of per-latent counts: contiguous index
            ranges make "ascending index" compose as an exclusive prefix over ranks.
  decoder : queries (phys) sharded, sources (latents) replicated -> no forward communication.
  backward: d latent from the local decoder is a partial sum -> all-reduce before the replicated
            transformer backward; GNO-side parameter grads are partial -> all-reduce (allreduce_partial_grads).
One process per GPU; collectives go through torch.distributed (NCCL over NVLink/NVSwitch on the box,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from .graph import PYG_MAX_NUM_NEIGHBORS, knn_graph, radius_graph, _coalesced, _tag


def shard_range(n: int, rank: int, world: int):
    """Contiguous, near-equal index ranges; the first n % world ranks get one extra point."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _AllReduceFwd(torch.autograd.Function):
    """y = sum_r x_r (replicated result).  Backward: identity -- the incoming gradient is already the
    total because the consumer's gradient is all-reduced by _AllReduceBwd downstream."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.contiguous().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class _AllReduceBwd(torch.autograd.Function):
    """Identity forward on a replicated tensor; backward sums the per-rank partial gradients."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def all_reduce_forward(x, group=None):
    return _AllReduceFwd.apply(x, group)


def all_reduce_backward(x, group=None):
    return _AllReduceBwd.apply(x, group)


def apply_global_radius_cap(lat_idx: torch.Tensor, phys_idx: torch.Tensor, num_latent: int, cap: int, group=None):
    """Local radius edges (grouped by latent, ascending local phys index, locally capped at `cap`) ->
    the subset that survives the GLOBAL cap: edge ordinal (edges of lower ranks + local ordinal) < cap."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = torch.bincount(lat_idx, minlength=num_latent)
    allc = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    before = torch.zeros_like(counts)
    for r in range(rank):
        before += allc[r]
    start = torch.cumsum(counts, 0) - counts                      # edges are grouped by ascending latent
    ordinal = torch.arange(lat_idx.numel(), device=lat_idx.device) - start[lat_idx]
    keep = (before[lat_idx] + ordinal) < cap
    return lat_idx[keep], phys_idx[keep]


def local_encoder_edges(strategy: str, phys_local, latent_pos, radius: float, k: int, group=None,
                        cap: int = PYG_MAX_NUM_NEIGHBORS) -> torch.Tensor:
    """[local phys idx; latent idx] edges of this rank, identical (after adding the shard offset) to
    the rows of the unsharded encoder graph whose phys index falls in this rank's range."""
    n_lat = latent_pos.shape[0]
    e_knn = e_rad = None
    if strategy in ("knn", "bidirectional"):
        e_knn = knn_graph(latent_pos, phys_local, k)                                   # [phys, latent]
    if strategy in ("radius", "bidirectional"):
        raw = radius_graph(phys_local, latent_pos, radius, max_num_neighbors=cap)      # [latent, phys]
        li, pi = apply_global_radius_cap(raw[0], raw[1], n_lat, cap, group)
        e_rad = torch.stack([pi, li])
    if strategy == "knn":
        return _tag(e_knn, False)
    if strategy == "radius":
        return _tag(e_rad, True)
    if strategy == "bidirectional":
        return _tag(_coalesced([e_knn, e_rad], phys_local.shape[0], n_lat), False)
    raise ValueError(f"Unknown encoder strategy: {strategy}")


def sharded_encoder_geo_features(phys_local, latent_pos, enc_edges, group=None) -> torch.Tensor:
    """Statistical geometric features [M, 9] of the latent tokens when their neighbours (physical points) are
    sharded: local moment sums -> all-reduce(SUM) -> eigenvalues + z-score, identical on every rank
    (reference geoembed.py:99-182 on the unsharded graph)."""
    csr = ops.csr_of(enc_edges, phys_local.shape[0], latent_pos.shape[0])
    mom = ops.geo_moments(phys_local, latent_pos, csr)
    dist.all_reduce(mom, op=dist.ReduceOp.SUM, group=group)
    return ops.geo_from_moments(mom, normalize=True)


def global_zscore(feat_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """z-score of row-sharded features over ALL rows (geoembed.py:177-180: mean, unbiased std, std < 1e-6 -> 1)."""
    f64 = feat_local.double()
    stats = torch.stack([f64.sum(0), (f64 * f64).sum(0)])
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    mean = stats[0] / n_total
    var = ((stats[1] - n_total * mean * mean) / max(n_total - 1, 1)).clamp(min=0)
    std = var.sqrt().float()
    std = torch.where(std < 1e-6, torch.ones_like(std), std)
    return (feat_local - mean.float()) / std


def sharded_forward(model, batch_local, tokens_pos, n_total: int, group=None, enc_edges: Optional[torch.Tensor] = None,
                    head_parallel: bool = True):
    """GAOT3D forward on this rank's shard of ONE sample (batch of one).  Returns the local rows of
    the output [N_local, C_out].  `model` is a gaot_3d_b200.GAOT3D (single scale, use_gno=True)."""
    enc, dec = model.encoder, model.decoder
    if len(enc.scales) != 1 or not enc.use_gno:
        raise NotImplementedError("sharded path: single scale with the GNO enabled")
    # options whose sharded form needs a collective this path does not have (silently wrong otherwise)
    if enc.gno.use_attn or dec.gno.use_attn:
        raise NotImplementedError("sharded path: use_attn needs a global segment softmax across the ranks")
    if enc.sampling_strategy is not None:
        raise NotImplementedError("sharded path: sampling_strategy must be None")
    if getattr(batch_local, "num_graphs", 1) != 1:
        raise NotImplementedError("sharded path: one example per step (num_graphs == 1)")
    for side in (enc, dec):
        if side.use_geoembed and side.geoembed.method != "statistical":
            raise NotImplementedError("sharded path: geometric embedding method must be 'statistical'")
    from .layers.magno import _apply_node_mlp
    dev = batch_local.pos.device
    lat = tokens_pos.to(dev)
    M = lat.shape[0]
    pos = batch_local.pos
    if enc_edges is None:
        enc_edges = local_encoder_edges(enc.encoder_strategy, pos, lat, enc.gno_radius, enc.k_neighbors, group)
    lifted = _apply_node_mlp(enc.lifting, enc.mlp_type, enc._features(batch_local))
    part = enc.gno(y_pos=pos, x_pos=lat, edge_index=enc_edges, f_y=lifted, reduce="sum")       # partial sums [M,C]
    cnt = torch.bincount(enc_edges[1], minlength=M).to(part.dtype)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    latent = all_reduce_forward(part, group) / cnt.clamp(min=1).unsqueeze(1)
    if enc.use_geoembed:
        # the neighbours of a latent token are spread over the ranks: all-reduce the moment SUMS, then every rank
        # finishes (eigenvalues, z-score over all tokens) identically; coordinates carry no gradient
        geo = enc.geoembed.mlp(sharded_encoder_geo_features(pos, lat, enc_edges, group))
        latent = _apply_node_mlp(enc.recovery, enc.mlp_type, torch.cat([latent, geo], dim=-1))
    # the transformer is replicated; its attention core splits by heads across the ranks when the head counts divide
    # (tblock.set_head_parallel: all-gather of the head outputs forward, of dqkv backward; same kernels, same numbers)
    from . import tblock
    tblock.set_head_parallel(head_parallel, group)
    try:
        rn = model.process(latent.view(1, M, -1))
    finally:
        tblock.set_head_parallel(False)
    rn = all_reduce_backward(rn.reshape(M, -1), group)
    if dec.decoder_strategy == "reverse":            # flip of the *bidirectional* encoder graph (magno.py:263-273)
        bi = enc_edges if enc.encoder_strategy == "bidirectional" else \
            local_encoder_edges("bidirectional", pos, lat, dec.gno_radius, dec.k_neighbors, group)
        dec_edges = _tag(bi.flip(0).contiguous(), True)
    else:
        from .graph import get_neighbor_strategy
        dec_edges = get_neighbor_strategy(dec.decoder_strategy, pos, None, lat, None, dec.gno_radius, dec.k_neighbors, True)
    out = dec.gno(y_pos=lat, x_pos=pos, edge_index=dec_edges, f_y=rn)
    if dec.use_geoembed:
        # queries are sharded, their neighbours (latents) are not: local statistics, global z-score
        raw = dec.geoembed.statistical_features(lat, pos, dec_edges, normalize=False)
        geo = dec.geoembed.mlp(global_zscore(raw, n_total, group))
        out = _apply_node_mlp(dec.recovery, dec.mlp_type, torch.cat([out, geo], dim=-1))
    return _apply_node_mlp(dec.projection, dec.mlp_type, out)


def allreduce_partial_grads(model, group=None):
    """GNO-side parameters see only this rank's points -> SUM their grads.  Everything computed on REPLICATED data
    already holds the total gradient on every rank and is left alone: the transformer (processor, patch_linear) and the
    encoder's geometric-embedding MLP / recovery layer, which run on the all-reduced latent tokens."""
    replicated = ("processor.", "patch_linear.", "encoder.geoembed.", "encoder.recovery.")
    bufs = [p.grad for n, p in model.named_parameters() if p.grad is not None and not n.startswith(replicated)]
    if not bufs:
        return
    flat = torch.cat([b.reshape(-1) for b in bufs])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for b in bufs:
        b.copy_(flat[off: off + b.numel()].view_as(b))
        off += b.numel()
