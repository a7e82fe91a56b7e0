"""Intra-sample sharding of one large mesh across the GPUs of a node (new -- the
reference only has sample-level DDP, src/trainer/stat.py:431-436).

Partitioning: rank r owns the contiguous physical-point index range [r*N/R, (r+1)*N/R) AND the r-th D-slab of the
latent grid (a contiguous token range [r*M/R, (r+1)*M/R) = a contiguous patch range); weights are replicated, no
activation is.
  encoder : local edges (local phys shard x all latents) -> GNO partial SUMS [M,C] with the edge COUNT as one more
            column -> ONE reduce-scatter(SUM) -> mean on the local token slab.   The radius cap (first 32 per latent by
            ascending phys index, globally) is restored with one all-gather of per-latent counts: contiguous index
            ranges make "ascending index" compose as an exclusive prefix over ranks.
  processor: sequence-parallel (tblock "sp" mode): every GEMM / norm on the S/R local tokens, two all-to-alls per block
            and direction around the head-sharded attention core, recorded into the processor's CUDA graphs (tgraph.py).
  decoder : all-gather of the processed latent slabs -> queries (phys) sharded, sources (latents) complete -> no further
            forward communication.
  backward: the all-gather's backward is a reduce-scatter of the partial d latent, the reduce-scatter's an all-gather;
            every parameter gradient is a partial sum (over local points or local tokens) -> one flat all-reduce
            (allreduce_partial_grads).
`mode="hp"` keeps round 1's layout (replicated transformer, attention core split by heads) for comparison.
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


class _ReduceScatterFwd(torch.autograd.Function):
    """y_r = (sum over ranks of x)[slab r]  (x [M, W] partial sums, y [M/R, W]).  Backward: all-gather of the slab
    gradients -- every rank's partial sums fed every slab."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        R = dist.get_world_size(group)
        x = x.contiguous()
        y = torch.empty((x.shape[0] // R,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        _reduce_scatter(y, x, group)
        return y

    @staticmethod
    def backward(ctx, g):
        R = dist.get_world_size(ctx.group)
        g = g.contiguous()
        out = torch.empty((g.shape[0] * R,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        _all_gather(out, g, ctx.group)
        return out, None


class _AllGatherFwd(torch.autograd.Function):
    """y = concat over ranks of x_r (row slabs).  Backward: reduce-scatter(SUM) of the per-rank partial gradients."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        R = dist.get_world_size(group)
        x = x.contiguous()
        out = torch.empty((x.shape[0] * R,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        _all_gather(out, x, group)
        return out

    @staticmethod
    def backward(ctx, g):
        R = dist.get_world_size(ctx.group)
        g = g.contiguous()
        y = torch.empty((g.shape[0] // R,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        _reduce_scatter(y, g, ctx.group)
        return y, None


def _reduce_scatter(out, inp, group):
    if inp.is_cuda:
        from . import p2p
        res = p2p.reduce_scatter(inp, group)                 # peer-memory pull-reduce (deterministic order) when available
        if res is not None:
            out.copy_(res)
            return
        dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)
    else:       # gloo (CPU tests) has no reduce_scatter: all-reduce a copy and keep the own slab
        tmp = inp.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        n = out.shape[0]
        out.copy_(tmp[dist.get_rank(group) * n:(dist.get_rank(group) + 1) * n])


def _all_gather(out, inp, group):
    if inp.is_cuda:
        from . import p2p
        res = p2p.all_gather(inp, group)
        if res is not None:
            out.copy_(res.view(out.shape))
            return
        dist.all_gather_into_tensor(out, inp, group=group)
    else:
        parts = [torch.empty_like(inp) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, inp, group=group)
        out.copy_(torch.cat(parts, dim=0))


def reduce_scatter_forward(x, group=None):
    return _ReduceScatterFwd.apply(x, group)


def all_gather_forward(x, group=None):
    return _AllGatherFwd.apply(x, group)


def all_reduce_forward(x, group=None):
    return _AllReduceFwd.apply(x, group)


def all_reduce_backward(x, group=None):
    return _AllReduceBwd.apply(x, group)


def apply_global_radius_cap(lat_idx: torch.Tensor, phys_idx: torch.Tensor, num_latent: int, cap: int, group=None):
    """Local radius edges (grouped by latent, ascending local phys index, locally capped at `cap`) ->
    the subset that survives the GLOBAL cap: edge ordinal (edges of lower ranks + local ordinal) < cap."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = _count(lat_idx, num_latent)
    allc = torch.empty(world, num_latent, dtype=counts.dtype, device=counts.device)
    _all_gather(allc.view(-1), counts, group)
    before = allc[:rank].sum(dim=0)                               # edges of the lower ranks, per latent
    start = torch.cumsum(counts, 0) - counts                      # edges are grouped by ascending latent
    ordinal = torch.arange(lat_idx.numel(), device=lat_idx.device) - start[lat_idx]
    keep = (before[lat_idx] + ordinal) < cap
    return lat_idx[keep], phys_idx[keep]


def _count(idx: torch.Tensor, n: int) -> torch.Tensor:
    """bincount without the host synchronisation torch.bincount makes (it reads the maximum back)."""
    return torch.zeros(n, dtype=torch.long, device=idx.device).scatter_add_(0, idx, torch.ones_like(idx))


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


_DEFAULT = {"mode": "sp"}


def set_default_mode(mode: str) -> None:
    """Process-wide default of `sharded_forward(mode=None)` / `allreduce_partial_grads(mode=None)`: "sp" or "hp"."""
    if mode not in ("sp", "hp"):
        raise ValueError("shard mode must be 'sp' or 'hp'")
    _DEFAULT["mode"] = mode


def sharded_forward(model, batch_local, tokens_pos, n_total: int, group=None, enc_edges: Optional[torch.Tensor] = None,
                    head_parallel: bool = True, mode: Optional[str] = None):
    from .graph import sample_scope
    with sample_scope():
        return _sharded_forward(model, batch_local, tokens_pos, n_total, group, enc_edges, head_parallel, mode)


def _sharded_forward(model, batch_local, tokens_pos, n_total: int, group=None, enc_edges: Optional[torch.Tensor] = None,
                     head_parallel: bool = True, mode: Optional[str] = None):
    """GAOT3D forward on this rank's shard of ONE sample (batch of one).  Returns the local rows of
    the output [N_local, C_out].  `model` is a gaot_3d_b200.GAOT3D (single scale, use_gno=True).
    mode "sp": token-sharded processor (nothing replicated); "hp": round 1's replicated processor with a head-parallel
    attention core (`head_parallel=False`: fully replicated)."""
    mode = mode or _DEFAULT["mode"]
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
    from . import tblock
    dev = batch_local.pos.device
    lat = tokens_pos.to(dev)
    M = lat.shape[0]
    pos = batch_local.pos
    R, rank = dist.get_world_size(group), dist.get_rank(group)
    heads = model.processor.encoder_layers[0].attn if len(model.processor.encoder_layers) else model.processor.middle_layer.attn
    sp = mode == "sp" and R > 1
    if sp and ((model.D // model.patch_size) % R or heads.num_heads % R or heads.num_kv_heads % R):
        raise NotImplementedError("sequence-parallel processor: patch planes and head counts must be divisible by the ranks")
    if enc_edges is None:
        enc_edges = local_encoder_edges(enc.encoder_strategy, pos, lat, enc.gno_radius, enc.k_neighbors, group)
    lifted = _apply_node_mlp(enc.lifting, enc.mlp_type, enc._features(batch_local))
    part = enc.gno(y_pos=pos, x_pos=lat, edge_index=enc_edges, f_y=lifted, reduce="sum")       # partial sums [M,C]
    if pos.is_cuda:        # the CSR the GNO kernel just built (cached on the edge tensor) already holds the per-token counts
        rp = ops.csr_of(enc_edges, pos.shape[0], M).rowptr
        cnt = (rp[1:] - rp[:-1]).to(part.dtype)
    else:
        cnt = _count(enc_edges[1], M).to(part.dtype)
    fused = torch.cat([part, cnt.unsqueeze(1)], dim=1)                # sums and counts travel in ONE collective
    lo, hi = (rank * (M // R), (rank + 1) * (M // R)) if sp else (0, M)
    tot = reduce_scatter_forward(fused, group) if sp else all_reduce_forward(fused, group)
    latent = tot[:, :-1] / tot[:, -1:].detach().clamp(min=1)
    if enc.use_geoembed:
        # the neighbours of a latent token are spread over the ranks: all-reduce the moment SUMS, then every rank
        # finishes (eigenvalues, z-score over all tokens) identically; coordinates carry no gradient
        feats = sharded_encoder_geo_features(pos, lat, enc_edges, group)[lo:hi]
        latent = _apply_node_mlp(enc.recovery, enc.mlp_type, torch.cat([latent, enc.geoembed.mlp(feats)], dim=-1))
    tblock.set_head_parallel(sp or head_parallel, group, "sp" if sp else "hp")
    try:
        rn = model.process(latent.reshape(1, hi - lo, -1), slab=(rank, R) if sp else (0, 1))
    finally:
        tblock.set_head_parallel(False)
    rn = rn.reshape(hi - lo, -1)
    rn = all_gather_forward(rn, group) if sp else all_reduce_backward(rn, group)
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


def allreduce_partial_grads(model, group=None, mode: Optional[str] = None):
    """SUM the parameter gradients that are partial on every rank, in ONE flat all-reduce.  "sp": all of them (GNO-side
    parameters see the local points, everything else the local token slab).  "hp": only the GNO side; what ran on
    REPLICATED data (processor, patch_linear, encoder geoembed MLP / recovery) already holds the total."""
    mode = mode or _DEFAULT["mode"]
    if mode == "sp" and dist.get_world_size(group) > 1:
        heads = model.processor.encoder_layers[0].attn if len(model.processor.encoder_layers) else model.processor.middle_layer.attn
        R = dist.get_world_size(group)
        if (model.D // model.patch_size) % R or heads.num_heads % R or heads.num_kv_heads % R:
            mode = "hp"
    replicated = ("processor.", "patch_linear.", "encoder.geoembed.", "encoder.recovery.") if mode == "hp" else ()
    bufs = [p.grad for n, p in model.named_parameters() if p.grad is not None and not (replicated and n.startswith(replicated))]
    if not bufs:
        return
    flat = torch.cat([b.reshape(-1) for b in bufs])
    from . import p2p
    if not (flat.dtype == torch.float32 and p2p.all_reduce_(flat, group)):       # two-shot over peer memory, else NCCL
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    torch._foreach_copy_(bufs, [flat[o:o + b.numel()].view_as(b) for o, b in zip(_offsets(bufs), bufs)])


def _offsets(bufs):
    off = 0
    for b in bufs:
        yield off
        off += b.numel()
