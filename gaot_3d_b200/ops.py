"""Tensor-level host API over the C ABI (include/gaot_b200.h).

PyTorch is plumbing only: device memory (torch.empty), the current CUDA stream and autograd
bookkeeping.  Every function here launches hand-written sm_100a kernels from
libgaot_b200.so through ctypes; CPU tensors are rejected (there is no fallback path).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import MlpDesc, check

_PRECISION = {"gno": 0}     # 0 = fp32 CUDA cores, 1 = bf16 tcgen05


def set_gno_precision(name: str) -> None:
    if name not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _PRECISION["gno"] = 0 if name == "fp32" else 1


def get_gno_precision() -> str:
    return "fp32" if _PRECISION["gno"] == 0 else "bf16"


_TRANSFORMER = {"precision": "bf16"}


def set_transformer_precision(name: str) -> None:
    """'bf16' (default): the latent transformer on this library's tcgen05 kernels -- bf16 operands, fp32 accumulation, the
    north star's rtol 2e-2 tier.  'fp32': the strict tier -- the same modules through the reference's own library calls
    (fp32 cuBLAS GEMMs with TF32 off, fp32 F.scaled_dot_product_attention, reference attn.py:100-128), so that a whole-model
    rtol 1e-5 comparison exists; graph build, GNO, geometric embedding stay on this library's fp32 kernels."""
    if name not in ("bf16", "fp32"):
        raise ValueError("transformer precision must be 'bf16' or 'fp32'")
    _TRANSFORMER["precision"] = name


def transformer_precision() -> str:
    return _TRANSFORMER["precision"]


_NODE_MLP = {"mode": "torch"}


def set_node_mlp_mode(mode: str) -> None:
    """'torch' (default: torch GEMMs in whatever precision torch's flags say), 'tf32' (torch GEMMs with the TF32 flags
    on), or 'fused' (TF32 flags on, and two-layer GELU MLPs of the projection-head shape 32 -> 256 -> <= 8 as one fused
    tensor-core kernel with f16/bf16 operands -- the rtol 2e-2 tier)."""
    if mode not in ("torch", "tf32", "fused"):
        raise ValueError("node MLP mode must be 'torch', 'tf32' or 'fused'")
    _NODE_MLP["mode"] = mode
    if mode != "torch":
        set_node_mlp_tf32(True)


def node_mlp_mode() -> str:
    return _NODE_MLP["mode"]


def set_node_mlp_tf32(on: bool = True) -> None:
    """Node-level MLPs (lifting, projection, recovery, geo-embedding MLP) are torch fp32 GEMMs on [N_points, C]
    tensors.  The reference's default mlp_type='channel' routes them through Conv1d, for which cuDNN allows TF32
    (SURVEY.md appendix C.10); this switches the same tensor-core mode on for the Linear form as well.  It is
    torch's process-wide flag pair -- the caller (bench.py, a trainer) decides."""
    torch.backends.cuda.matmul.allow_tf32 = bool(on)
    torch.backends.cudnn.allow_tf32 = bool(on)


def _lib_():
    return _lib.load()


# ---- optional per-op device timing (bench.py): CUDA events on the launching stream ----
_TIMING = {"on": False, "events": []}


def enable_timing(on: bool = True) -> None:
    _TIMING["on"] = on
    _TIMING["events"] = []


class _timed:
    def __init__(self, name, dev):
        self.name, self.dev = name, dev

    def __enter__(self):
        if _TIMING["on"]:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.dev))
        return self

    def __exit__(self, *a):
        if _TIMING["on"]:
            self.e.record(torch.cuda.current_stream(self.dev))
            _TIMING["events"].append((self.name, self.s, self.e))
        return False


def timing_summary() -> dict:
    """name -> (calls, total ms); call after torch.cuda.synchronize()."""
    out = {}
    for name, s, e in _TIMING["events"]:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + s.elapsed_time(e))
    return out


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gaot_3d_b200: the B200 hot path has no CPU fallback; got a CPU tensor "
                               "(build graphs / run the model on a CUDA device)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ws(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def _pos3(p: torch.Tensor) -> torch.Tensor:
    """float32 contiguous [n,3]; 2-D inputs are zero padded (dz*dz = 0 keeps distances exact)."""
    if p.dim() != 2 or p.shape[1] not in (2, 3):
        raise ValueError("positions must be [N, 2] or [N, 3]")
    p = p.detach().to(torch.float32)
    if p.shape[1] == 2:
        p = torch.cat([p, p.new_zeros(p.shape[0], 1)], dim=1)
    return p.contiguous()


# ----------------------------------------------------------------------------- graph build
def radius(x: torch.Tensor, y: torch.Tensor, r: float, max_num_neighbors: int = 32):
    """For each y all x within r (strict), first `max_num_neighbors` by ascending x index.
    Returns (row_y, col_x) int64, grouped by ascending y -- torch_cluster.radius semantics
    (reference magno.py:193-200, :253-260)."""
    _need_cuda(x, y)
    x, y = _pos3(x), _pos3(y)
    nx, ny, dev = x.shape[0], y.shape[0], x.device
    lib = _lib_()
    empty = torch.empty(0, dtype=torch.long, device=dev)
    if nx == 0 or ny == 0:
        return empty, empty.clone()
    wsb = lib.gaot_radius_workspace_bytes(nx, ny)
    ws = _ws(wsb, dev)
    rowptr = torch.empty(ny + 1, dtype=torch.int32, device=dev)
    E = ctypes.c_int64(0)
    with torch.cuda.device(dev), _timed("radius", dev):
        check(lib.gaot_radius_count(_p(x), nx, _p(y), ny, float(r), int(max_num_neighbors), _p(ws), wsb,
                                    _p(rowptr), ctypes.byref(E), _stream(dev)), "radius_count")
        out_y = torch.empty(E.value, dtype=torch.long, device=dev)
        out_x = torch.empty(E.value, dtype=torch.long, device=dev)
        if E.value:
            check(lib.gaot_radius_emit(_p(x), nx, _p(y), ny, float(r), int(max_num_neighbors), _p(ws), wsb,
                                       _p(rowptr), _p(out_y), _p(out_x), _stream(dev)), "radius_emit")
    return out_y, out_x


def knn(x: torch.Tensor, y: torch.Tensor, k: int):
    """For each y its k nearest x, ascending (distance, index).  Returns (row_y, col_x) int64 of
    length ny*min(k, nx), grouped by y -- torch_cluster.knn semantics (magno.py:183-189, :242-248)."""
    _need_cuda(x, y)
    x, y = _pos3(x), _pos3(y)
    nx, ny, dev = x.shape[0], y.shape[0], x.device
    lib = _lib_()
    if nx == 0 or ny == 0:
        e = torch.empty(0, dtype=torch.long, device=dev)
        return e, e.clone()
    k = min(int(k), nx)
    wsb = lib.gaot_knn_workspace_bytes(nx, ny)
    ws = _ws(wsb, dev)
    out_y = torch.empty(ny * k, dtype=torch.long, device=dev)
    out_x = torch.empty(ny * k, dtype=torch.long, device=dev)
    with torch.cuda.device(dev), _timed("knn", dev):
        check(lib.gaot_knn(_p(x), nx, _p(y), ny, k, _p(ws), wsb, _p(out_y), _p(out_x), _stream(dev)), "knn")
    return out_y, out_x


def coalesce(row0: torch.Tensor, row1: torch.Tensor, max_row0: int, max_row1: int):
    """Lexicographic sort + unique of an edge list (torch_geometric.utils.coalesce, magno.py:220,:293)."""
    _need_cuda(row0, row1)
    row0, row1 = row0.contiguous(), row1.contiguous()
    E, dev = row0.numel(), row0.device
    lib = _lib_()
    if E == 0:
        return row0.clone(), row1.clone()
    wsb = lib.gaot_coalesce_workspace_bytes(E)
    ws = _ws(wsb, dev)
    o0, o1 = torch.empty_like(row0), torch.empty_like(row1)
    e_dev = torch.empty(1, dtype=torch.long, device=dev)
    e_host = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        check(lib.gaot_coalesce(_p(row0), _p(row1), E, int(max_row0), int(max_row1), _p(ws), wsb, _p(o0), _p(o1),
                                _p(e_dev), ctypes.byref(e_host), _stream(dev)), "coalesce")
    return o0[: e_host.value], o1[: e_host.value]


def edge_mask(row0: torch.Tensor, row1: torch.Tensor, p_drop: float, seed: int, offset: int = 0):
    """Philox edge dropout + order-preserving compaction (torch_geometric dropout_edge, magno.py:367)."""
    _need_cuda(row0, row1)
    row0, row1 = row0.contiguous(), row1.contiguous()
    E, dev = row0.numel(), row0.device
    lib = _lib_()
    if E == 0:
        return row0.clone(), row1.clone()
    wsb = lib.gaot_edge_mask_workspace_bytes(E)
    ws = _ws(wsb, dev)
    o0, o1 = torch.empty_like(row0), torch.empty_like(row1)
    e_dev = torch.empty(1, dtype=torch.long, device=dev)
    e_host = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        check(lib.gaot_edge_mask(_p(row0), _p(row1), E, float(p_drop), int(seed) & (2 ** 64 - 1), int(offset),
                                 _p(ws), wsb, _p(o0), _p(o1), _p(e_dev), ctypes.byref(e_host), _stream(dev)),
              "edge_mask")
    return o0[: e_host.value], o1[: e_host.value]


@dataclass
class Csr:
    """Query-major CSR side-band of an edge list (int32)."""
    rowptr: torch.Tensor    # [nq + 1]
    src: torch.Tensor       # [E] source index per CSR slot
    qry: torch.Tensor       # [E] query index per CSR slot
    perm: torch.Tensor      # [E] original edge id per CSR slot
    n_src: int
    nq: int

    @property
    def E(self) -> int:
        return int(self.src.numel())


def build_csr(src: torch.Tensor, qry: torch.Tensor, n_src: int, nq: int, query_sorted: bool = False,
              validate: bool = True) -> Csr:
    """`validate` (default): indices are range-checked on the device (and the `query_sorted` claim verified) before any
    kernel gathers by them; a bad edge list raises ValueError.  Graphs this package built itself skip the check."""
    _need_cuda(src, qry)
    src = src.to(torch.long).contiguous()
    qry = qry.to(torch.long).contiguous()
    E, dev = src.numel(), src.device
    lib = _lib_()
    rowptr = torch.empty(nq + 1, dtype=torch.int32, device=dev)
    cs = torch.empty(E, dtype=torch.int32, device=dev)
    cq = torch.empty(E, dtype=torch.int32, device=dev)
    pm = torch.empty(E, dtype=torch.int32, device=dev)
    wsb = lib.gaot_csr_workspace_bytes(E, nq)
    ws = _ws(wsb, dev)
    with torch.cuda.device(dev):
        check(lib.gaot_csr_from_edges(_p(src), _p(qry), E, int(n_src), int(nq), (1 if query_sorted else 0) | (2 if validate else 0),
                                      _p(ws), wsb, _p(rowptr), _p(cs), _p(cq), _p(pm), _stream(dev)), "csr_from_edges")
    return Csr(rowptr, cs, cq, pm, int(n_src), int(nq))


def csr_of(edge_index: torch.Tensor, n_src: int, nq: int) -> Csr:
    """CSR of an edge_index [2,E] (row 0 = source, row 1 = query); cached on the tensor object so the
    GNO and the geometric embedding of one scale share it."""
    cache = getattr(edge_index, "_gaot_csr", None)
    if cache is not None and cache.n_src == n_src and cache.nq == nq and cache.E == edge_index.shape[1]:
        return cache
    # `_gaot_trusted` is set by graph._tag on edge lists produced by this package's own search / coalesce / mask kernels;
    # anything else (precomputed edges from dataset files, reference magno.py:506-516) is validated once here
    csr = build_csr(edge_index[0], edge_index[1], n_src, nq, bool(getattr(edge_index, "_gaot_query_sorted", False)),
                    validate=not bool(getattr(edge_index, "_gaot_trusted", False)))
    try:
        edge_index._gaot_csr = csr
    except Exception:
        pass
    return csr


# ----------------------------------------------------------------------------- GNO
_TRANSFORMS = {"linear": 0, "nonlinear": 1, "nonlinear_kernelonly": 2}


def _mlp_desc(dims: Sequence[int]) -> MlpDesc:
    d = MlpDesc()
    d.n_layers = len(dims) - 1
    if d.n_layers < 1 or d.n_layers > 6:
        raise NotImplementedError("gno: the fused kernel supports 1..6 MLP layers")
    for i, v in enumerate(dims):
        d.dims[i] = int(v)
    return d


def _flat_params(weights, biases) -> torch.Tensor:
    parts = []
    for w, b in zip(weights, biases):
        parts.append(w.reshape(-1))
        parts.append(b.reshape(-1))
    return torch.cat(parts).to(torch.float32).contiguous()


class _GnoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pos, x_pos, f_y, csr: Csr, dims, transform, reduce, precision, edge_w, *wb):
        nl = len(dims) - 1
        weights, biases = wb[:nl], wb[nl:]
        params = _flat_params(weights, biases)
        dev = x_pos.device
        lib = _lib_()
        desc = _mlp_desc(dims)
        nq, n_src, E = csr.nq, csr.n_src, csr.E
        c_f = 0 if f_y is None else f_y.shape[1]
        fy = None if f_y is None else f_y.detach().to(torch.float32).contiguous()
        out = torch.empty(nq, dims[-1], dtype=torch.float32, device=dev)
        wsb = lib.gaot_gno_workspace_bytes(E, nq, ctypes.byref(desc))
        ws = _ws(wsb, dev)
        ew = None if edge_w is None else edge_w.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev), _timed("gno_fwd", dev):
            check(lib.gaot_gno_forward_weighted(_p(y_pos), n_src, _p(x_pos), nq, _p(fy), c_f, _p(csr.rowptr), _p(csr.src),
                                                _p(csr.qry), E, ctypes.byref(desc), _p(params), transform, reduce, precision,
                                                _p(ew), _p(ws), wsb, _p(out), _stream(dev)), "gno_forward")
        ctx.save_for_backward(y_pos, x_pos, fy, params, ew)
        ctx.need_ew = edge_w is not None and edge_w.requires_grad
        ctx.csr, ctx.dims, ctx.transform, ctx.reduce, ctx.precision = csr, dims, transform, reduce, precision
        ctx.need_f = f_y is not None and f_y.requires_grad
        ctx.shapes = [(tuple(w.shape), tuple(b.shape)) for w, b in zip(weights, biases)]
        return out

    @staticmethod
    def backward(ctx, d_out):
        y_pos, x_pos, fy, params, ew = ctx.saved_tensors
        csr, dims = ctx.csr, ctx.dims
        dev = x_pos.device
        lib = _lib_()
        desc = _mlp_desc(dims)
        d_out = d_out.to(torch.float32).contiguous()
        d_params = torch.empty_like(params)
        c_f = 0 if fy is None else fy.shape[1]
        d_f = torch.empty_like(fy) if ctx.need_f else None
        wsb = lib.gaot_gno_workspace_bytes(csr.E, csr.nq, ctypes.byref(desc))
        ws = _ws(wsb, dev)
        d_ew = torch.empty_like(ew) if (ew is not None and ctx.need_ew) else None
        with torch.cuda.device(dev), _timed("gno_bwd", dev):
            check(lib.gaot_gno_backward_weighted(_p(y_pos), csr.n_src, _p(x_pos), csr.nq, _p(fy), c_f, _p(csr.rowptr),
                                                 _p(csr.src), _p(csr.qry), csr.E, ctypes.byref(desc), _p(params),
                                                 ctx.transform, ctx.reduce, ctx.precision, _p(ew), _p(d_out), _p(ws), wsb,
                                                 _p(d_params), _p(d_f), _p(d_ew), _stream(dev)), "gno_backward")
        gw, gb, off = [], [], 0
        for (ws_, bs_) in ctx.shapes:
            nw = ws_[0] * ws_[1]
            gw.append(d_params[off: off + nw].view(ws_)); off += nw
            gb.append(d_params[off: off + bs_[0]].view(bs_)); off += bs_[0]
        return (None, None, d_f, None, None, None, None, None, d_ew, *gw, *gb)


def gno(y_pos, x_pos, f_y, csr: Csr, weights, biases, transform_type: str = "linear",
        reduce: str = "mean", precision: Optional[str] = None, edge_w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused gather -> kernel MLP -> (* f_y) -> segmented mean (reference integral_transform.py:114-171).
    `edge_w` [E] (CSR order, see Csr.src / Csr.qry): per-edge attention weights multiplied in before a SUM reduction
    (use_attn, :161-165); differentiable."""
    if edge_w is not None:
        reduce = "sum"
    _need_cuda(y_pos, x_pos, f_y)
    if f_y is None:
        t = 3
    else:
        if transform_type not in _TRANSFORMS:
            raise ValueError(f"unknown transform_type {transform_type}")
        t = _TRANSFORMS[transform_type]
    if y_pos.shape[1] == 2:
        # 2-D coordinates (MAGNOConfig.gno_coord_dim = 2, the dataclass default): the kernels gather zero-padded 3-D points, so the
        # first layer [h, 2 + 2 (+ C)] gets zero columns where the padded z of y and of x sit -- same products, and autograd
        # hands the gradient of the real columns back through the concatenation
        w0 = weights[0]
        z = w0.new_zeros(w0.shape[0], 1)
        weights = [torch.cat([w0[:, :2], z, w0[:, 2:4], z, w0[:, 4:]], dim=1), *weights[1:]]
    dims = [int(weights[0].shape[1])] + [int(w.shape[0]) for w in weights]
    prec = _PRECISION["gno"] if precision is None else (0 if precision == "fp32" else 1)
    return _GnoFn.apply(_pos3(y_pos), _pos3(x_pos), f_y, csr, tuple(dims), t, 0 if reduce == "mean" else 1, prec,
                        edge_w, *weights, *biases)


# ----------------------------------------------------------------------------- geometric embedding
def geo_stats(source_pos, query_pos, csr: Csr, normalize: bool = True) -> torch.Tensor:
    """[nq, 9] statistical features (reference geoembed.py:99-182); z-scored over the queries when
    `normalize` (the sharded encoder normalises after its all-reduce instead)."""
    _need_cuda(source_pos, query_pos)
    sp, qp = _pos3(source_pos), _pos3(query_pos)
    dev = qp.device
    lib = _lib_()
    feat = torch.empty(csr.nq, 9, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.gaot_geo_stats(_p(sp), csr.n_src, _p(qp), csr.nq, _p(csr.rowptr), _p(csr.src), _p(feat),
                                 _stream(dev)), "geo_stats")
        if normalize:
            zscore_(feat)
    return feat


def geo_moments(source_pos, query_pos, csr: Csr) -> torch.Tensor:
    """[nq, 12] per-query moment SUMS of (y - x) over this rank's sources (count, sum d, sum d^2, sum delta (3),
    sum delta delta^T (6)); partials of disjoint source shards add up (sharded encoder)."""
    _need_cuda(source_pos, query_pos)
    sp, qp = _pos3(source_pos), _pos3(query_pos)
    dev = qp.device
    lib = _lib_()
    mom = torch.empty(csr.nq, 12, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.gaot_geo_moments(_p(sp), csr.n_src, _p(qp), csr.nq, _p(csr.rowptr), _p(csr.src), _p(mom),
                                   _stream(dev)), "geo_moments")
    return mom


def geo_from_moments(moments: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """moment sums [nq, 12] -> the [nq, 9] statistical features of geo_stats (same finalisation code)."""
    _need_cuda(moments)
    mom = moments.to(torch.float32).contiguous()
    dev = mom.device
    lib = _lib_()
    feat = torch.empty(mom.shape[0], 9, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.gaot_geo_from_moments(_p(mom), mom.shape[0], _p(feat), _stream(dev)), "geo_from_moments")
        if normalize:
            zscore_(feat)
    return feat


class _PointNetPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sp, qp, csr, pooling, w1, b1, w2, b2):
        lib = _lib_()
        dev = qp.device
        params = torch.cat([w1.reshape(-1), b1.reshape(-1), w2.reshape(-1), b2.reshape(-1)]).detach().to(torch.float32).contiguous()
        pooled = torch.empty(csr.nq, 32, dtype=torch.float32, device=dev)
        argmax = torch.empty(csr.nq, 32, dtype=torch.int32, device=dev) if pooling == 0 else None
        with torch.cuda.device(dev):
            check(lib.gaot_pointnet_forward(_p(sp), csr.n_src, _p(qp), csr.nq, _p(csr.rowptr), _p(csr.src), _p(params), pooling,
                                            _p(pooled), _p(argmax), _stream(dev)), "pointnet_forward")
        ctx.save_for_backward(sp, qp, params, argmax)
        ctx.csr, ctx.pooling = csr, pooling
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        sp, qp, params, argmax = ctx.saved_tensors
        csr = ctx.csr
        lib = _lib_()
        dev = qp.device
        d_pooled = d_pooled.to(torch.float32).contiguous()
        dpar = torch.empty_like(params)
        wsb = lib.gaot_pointnet_workspace_bytes()
        ws = _ws(wsb, dev)
        with torch.cuda.device(dev):
            check(lib.gaot_pointnet_backward(_p(sp), csr.n_src, _p(qp), csr.nq, _p(csr.rowptr), _p(csr.src), _p(params), ctx.pooling,
                                             _p(d_pooled), _p(argmax), _p(ws), wsb, _p(dpar), _stream(dev)), "pointnet_backward")
        return None, None, None, None, dpar[:96].view(32, 3), dpar[96:128], dpar[128:1152].view(32, 32), dpar[1152:]


def pointnet_pool(source_pos, query_pos, csr: Csr, w1, b1, w2, b2, pooling: str = "max") -> torch.Tensor:
    """[nq, 32] pooled PointNet features of the centred neighbour coordinates (reference geoembed.py:184-213):
    relu(W2 relu(W1 (y - x) + b1) + b2) per edge, max or mean over each query's edges, 0 for empty queries."""
    _need_cuda(source_pos, query_pos, w1)
    if tuple(w1.shape) != (32, 3) or tuple(w2.shape) != (32, 32):
        raise NotImplementedError("pointnet_pool: the fused kernel is built for the reference's 3 -> 32 -> 32 MLP")
    return _PointNetPoolFn.apply(_pos3(source_pos), _pos3(query_pos), csr, 0 if pooling == "max" else 1, w1, b1, w2, b2)


def zscore_(feat: torch.Tensor) -> torch.Tensor:
    _need_cuda(feat)
    lib = _lib_()
    dev = feat.device
    wsb = lib.gaot_geo_zscore_workspace_bytes(feat.shape[0])
    ws = _ws(wsb, dev)
    with torch.cuda.device(dev):
        check(lib.gaot_geo_zscore(_p(feat), feat.shape[0], feat.shape[1], _p(ws), wsb, _stream(dev)), "geo_zscore")
    return feat


# ----------------------------------------------------------------------------- fused two-layer node MLP
def node_mlp2_supported(c_in: int, hidden: int, c_out: int) -> bool:
    return bool(_lib_().gaot_node_mlp2_supported(int(c_in), int(hidden), int(c_out)))


class _NodeMlp2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        lib = _lib_()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        xd, w1d, b1d, w2d, b2d = f32(x), f32(w1), f32(b1), f32(w2), f32(b2)
        n, c_in = xd.shape
        hidden, c_out = w1d.shape[0], w2d.shape[0]
        dev = xd.device
        y = torch.empty(n, c_out, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.gaot_node_mlp2_forward(_p(xd), n, c_in, hidden, c_out, _p(w1d), _p(b1d), _p(w2d), _p(b2d), _p(y),
                                             _stream(dev)), "node_mlp2_forward")
        ctx.save_for_backward(xd, w1d, b1d, w2d)
        return y

    @staticmethod
    def backward(ctx, dy):
        xd, w1d, b1d, w2d = ctx.saved_tensors
        lib = _lib_()
        n, c_in = xd.shape
        hidden, c_out = w1d.shape[0], w2d.shape[0]
        dev = xd.device
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(xd) if ctx.needs_input_grad[0] else None
        dpar = torch.empty(hidden * c_in + hidden + c_out * hidden + c_out, dtype=torch.float32, device=dev)
        wsb = lib.gaot_node_mlp2_workspace_bytes(c_in, hidden, c_out)
        ws = _ws(wsb, dev)
        with torch.cuda.device(dev):
            check(lib.gaot_node_mlp2_backward(_p(xd), _p(dy), n, c_in, hidden, c_out, _p(w1d), _p(b1d), _p(w2d), _p(ws), wsb,
                                              _p(dx), _p(dpar), _stream(dev)), "node_mlp2_backward")
        o1, o2, o3 = hidden * c_in, hidden * c_in + hidden, hidden * c_in + hidden + c_out * hidden
        return dx, dpar[:o1].view(hidden, c_in), dpar[o1:o2], dpar[o2:o3].view(c_out, hidden), dpar[o3:]


def node_mlp2(x, w1, b1, w2, b2):
    """y = W2 gelu(W1 x + b1) + b2 on [N, 32] rows as one fused tensor-core kernel (forward) and one more (backward):
    the projection head of the decoder (reference magno.py:640-644, :796-797).  Gradients flow to x and all four
    parameters; weight shapes [hidden, c_in] / [c_out, hidden] (a kernel-size-1 Conv1d weight reshaped is the same)."""
    _need_cuda(x, w1, b1, w2, b2)
    w1_, w2_ = w1.reshape(w1.shape[0], -1), w2.reshape(w2.shape[0], -1)
    y = _NodeMlp2Fn.apply(x, w1_, b1, w2_, b2)
    return y


# ----------------------------------------------------------------------------- one-layer node MLP, tiny input width
def node_linear_supported(k_in: int, c_out: int) -> bool:
    return bool(_lib_().gaot_node_linear_supported(int(k_in), int(c_out)))


class _NodeLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        lib = _lib_()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        xd, wd = f32(x), f32(w)
        bd = None if b is None else f32(b)
        n, k_in = xd.shape
        c_out = wd.shape[0]
        dev = xd.device
        y = torch.empty(n, c_out, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.gaot_node_linear_forward(_p(xd), n, k_in, c_out, _p(wd), _p(bd), _p(y), _stream(dev)), "node_linear_forward")
        ctx.save_for_backward(xd, wd)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xd, wd = ctx.saved_tensors
        lib = _lib_()
        n, k_in = xd.shape
        c_out = wd.shape[0]
        dev = xd.device
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(xd) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(wd)
        db = torch.empty(c_out, dtype=torch.float32, device=dev) if ctx.has_bias else None
        wsb = lib.gaot_node_linear_workspace_bytes(k_in, c_out)
        ws = _ws(wsb, dev)
        with torch.cuda.device(dev):
            check(lib.gaot_node_linear_backward(_p(xd), _p(dy), n, k_in, c_out, _p(wd), _p(ws), wsb, _p(dx), _p(dw), _p(db),
                                                _stream(dev)), "node_linear_backward")
        return dx, dw, db


def node_linear(x, w, b=None):
    """y = x W^T + b for [N, k_in <= 16] rows in fp32 as one streaming kernel each way: the encoder's lifting layer (reference
    magno.py:421-424, :540-545).  Weight [c_out, k_in] (a kernel-size-1 Conv1d weight reshaped is the same)."""
    _need_cuda(x, w, b)
    return _NodeLinearFn.apply(x, w.reshape(w.shape[0], -1), b)


# ----------------------------------------------------------------------------- dense layers (nn.Linear)
_DT = {torch.float32: 0, torch.bfloat16: 1}


def linear_supported(in_features: int, out_features: int) -> bool:
    """Shapes the tcgen05 GEMM takes (16-byte aligned rows for fp32 and bf16 operands)."""
    return in_features % 8 == 0 and out_features % 8 == 0 and in_features >= 16 and out_features >= 16


def _rows(t: torch.Tensor) -> torch.Tensor:
    t = t.reshape(-1, t.shape[-1])
    if t.dtype not in _DT:
        t = t.to(torch.float32)
    return t if t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0 else t.contiguous()


def _lin_ws(lib, M, N, K, dev):
    wsb = lib.gaot_linear_workspace_bytes(M, N, K)
    return _ws(wsb, dev), wsb


def _linear_fwd_raw(x, x2, w, bias, residual, out_dtype=torch.float32):
    lib = _lib_()
    M, K1 = x.shape
    K = K1 + (0 if x2 is None else x2.shape[1])
    N = w.shape[0]
    dev = x.device
    y = torch.empty(M, N, dtype=out_dtype, device=dev)
    if M == 0:
        return y
    ws, wsb = _lin_ws(lib, M, N, K, dev)
    with torch.cuda.device(dev):
        check(lib.gaot_linear_forward(_p(x), _DT[x.dtype], x.stride(0), _p(x2), 0 if x2 is None else x2.stride(0), K1,
                                      _p(w), _DT[w.dtype], M, N, K, _p(bias), _p(residual),
                                      0 if residual is None else residual.stride(0), _p(y), _DT[out_dtype], y.stride(0),
                                      _p(ws), wsb, _stream(dev)), "linear_forward")
    return y


def _linear_bwd_x_raw(dy, w, out_dtype=torch.float32, residual=None):
    lib = _lib_()
    M, N = dy.shape
    K = w.shape[1]
    dev = dy.device
    dx = torch.empty(M, K, dtype=out_dtype, device=dev)
    if M == 0:
        return dx
    ws, wsb = _lin_ws(lib, M, N, K, dev)
    with torch.cuda.device(dev):
        check(lib.gaot_linear_backward_input(_p(dy), _DT[dy.dtype], dy.stride(0), _p(w), _DT[w.dtype], M, N, K,
                                             _p(residual), 0 if residual is None else residual.stride(0),
                                             _p(dx), _DT[out_dtype], dx.stride(0), 0, _p(ws), wsb, _stream(dev)),
              "linear_backward_input")
    return dx


def _linear_bwd_w_raw(dy, x, dw, accumulate=False):
    """dw[N, K'] (a column slice of the weight gradient, row stride dw.stride(0)) = dy^T x."""
    lib = _lib_()
    M, N = dy.shape
    K = x.shape[1]
    dev = dy.device
    if M == 0:
        if not accumulate:
            dw.zero_()
        return dw
    ws, wsb = _lin_ws(lib, M, N, K, dev)
    with torch.cuda.device(dev):
        check(lib.gaot_linear_backward_weight(_p(dy), _DT[dy.dtype], dy.stride(0), _p(x), _DT[x.dtype], x.stride(0),
                                              M, N, K, _p(dw), dw.stride(0), 1 if accumulate else 0, _p(ws), wsb,
                                              _stream(dev)), "linear_backward_weight")
    return dw


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x2, weight, bias, residual):
        shape = x.shape
        xr = _rows(x.detach())
        x2r = None if x2 is None else _rows(x2.detach())
        w = weight.detach().to(torch.float32).contiguous()
        b = None if bias is None else bias.detach().to(torch.float32).contiguous()
        r = None if residual is None else _rows(residual.detach()).to(torch.float32)
        y = _linear_fwd_raw(xr, x2r, w, b, r)
        ctx.save_for_backward(xr, x2r, w)
        ctx.meta = (shape, None if x2 is None else x2.shape, None if residual is None else residual.shape, bias is not None)
        return y.view(*shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xr, x2r, w = ctx.saved_tensors
        shape, shape2, rshape, has_bias = ctx.meta
        dyr = _rows(dy)
        need_x, need_x2, need_w, need_b, need_r = ctx.needs_input_grad
        dx = dx2 = dw = db = dr = None
        if need_x or (need_x2 and x2r is not None):
            dfull = _linear_bwd_x_raw(dyr, w)
            k1 = xr.shape[1]
            if need_x:
                dx = dfull[:, :k1].reshape(shape) if x2r is None else dfull[:, :k1].view(*shape[:-1], k1)
            if need_x2 and x2r is not None:
                dx2 = dfull[:, k1:].view(*shape2[:-1], shape2[-1])
        if need_w:
            dw = torch.empty_like(w)
            k1 = xr.shape[1]
            _linear_bwd_w_raw(dyr, xr, dw[:, :k1] if x2r is not None else dw)
            if x2r is not None:
                _linear_bwd_w_raw(dyr, x2r, dw[:, k1:])
        if need_b and has_bias:
            db = dyr.sum(0)
        if need_r and rshape is not None:
            dr = dy.reshape(rshape)
        return dx, dx2, dw, db, dr


def linear(x, weight, bias=None, residual=None, x2=None):
    """y = [x | x2] weight^T (+ bias) (+ residual) on tcgen05 (bf16 operands, fp32 accumulate) -- the nn.Linear
    of the latent transformer (reference attn.py:104-106,:129,:163,:223; gaot_3d.py:205).  `x2` is the second
    half of a concatenated input (skip_proj over cat[x, skip]) read in place."""
    _need_cuda(x, weight, bias, residual, x2)
    k1 = x.shape[-1]
    if not linear_supported(weight.shape[1], weight.shape[0]) or (x2 is not None and k1 % 64 != 0):
        raise NotImplementedError(f"linear: shape [{weight.shape[0]}, {weight.shape[1]}] outside the tensor-core kernel envelope")
    return _LinearFn.apply(x, x2, weight, bias, residual)


# ----------------------------------------------------------------------------- attention
class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, num_heads, num_kv_heads, freqs, dropout_p, seed):
        lib = _lib_()
        B, S, HD = q.shape
        d = HD // num_heads
        dev = q.device
        q, k, v = (t.to(torch.float32).contiguous() for t in (q, k, v))
        fr = None if freqs is None else freqs.detach().to(torch.float32).contiguous()
        out = torch.empty(B, S, HD, dtype=torch.float32, device=dev)
        lse = torch.empty(B, num_heads, S, dtype=torch.float32, device=dev)
        wsb = lib.gaot_attn_workspace_bytes(B, S, num_heads, num_kv_heads, d)
        ws = _ws(wsb, dev)
        with torch.cuda.device(dev), _timed("attn_fwd", dev):
            check(lib.gaot_attn_forward(_p(q), _p(k), _p(v), B, S, num_heads, num_kv_heads, d, _p(fr), float(dropout_p),
                                        int(seed), _p(ws), wsb, _p(out), _p(lse), _stream(dev)), "attn_forward")
        ctx.save_for_backward(q, k, v, out, lse, fr)
        ctx.cfg = (B, S, num_heads, num_kv_heads, d, float(dropout_p), int(seed))
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, out, lse, fr = ctx.saved_tensors
        B, S, H, Hkv, d, dropout_p, seed = ctx.cfg
        lib = _lib_()
        dev = q.device
        d_out = d_out.to(torch.float32).contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        wsb = lib.gaot_attn_workspace_bytes(B, S, H, Hkv, d)
        ws = _ws(wsb, dev)
        with torch.cuda.device(dev), _timed("attn_bwd", dev):
            check(lib.gaot_attn_backward(_p(q), _p(k), _p(v), _p(out), _p(d_out), _p(lse), B, S, H, Hkv, d, _p(fr),
                                         dropout_p, seed, _p(ws), wsb, _p(dq), _p(dk), _p(dv), _stream(dev)), "attn_backward")
        return dq, dk, dv, None, None, None, None, None


def attention(q, k, v, num_heads: int, num_kv_heads: int, rope_freqs: Optional[torch.Tensor] = None,
              dropout_p: float = 0.0, seed: Optional[int] = None):
    """softmax(QK^T/sqrt(d))V on token-major projections q [B,S,H*d], k/v [B,S,Hkv*d]
    (reference attn.py:110-128), optional 1-D RoPE with the module's `freqs`, optional dropout on the
    probabilities (counter-based mask; `seed` defaults to a draw from torch's CPU generator)."""
    _need_cuda(q, k, v)
    if dropout_p > 0.0 and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _AttnFn.apply(q, k, v, int(num_heads), int(num_kv_heads), rope_freqs, float(dropout_p), int(seed or 0))


def launch_count() -> int:
    return int(_lib_().gaot_launch_count())


def reset_launch_count() -> None:
    _lib_().gaot_launch_count_reset()
