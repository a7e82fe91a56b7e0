"""Online bipartite graph build between physical points and latent tokens.

Drop-in for reference src/model/layers/magno.py:72-371 (`parse_neighbor_strategy`,
`get_neighbor_strategy`, `apply_neighbor_sampling`): same signatures, same edge_index
conventions (row 0 = source index, row 1 = query index, int64, global batched indices), same
error behaviour (ValueError on an unknown strategy, empty [2,0] tensor when nothing matches).
Underneath: the cell-list radius / kNN kernels and the radix-sort coalesce of libgaot_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from . import ops

PYG_MAX_NUM_NEIGHBORS = 32   # torch_geometric.nn.radius default, never overridden by the reference


def parse_neighbor_strategy(neighbor_strategy: Union[str, List[str]]) -> Tuple[str, str]:
    if isinstance(neighbor_strategy, str):
        return neighbor_strategy, neighbor_strategy
    if isinstance(neighbor_strategy, (list, tuple)) and len(neighbor_strategy) == 2:
        return neighbor_strategy[0], neighbor_strategy[1]
    raise ValueError(f"neighbor_strategy must be str or list of length 2, got {neighbor_strategy}")


def parse_geoembed_strategy(use_geoembed: Union[bool, List[bool]]) -> Tuple[bool, bool]:
    if isinstance(use_geoembed, bool):
        return use_geoembed, use_geoembed
    if isinstance(use_geoembed, (list, tuple)) and len(use_geoembed) == 2:
        return bool(use_geoembed[0]), bool(use_geoembed[1])
    raise ValueError(f"use_geoembed must be bool or list of length 2, got {use_geoembed}")


def _example_slices(batch: Optional[torch.Tensor], n: int, num_examples: int):
    """[(start, end)] per example from a sorted batch vector (PyG ptr semantics)."""
    if batch is None or num_examples == 1:
        return [(0, n)]
    ptr = torch.searchsorted(batch.contiguous(), torch.arange(num_examples + 1, device=batch.device, dtype=batch.dtype))
    ptr = ptr.tolist()
    return [(ptr[i], ptr[i + 1]) for i in range(num_examples)]


def _num_examples(*batches) -> int:
    b = 1
    for bt in batches:
        if bt is not None and bt.numel():
            b = max(b, int(bt[-1].item()) + 1)      # sorted ascending
    return b


def _per_example(fn, x, y, batch_x, batch_y):
    """Run a single-example search per batch element; indices are offset back to global."""
    B = _num_examples(batch_x, batch_y)
    if B == 1:
        return fn(x, y)
    if batch_x is None or batch_y is None:
        raise ValueError("batched neighbour search needs BOTH batch vectors (got one None with more than one example)")
    sx, sy = _example_slices(batch_x, x.shape[0], B), _example_slices(batch_y, y.shape[0], B)
    rows, cols = [], []
    for (x0, x1), (y0, y1) in zip(sx, sy):
        r, c = fn(x[x0:x1], y[y0:y1])
        rows.append(r + y0)
        cols.append(c + x0)
    return torch.cat(rows), torch.cat(cols)


# ----------------------------------------------------------------------------- sample-keyed search cache
# One forward asks for the same search several times: the decoder's 'reverse' graph is the flip of the *bidirectional* encoder
# graph (reference magno.py:263-273 rebuilds it from scratch), a knn decoder repeats the knn encoder's search, and every scale of a
# multi-scale model repeats the scale-independent knn part (magno.py:502, :711 loop over scales).  INSIDE ONE FORWARD (a
# `sample_scope()`, opened by GAOT3D.forward and shard.sharded_forward) results are kept in a small LRU keyed by the identity AND
# version counter of the position / batch tensors (an in-place update of the coordinates misses) plus the search parameters; the
# scope's exit drops everything, so nothing is ever reused from one sample / training step to the next (the online graph build
# stays inside every step) and nothing outlives the forward.  The entry holds references to its inputs, so a data pointer cannot
# be recycled while it is cached.  The cached tensors are handed out as they are: treat a returned edge_index as read-only.
_CACHE = {"on": True, "max": 8, "entries": {}, "depth": 0}


def set_graph_cache(on: bool = True, max_entries: int = 8) -> None:
    """Process-wide switch of the per-forward search cache (MAGNOConfig.use_graph_cache, unused by the reference, names the intent)."""
    _CACHE["on"], _CACHE["max"] = bool(on), int(max_entries)
    _CACHE["entries"].clear()


class sample_scope:
    """Context manager around the graph builds of ONE sample (one model forward); nests."""

    def __enter__(self):
        _CACHE["depth"] += 1
        return self

    def __exit__(self, *exc):
        _CACHE["depth"] -= 1
        if _CACHE["depth"] == 0:
            _CACHE["entries"].clear()
        return False


def _tkey(t: Optional[torch.Tensor]):
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape), t.dtype, t.device.index)


def _cached(kind, tensors, params, build):
    if not _CACHE["on"] or _CACHE["depth"] == 0:
        return build()
    key = (kind, tuple(_tkey(t) for t in tensors), params)
    ent = _CACHE["entries"].get(key)
    if ent is not None:
        _CACHE["entries"][key] = _CACHE["entries"].pop(key)          # most recently used last
        return ent[1]
    out = build()
    while len(_CACHE["entries"]) >= _CACHE["max"]:
        _CACHE["entries"].pop(next(iter(_CACHE["entries"])))
    _CACHE["entries"][key] = (tensors, out)
    return out


def radius_graph(x, y, r, batch_x=None, batch_y=None, max_num_neighbors: int = PYG_MAX_NUM_NEIGHBORS):
    """torch_geometric.nn.radius(x, y, r, batch_x, batch_y): returns [2,E], row 0 = y index, row 1 = x index."""
    def build():
        ry, cx = _per_example(lambda a, b: ops.radius(a, b, r, max_num_neighbors), x, y, batch_x, batch_y)
        return torch.stack([ry, cx])
    return _cached("radius", (x, y, batch_x, batch_y), (float(r), int(max_num_neighbors)), build)


def knn_graph(x, y, k, batch_x=None, batch_y=None):
    """torch_geometric.nn.knn(x, y, k, batch_x, batch_y): returns [2, ny*k], row 0 = y index, row 1 = x index."""
    def build():
        ry, cx = _per_example(lambda a, b: ops.knn(a, b, k), x, y, batch_x, batch_y)
        return torch.stack([ry, cx])
    return _cached("knn", (x, y, batch_x, batch_y), (int(k),), build)


def _tag(ei: torch.Tensor, query_sorted: bool) -> torch.Tensor:
    ei._gaot_query_sorted = query_sorted     # side-band hint: CSR build can skip its sort
    ei._gaot_trusted = True                  # built by this package's kernels: ops.csr_of skips the index validation
    return ei


def _coalesced(parts, n0: int, n1: int) -> torch.Tensor:
    cat = torch.cat(parts, dim=1)
    r0, r1 = ops.coalesce(cat[0], cat[1], max(n0 - 1, 0), max(n1 - 1, 0))
    return torch.stack([r0, r1])


def _encoder_edges(strategy, phys_pos, batch_phys, latent_pos, batch_latent, radius, k):
    n_phys, n_lat = phys_pos.shape[0], latent_pos.shape[0]
    if strategy == "knn":          # each physical point -> its k nearest latent tokens: [phys, latent]
        return _tag(knn_graph(latent_pos, phys_pos, k, batch_latent, batch_phys), False)
    if strategy == "radius":       # latent tokens as centres: raw [latent, phys] -> flip
        return _tag(radius_graph(phys_pos, latent_pos, radius, batch_phys, batch_latent).flip(0).contiguous(), True)
    if strategy == "bidirectional":
        def build():
            e_knn = knn_graph(latent_pos, phys_pos, k, batch_latent, batch_phys)
            e_rad = radius_graph(phys_pos, latent_pos, radius, batch_phys, batch_latent).flip(0)
            return _tag(_coalesced([e_knn, e_rad], n_phys, n_lat), False)
        return _cached("enc_bidirectional", (phys_pos, latent_pos, batch_phys, batch_latent), (float(radius), int(k)), build)
    raise ValueError(f"Unknown encoder strategy: {strategy}")


def _decoder_edges(strategy, phys_pos, batch_phys, latent_pos, batch_latent, radius, k):
    n_phys, n_lat = phys_pos.shape[0], latent_pos.shape[0]
    if strategy == "reverse":      # flip of the *bidirectional* encoder graph, whatever the encoder uses
        enc = _encoder_edges("bidirectional", phys_pos, batch_phys, latent_pos, batch_latent, radius, k)
        return _tag(enc.flip(0).contiguous(), True)
    if strategy == "knn":
        return _tag(knn_graph(latent_pos, phys_pos, k, batch_latent, batch_phys).flip(0).contiguous(), True)
    if strategy == "radius":
        return _tag(radius_graph(latent_pos, phys_pos, radius, batch_latent, batch_phys).flip(0).contiguous(), True)
    if strategy == "bidirectional":
        e_knn = knn_graph(latent_pos, phys_pos, k, batch_latent, batch_phys).flip(0)
        e_rad = radius_graph(latent_pos, phys_pos, radius, batch_latent, batch_phys).flip(0)
        return _tag(_coalesced([e_knn, e_rad], n_lat, n_phys), False)
    raise ValueError(f"Unknown decoder strategy: {strategy}")


def get_neighbor_strategy(neighbor_strategy: str, phys_pos: torch.Tensor, batch_idx_phys: torch.Tensor,
                          latent_tokens_pos: torch.Tensor, batch_idx_latent: torch.Tensor, radius: float,
                          k_neighbors: int = 1, is_decoder: bool = False) -> torch.Tensor:
    """Same contract as reference magno.py:116-163.  Encoder output is [phys_idx; latent_idx],
    decoder output [latent_idx; phys_idx].  CUDA tensors only (no CPU fallback by design: the
    reference's CPU callers -- collate_functions.py:97, stat.py:182 -- become unnecessary)."""
    if not phys_pos.is_cuda:
        raise RuntimeError("gaot_3d_b200.get_neighbor_strategy builds graphs on the GPU only; move the positions "
                           "to a CUDA device (set asynchronous_graph_building=False, precompute_edges=False)")
    fn = _decoder_edges if is_decoder else _encoder_edges
    return fn(neighbor_strategy, phys_pos, batch_idx_phys, latent_tokens_pos, batch_idx_latent,
              float(radius), int(k_neighbors))


# Philox stream position of the edge-dropout masks: one running counter per torch seed, advanced by the number of
# counter values a call consumes (ceil(E/4)), so the masks of different calls (encoder / decoder, scales, samples) never
# share random words; torch.manual_seed() restarts the stream (reproducible runs).
_mask_stream = {"seed": None, "offset": 0}


def apply_neighbor_sampling(edge_index: torch.Tensor, num_query_nodes: int, device=None,
                            sampling_strategy: Optional[str] = None, max_neighbors: Optional[int] = None,
                            sample_ratio: Optional[float] = None, training: bool = True) -> torch.Tensor:
    """Reference magno.py:297-371.  None -> passthrough; 'ratio' -> Philox edge mask kernel (training
    only); 'max_neighbors' -> random <= K edges per over-full query (kept in torch, rarely used)."""
    if sampling_strategy is None:
        return edge_index
    E = edge_index.shape[1]
    if num_query_nodes == 0 or E == 0:
        return edge_index
    if sampling_strategy == "ratio":
        if sample_ratio is None:
            raise ValueError("sample_ratio must be provided when using 'ratio' sampling strategy")
        if sample_ratio >= 1.0 or not training:
            return edge_index
        seed = int(torch.initial_seed())
        if _mask_stream["seed"] != seed:
            _mask_stream["seed"], _mask_stream["offset"] = seed, 0
        offset = _mask_stream["offset"]
        _mask_stream["offset"] = (offset + (E + 3) // 4) & (2 ** 64 - 1)
        r0, r1 = ops.edge_mask(edge_index[0], edge_index[1], 1.0 - float(sample_ratio), seed, offset=offset)
        out = torch.stack([r0, r1])
        out._gaot_query_sorted = bool(getattr(edge_index, "_gaot_query_sorted", False))
        out._gaot_trusted = bool(getattr(edge_index, "_gaot_trusted", False))
        return out
    if sampling_strategy == "max_neighbors":
        if max_neighbors is None:
            raise ValueError("max_neighbors must be provided when using 'max_neighbors' sampling strategy")
        dest = edge_index[1]
        counts = torch.bincount(dest, minlength=num_query_nodes)
        if not bool((counts > max_neighbors).any()):
            return edge_index
        # random rank inside each query segment; keep ranks < max_neighbors (one vectorised pass
        # instead of the reference's Python loop over over-full queries)
        key = torch.rand(E, device=edge_index.device)
        order = torch.argsort(dest.to(torch.float64) + key.to(torch.float64) * 0.5, stable=True)
        sorted_dest = dest[order]
        start = torch.cumsum(counts, 0) - counts
        rank = torch.arange(E, device=edge_index.device) - start[sorted_dest]
        keep = torch.zeros(E, dtype=torch.bool, device=edge_index.device)
        keep[order[rank < max_neighbors]] = True
        return edge_index[:, keep]
    raise ValueError(f"Invalid sampling strategy: {sampling_strategy}")
