"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the bipartite graph build.

Restates, in numpy (exact fp32 arithmetic, int64 indices):

* ``torch_cluster.radius`` / ``torch_cluster.knn`` as called through
  ``torch_geometric.nn.{radius,knn}`` by the reference at
  src/model/layers/magno.py:183-200 and :242-260.  torch-cluster is an
  UN-VENDORED, UNPINNED third-party wheel (README.md:112-113 only suggests the
  wheel index for torch 2.7.0); its source is not under /root/reference, so the
  published algorithm is restated here:
    - radius: for every y, scan x of the same example in ascending index,
      keep x when fp32 ``((dx*dx + dy*dy) + dz*dz) < fl32(double(r)*double(r))``,
      stop after ``max_num_neighbors`` (PyG default 32 -- the reference leaves
      the argument commented out, magno.py:199,259).  This is the CUDA kernel's
      "first 32 by ascending x index"; the CPU wheel's nanoflann traversal order
      is not reproducible by any other implementation.
    - knn: for every y the k nearest x (same fp32 distance), ascending distance,
      ties -> lower x index first (strict ``best > d`` insertion), fewer than k
      rows when the example has fewer than k points.
* ``torch_geometric.utils.coalesce`` (App. A3): lexicographic sort + unique.
* ``torch_geometric.utils.dropout_edge`` (App. A4).
* the composition rules of ``get_neighbor_strategy`` (magno.py:116-295):
  flips, bidirectional = coalesce(cat(knn, radius)), reverse = flip of the
  *bidirectional* encoder graph (magno.py:263-273).

PARITY UNPINNED against a real torch_cluster install -- the reference ships no tests / golden vectors for this path
 and torch_cluster is absent from this image, so this
restatement is pinned (a) against an independent brute-force O(N*M) evaluation
of the same definition (tests/test_oracle_graph.py) and (b) against the
committed fixtures in tests/golden/graph_*.npz produced by that brute force.
"parity unpinned" w.r.t. a real torch_cluster install -- stated in DESIGN.md.
"""
from __future__ import annotations

import numpy as np

try:  # scipy is present in this image; brute force is the fallback for tiny cases
    from scipy.spatial import cKDTree
except Exception:  # pragma: no cover
    cKDTree = None

F32 = np.float32


def _as_f32(a):
    a = np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("positions must be [N, D]")
    return a


def r2_f32(r: float) -> np.float32:
    """fl32(double(r) * double(r)) -- torch_cluster passes ``r * r`` (double) cast to scalar_t."""
    return np.float32(float(r) * float(r))


def dist2_f32(xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """fp32, non-fused, left-to-right: ((dx*dx + dy*dy) + dz*dz)."""
    d = (xs - ys).astype(F32)
    acc = (d[:, 0] * d[:, 0]).astype(F32)
    for j in range(1, d.shape[1]):
        acc = (acc + (d[:, j] * d[:, j]).astype(F32)).astype(F32)
    return acc


def _ptr(batch, n, num_examples):
    if batch is None:
        return np.array([0, n], dtype=np.int64)
    batch = np.asarray(batch, dtype=np.int64)
    if batch.size and np.any(np.diff(batch) < 0):
        raise ValueError("batch vector must be sorted ascending")
    return np.searchsorted(batch, np.arange(num_examples + 1), side="left").astype(np.int64)


def _num_examples(batch_x, batch_y):
    b = 1
    for bt in (batch_x, batch_y):
        if bt is not None and len(bt):
            b = max(b, int(np.asarray(bt).max()) + 1)
    return b


# --------------------------------------------------------------------------- radius
def radius_bruteforce(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """Literal O(N_x*N_y) evaluation of the definition (small cases only)."""
    x, y = _as_f32(x), _as_f32(y)
    B = _num_examples(batch_x, batch_y)
    px, py = _ptr(batch_x, len(x), B), _ptr(batch_y, len(y), B)
    r2 = r2_f32(r)
    rows, cols = [], []
    for b in range(B):
        xs = x[px[b]:px[b + 1]]
        for j in range(py[b], py[b + 1]):
            if len(xs) == 0:
                continue
            d2 = dist2_f32(xs, y[j][None, :])
            hit = np.nonzero(d2 < r2)[0][:max_num_neighbors]
            rows.append(np.full(len(hit), j, dtype=np.int64))
            cols.append(hit.astype(np.int64) + px[b])
    if not rows:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(rows), np.concatenate(cols)])


def radius_np(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, workers=1):
    """KD-tree candidate generation + exact fp32 membership test + index-ordered cap."""
    x, y = _as_f32(x), _as_f32(y)
    if len(x) == 0 or len(y) == 0:
        return np.zeros((2, 0), dtype=np.int64)
    if cKDTree is None:
        return radius_bruteforce(x, y, r, batch_x, batch_y, max_num_neighbors)
    B = _num_examples(batch_x, batch_y)
    px, py = _ptr(batch_x, len(x), B), _ptr(batch_y, len(y), B)
    r2 = r2_f32(r)
    scale = max(1.0, float(np.abs(x).max()), float(np.abs(y).max()))
    r_search = float(r) * (1.0 + 1e-5) + 4e-7 * scale
    out_r, out_c = [], []
    for b in range(B):
        xs, ys = x[px[b]:px[b + 1]], y[py[b]:py[b + 1]]
        if len(xs) == 0 or len(ys) == 0:
            continue
        tree = cKDTree(xs.astype(np.float64))
        lists = tree.query_ball_point(ys.astype(np.float64), r_search, workers=workers, return_sorted=True)
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        if lens.sum() == 0:
            continue
        cand = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists if len(l)])
        qry = np.repeat(np.arange(len(ys), dtype=np.int64), lens)
        keep = dist2_f32(xs[cand], ys[qry]) < r2
        cand, qry = cand[keep], qry[keep]
        # rank of each kept candidate inside its (ascending-x) query group
        cnt = np.bincount(qry, minlength=len(ys))
        start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        rank = np.arange(len(qry), dtype=np.int64) - start[qry]
        sel = rank < max_num_neighbors
        out_r.append(qry[sel] + py[b])
        out_c.append(cand[sel] + px[b])
    if not out_r:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(out_r), np.concatenate(out_c)]).astype(np.int64)


# --------------------------------------------------------------------------- knn
def knn_bruteforce(x, y, k, batch_x=None, batch_y=None):
    x, y = _as_f32(x), _as_f32(y)
    B = _num_examples(batch_x, batch_y)
    px, py = _ptr(batch_x, len(x), B), _ptr(batch_y, len(y), B)
    rows, cols = [], []
    for b in range(B):
        xs = x[px[b]:px[b + 1]]
        for j in range(py[b], py[b + 1]):
            if len(xs) == 0:
                continue
            d2 = dist2_f32(xs, y[j][None, :])
            order = np.lexsort((np.arange(len(xs)), d2))[:k]   # (d2, idx) ascending, stable ties
            order = order[d2[order] < np.float32(1e10)]
            rows.append(np.full(len(order), j, dtype=np.int64))
            cols.append(order.astype(np.int64) + px[b])
    if not rows:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(rows), np.concatenate(cols)])


def knn_np(x, y, k, batch_x=None, batch_y=None, workers=1):
    x, y = _as_f32(x), _as_f32(y)
    if len(x) == 0 or len(y) == 0:
        return np.zeros((2, 0), dtype=np.int64)
    if cKDTree is None:
        return knn_bruteforce(x, y, k, batch_x, batch_y)
    B = _num_examples(batch_x, batch_y)
    px, py = _ptr(batch_x, len(x), B), _ptr(batch_y, len(y), B)
    scale = max(1.0, float(np.abs(x).max()), float(np.abs(y).max()))
    out_r, out_c = [], []
    for b in range(B):
        xs, ys = x[px[b]:px[b + 1]], y[py[b]:py[b + 1]]
        if len(xs) == 0 or len(ys) == 0:
            continue
        kk = min(k, len(xs))
        tree = cKDTree(xs.astype(np.float64))
        dk, _ = tree.query(ys.astype(np.float64), k=[kk], workers=workers)
        # every point that could beat/tie the k-th in fp32 arithmetic
        rad = dk[:, 0] * (1.0 + 1e-5) + 4e-7 * scale
        lists = tree.query_ball_point(ys.astype(np.float64), rad, workers=workers, return_sorted=True)
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        cand = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists])
        qry = np.repeat(np.arange(len(ys), dtype=np.int64), lens)
        d2 = dist2_f32(xs[cand], ys[qry])
        order = np.lexsort((cand, d2, qry))          # by query, then d2, then index
        cand, qry, d2 = cand[order], qry[order], d2[order]
        start = np.concatenate([[0], np.cumsum(lens)[:-1]])
        rank = np.arange(len(qry), dtype=np.int64) - start[qry]
        sel = (rank < kk) & (d2 < np.float32(1e10))
        out_r.append(qry[sel] + py[b])
        out_c.append(cand[sel] + px[b])
    if not out_r:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(out_r), np.concatenate(out_c)]).astype(np.int64)


# --------------------------------------------------------------------------- pyg utils
def coalesce_np(edge_index):
    """torch_geometric.utils.coalesce defaults: sort by (row0,row1), drop duplicates."""
    ei = np.asarray(edge_index, dtype=np.int64)
    if ei.shape[1] == 0:
        return ei.reshape(2, 0)
    num_nodes = int(ei.max()) + 1
    key = ei[0] * num_nodes + ei[1]
    key = np.unique(key)
    return np.stack([key // num_nodes, key % num_nodes]).astype(np.int64)


def get_neighbor_strategy_np(neighbor_strategy, phys_pos, batch_idx_phys, latent_pos, batch_idx_latent,
                             radius, k_neighbors=1, is_decoder=False, max_num_neighbors=32, workers=1):
    """Restates reference magno.py:116-295 on top of radius_np / knn_np / coalesce_np."""
    flip = lambda e: e[::-1].copy()
    empty = np.zeros((2, 0), dtype=np.int64)

    def enc(strategy):
        e_knn = e_rad = None
        if strategy in ("knn", "bidirectional"):       # magno.py:181-189 -> [phys, latent]
            e_knn = knn_np(latent_pos, phys_pos, k_neighbors, batch_idx_latent, batch_idx_phys, workers)
        if strategy in ("radius", "bidirectional"):    # magno.py:191-201 -> raw [latent, phys] -> flip
            e_rad = flip(radius_np(phys_pos, latent_pos, radius, batch_idx_phys, batch_idx_latent,
                                   max_num_neighbors, workers))
        if strategy == "knn":
            return e_knn
        if strategy == "radius":
            return e_rad
        if strategy == "bidirectional":
            return coalesce_np(np.concatenate([e_knn, e_rad], axis=1))
        raise ValueError(f"Unknown encoder strategy: {strategy}")

    if not is_decoder:
        return enc(neighbor_strategy)
    if neighbor_strategy == "reverse":                 # magno.py:263-273
        return flip(enc("bidirectional"))
    e_knn = e_rad = None
    if neighbor_strategy in ("knn", "bidirectional"):  # magno.py:240-249
        e_knn = flip(knn_np(latent_pos, phys_pos, k_neighbors, batch_idx_latent, batch_idx_phys, workers))
    if neighbor_strategy in ("radius", "bidirectional"):  # magno.py:251-261
        e_rad = flip(radius_np(latent_pos, phys_pos, radius, batch_idx_latent, batch_idx_phys,
                               max_num_neighbors, workers))
    if neighbor_strategy == "knn":
        return e_knn
    if neighbor_strategy == "radius":
        return e_rad
    if neighbor_strategy == "bidirectional":
        return coalesce_np(np.concatenate([e_knn, e_rad], axis=1))
    raise ValueError(f"Unknown decoder strategy: {neighbor_strategy}")


# --------------------------------------------------------------------------- torch adapters (ref_loader stubs)
def _t2n(t):
    return None if t is None else t.detach().cpu().numpy()


def radius_torch(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    import torch
    out = radius_np(_t2n(x), _t2n(y), r, _t2n(batch_x), _t2n(batch_y), max_num_neighbors)
    return torch.from_numpy(out).to(x.device)


def knn_torch(x, y, k, batch_x=None, batch_y=None):
    import torch
    out = knn_np(_t2n(x), _t2n(y), k, _t2n(batch_x), _t2n(batch_y))
    return torch.from_numpy(out).to(x.device)


def coalesce_torch(edge_index, *a, **kw):
    import torch
    return torch.from_numpy(coalesce_np(_t2n(edge_index))).to(edge_index.device)


def dropout_edge_torch(edge_index, p=0.5, force_undirected=False, training=True):
    import torch
    if not training or p == 0.0:
        return edge_index, edge_index.new_ones(edge_index.size(1), dtype=torch.bool)
    mask = torch.rand(edge_index.size(1), device=edge_index.device) >= p
    return edge_index[:, mask], mask


def sort_edges(ei: np.ndarray) -> np.ndarray:
    """Lexicographic (row0,row1) order -- the comparison form of SURVEY §8(c) protocol (1)."""
    ei = np.asarray(ei, dtype=np.int64)
    o = np.lexsort((ei[1], ei[0]))
    return ei[:, o]
