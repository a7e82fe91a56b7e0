"""TEST INFRASTRUCTURE ONLY -- travelling torch-CPU restatement of the GNO hot ops.

The *real* reference code (IntegralTransform, GeometricEmbedding, scatter_native)
is importable in the dev container through ``oracle.ref_loader``; it cannot
travel to the GPU box, so the same math is restated here and pinned against the
reference modules in tests/test_oracle_vs_reference.py (dev container) and
against the committed fixtures tests/golden/gno_*.pt (everywhere).

Each function cites the reference lines it follows.
"""
import torch
import torch.nn.functional as F


def scatter_mean(src, index, dim_size):
    """reference scatter_native.py:23-31 -- scatter_add_ then / bincount.clamp(min=1)."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=dim_size).to(src.dtype)
    return out / cnt.clamp(min=1).view([-1] + [1] * (src.dim() - 1))


def scatter_sum(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def mlp_forward(x, weights, biases):
    """LinearChannelMLP.forward, reference mlp.py:327-335: Linear (+ exact-erf GELU between layers)."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w, b)
        if i < n - 1:
            x = F.gelu(x)
    return x


def segment_softmax(scores, index, dim_size):
    """IntegralTransform._segment_softmax_pyg, reference integral_transform.py:68-78."""
    smax = torch.zeros(dim_size, dtype=scores.dtype, device=scores.device).scatter_reduce(0, index, scores, reduce="amax", include_self=False)
    ex = torch.exp(scores - smax[index])
    den = torch.clamp(scatter_sum(ex, index, dim_size), min=torch.finfo(ex.dtype).tiny)
    return ex / den[index]


def integral_transform(y_pos, x_pos, edge_index, f_y, weights, biases, transform_type="linear", attn=None):
    """IntegralTransform.forward, reference integral_transform.py:80-175.

    out[q] = mean_{e: qry(e)=q} MLP(cat[y_pos[src], x_pos[q] (, f_y[src])]) (* f_y[src])
    attn = dict(type='cosine' | 'dot_product', coord_dim=D[, wq, bq, wk, bk]) selects the attentional variant
    (use_attn, :128-141): the per-edge values are weighted by the segment softmax of the scores and SUMMED (:161-165).
    """
    nq = x_pos.shape[0]
    if edge_index.shape[1] == 0:                                   # :107-112
        return torch.zeros(nq, weights[-1].shape[0], dtype=weights[-1].dtype, device=x_pos.device)
    src, qry = edge_index[0].long(), edge_index[1].long()          # :114-115
    rep, slf = y_pos[src], x_pos[qry]                              # :117-118
    inf = f_y[src] if f_y is not None else None                    # :120-123
    agg = torch.cat([rep, slf], dim=-1)                            # :146
    if inf is not None and transform_type in ("nonlinear", "nonlinear_kernelonly"):
        agg = torch.cat([agg, inf], dim=-1)                        # :148-152
    k = mlp_forward(agg, weights, biases)                          # :154
    if inf is not None and transform_type != "nonlinear_kernelonly":
        k = k * inf                                                # :156-157
    if attn is not None:
        d = attn["coord_dim"]
        qc, kc = slf[:, :d], rep[:, :d]                            # :130-131
        if attn["type"] == "dot_product":                          # :132-135
            sc = (F.linear(qc, attn["wq"], attn["bq"]) * F.linear(kc, attn["wk"], attn["bk"])).sum(-1) / (attn["wq"].shape[0] ** 0.5)
        else:                                                      # :136-139
            sc = (F.normalize(qc, p=2, dim=-1) * F.normalize(kc, p=2, dim=-1)).sum(-1)
        return scatter_sum(k * segment_softmax(sc, qry, nq).unsqueeze(-1), qry, nq)    # :141, :161-171 (reduce='sum')
    return scatter_mean(k, qry, nq)                                # :163-171 (reduce='mean')


def sym3_eigvals_desc(cov):
    return torch.linalg.eigvalsh(cov).flip(dims=[1])


def geo_statistical_features(source_pos, query_pos, edge_index, normalize=True):
    """GeometricEmbedding._compute_statistical_features_pyg, reference geoembed.py:99-182."""
    nq, nd = query_pos.shape
    src, qry = edge_index[0].long(), edge_index[1].long()
    n_i = torch.bincount(qry, minlength=nq).to(query_pos.dtype)    # :119-125
    has = n_i > 0
    nbr, qc = source_pos[src], query_pos[qry]                      # :129-130
    dist = torch.norm(nbr - qc, dim=1)                             # :132
    d_avg = scatter_mean(dist, qry, nq)                            # :133
    e_x2 = scatter_mean(dist ** 2, qry, nq)                        # :135-136
    d_var = torch.clamp(e_x2 - d_avg ** 2, min=0.0)                # :137-139
    cen = scatter_mean(nbr, qry, nq)                               # :142
    delta = cen - query_pos                                        # :143
    c = nbr - cen[qry]                                             # :146
    cov = scatter_sum(c.unsqueeze(2) * c.unsqueeze(1), qry, nq)    # :147-148
    cov = cov / n_i.clamp(min=1).view(-1, 1, 1)                    # :149-151
    pca = torch.zeros(nq, nd, dtype=query_pos.dtype)
    if has.any():                                                  # :155-162
        reg = cov[has] + 1e-6 * torch.eye(nd, dtype=cov.dtype).unsqueeze(0)
        pca[has] = sym3_eigvals_desc(reg)
    feat = torch.cat([n_i[:, None], d_avg[:, None], d_var[:, None], delta, pca], dim=1)  # :169-172
    feat[~has] = 0.0                                               # :175
    if not normalize:
        return feat
    mean = feat.mean(dim=0, keepdim=True)                          # :177-180
    std = feat.std(dim=0, keepdim=True)
    std[std < 1e-6] = 1.0
    return (feat - mean) / std


def geo_pointnet_embedding(source_pos, query_pos, edge_index, w1, b1, w2, b2, wf, bf, pooling="max"):
    """GeometricEmbedding._compute_pointnet_features_pyg, reference geoembed.py:184-222: per-edge
    relu(Linear(relu(Linear(y - x)))) (:206), scatter max | mean with a zero-initialised output (:208-213, the native
    scatter's amax with include_self=False keeps 0 for empty queries), fc (:215), rows without neighbours zeroed (:216)."""
    nq = query_pos.shape[0]
    out_dim = wf.shape[0]
    out = torch.zeros(nq, out_dim, dtype=query_pos.dtype)
    if edge_index.numel() == 0:
        return out
    src, qry = edge_index[0].long(), edge_index[1].long()
    has = torch.bincount(qry, minlength=nq) > 0
    h = F.relu(F.linear(F.relu(F.linear(source_pos[src] - query_pos[qry], w1, b1)), w2, b2))
    if pooling == "max":
        pooled = torch.zeros(nq, h.shape[1], dtype=h.dtype).scatter_reduce(0, qry[:, None].expand_as(h), h, reduce="amax", include_self=False)
    else:
        pooled = scatter_mean(h, qry, nq)
    res = F.linear(pooled, wf, bf)
    return torch.where(has[:, None], res, out)


def geo_embedding(source_pos, query_pos, edge_index, w0, b0, w1, b1):
    """GeometricEmbedding.forward('statistical'), reference geoembed.py:37-41,80-84."""
    f = geo_statistical_features(source_pos, query_pos, edge_index)
    return F.linear(F.relu(F.linear(f, w0, b0)), w1, b1)
