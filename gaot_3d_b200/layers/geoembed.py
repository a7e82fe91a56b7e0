"""Geometric embedding -- drop-in for reference src/model/layers/geoembed.py ('statistical' and 'pointnet').

Parameter names match (`mlp.0`, `mlp.2`; `pointnet_mlp.0`, `pointnet_mlp.2`, `fc.0`).  'statistical': the per-query
statistics (count, mean/var of distance, centroid offset, covariance eigenvalues) and the global z-score are CUDA
kernels (geo.cu); the 9->64->out MLP on [N_q, 9] stays a torch GEMM.  'pointnet': the per-edge 3->32->32 ReLU MLP and
the max / mean pool over each query's edges are one CUDA kernel (pointnet.cu); the Linear(32->out) behind it is torch.
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops


class GeometricEmbedding(nn.Module):
    def __init__(self, input_dim, output_dim, method="statistical", pooling="max", **kwargs):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.method, self.pooling, self.kwargs = method.lower(), pooling.lower(), kwargs
        if self.pooling not in ("max", "mean"):
            raise ValueError(f"Unsupported pooling method: {self.pooling}. Supported methods: 'max', 'mean'.")
        if self.method == "statistical":
            self.mlp = nn.Sequential(nn.Linear(self._get_stat_feature_dim(), 64), nn.ReLU(), nn.Linear(64, output_dim))
        elif self.method == "pointnet":                      # reference geoembed.py:42-53
            self.pointnet_mlp = nn.Sequential(nn.Linear(input_dim, 32), nn.ReLU(), nn.Linear(32, 32), nn.ReLU())
            self.fc = nn.Sequential(nn.Linear(32, output_dim))
        else:
            raise ValueError(f"Unknown method: {self.method}")

    def _get_stat_feature_dim(self):
        return 3 + 2 * self.input_dim

    def statistical_features(self, source_pos, query_pos, edge_index, normalize=True):
        csr = ops.csr_of(edge_index, source_pos.shape[0], query_pos.shape[0])
        feat = ops.geo_stats(source_pos, query_pos, csr, normalize=normalize)
        if self.input_dim == 2:
            # points padded with z = 0: the 3x3 covariance has a zero row / column, its eigenvalues (+ 1e-6) in descending order
            # are the two of the 2x2 problem followed by 1e-6, the centroid offset's z is 0 -> the reference's 7 features
            # (geoembed.py:95-97) are columns [N, Davg, Dvar, dx, dy, l0, l1]; the z-score is per column, so selecting after it
            # is the same as before it
            feat = feat[:, [0, 1, 2, 3, 4, 6, 7]]
        return feat

    def forward(self, source_pos: torch.Tensor, query_pos: torch.Tensor, edge_index: torch.Tensor,
                batch_source: Optional[torch.Tensor] = None, batch_query: Optional[torch.Tensor] = None,
                neighbors_counts: Optional[torch.Tensor] = None) -> torch.Tensor:
        if neighbors_counts is not None:
            raise NotImplementedError("neighbors_counts override is unused by the reference (magno.py:513-516)")
        if self.input_dim not in (2, 3):
            raise NotImplementedError("geometric embedding: 2-D or 3-D coordinates")
        if self.method == "pointnet":
            return self.pointnet_features(source_pos, query_pos, edge_index)
        return self.mlp(self.statistical_features(source_pos, query_pos, edge_index))

    def pointnet_features(self, source_pos, query_pos, edge_index):
        """Reference geoembed.py:184-222: pooled per-edge MLP features -> fc; rows of queries without neighbours are 0."""
        nq = query_pos.shape[0]
        if edge_index.numel() == 0:
            return torch.zeros(nq, self.output_dim, device=query_pos.device, dtype=query_pos.dtype)
        csr = ops.csr_of(edge_index, source_pos.shape[0], nq)
        l1, l2 = self.pointnet_mlp[0], self.pointnet_mlp[2]
        w1 = l1.weight
        if self.input_dim == 2:        # the kernel reads zero-padded 3-D offsets: a zero weight column for the padded axis
            w1 = torch.cat([w1, w1.new_zeros(w1.shape[0], 1)], dim=1)
        pooled = ops.pointnet_pool(source_pos, query_pos, csr, w1, l1.bias, l2.weight, l2.bias, self.pooling)
        has = (csr.rowptr[1:] > csr.rowptr[:-1]).to(pooled.dtype).unsqueeze(1)
        return self.fc(pooled) * has
