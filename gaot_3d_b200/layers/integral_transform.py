"""GNO kernel integral -- drop-in for reference src/model/layers/integral_transform.py.

Same constructor, forward signature and parameter names (`channel_mlp.fcs.{i}`); the forward
is ONE fused CUDA kernel (gather -> per-edge MLP -> (* f_y) -> CSR segmented mean) and the
backward one more (recompute-in-backward), instead of ~20 library kernels and [E,*] temporaries.
"""
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .mlp import LinearChannelMLP
from .. import ops


class IntegralTransform(nn.Module):
    def __init__(self, channel_mlp=None, channel_mlp_layers=None, channel_mlp_non_linearity=F.gelu,
                 transform_type="linear", use_attn=None, coord_dim=None, attention_type="cosine"):
        super().__init__()
        self.transform_type = transform_type
        self.use_attn = use_attn
        self.coord_dim = coord_dim
        self.attention_type = attention_type
        if channel_mlp is None:
            if channel_mlp_layers is None:
                raise ValueError("Need channel_mlp or layers")
            channel_mlp = LinearChannelMLP(layers=channel_mlp_layers, non_linearity=channel_mlp_non_linearity)
        self.channel_mlp = channel_mlp
        if channel_mlp_non_linearity is not F.gelu:
            raise NotImplementedError("the fused GNO kernel implements the reference's exact-erf GELU only")
        if self.use_attn:
            # segment-softmax attention weights are a later row
            raise NotImplementedError("use_attn (segment-softmax GNO weights) is not built yet; the reference default is None")
        if transform_type not in ("linear", "nonlinear", "nonlinear_kernelonly"):
            raise ValueError(f"unknown transform_type {transform_type}")

    def forward(self, y_pos: torch.Tensor, x_pos: torch.Tensor, edge_index: torch.Tensor,
                f_y: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None,
                batch_y=None, batch_x=None, reduce: str = "mean") -> torch.Tensor:
        """y_pos [N_y,D] sources, x_pos [N_x,D] queries, edge_index [2,E] (row 0 -> y, row 1 -> x),
        f_y [N_y,C].  `weights`, `batch_y`, `batch_x` are accepted and unused, as in the reference."""
        device = x_pos.device
        nq = x_pos.shape[0]
        fcs = self.channel_mlp.fcs
        edge_index = edge_index.to(device)
        if edge_index.shape[1] == 0:                       # reference :107-112
            return torch.zeros(nq, fcs[-1].out_features, device=device, dtype=fcs[-1].weight.dtype)
        csr = ops.csr_of(edge_index, y_pos.shape[0], nq)
        return ops.gno(y_pos, x_pos, f_y, csr, [fc.weight for fc in fcs], [fc.bias for fc in fcs],
                       transform_type=self.transform_type, reduce=reduce)
