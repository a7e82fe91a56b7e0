"""MAGNO encoder / decoder modules -- drop-in for reference src/model/layers/magno.py:21-66,377-798.

`MAGNOConfig` keeps every field name and default of the reference dataclass (it is the config
schema); the module shells keep constructor arguments, forward signatures, attribute and
parameter names (`gno.channel_mlp.fcs.*`, `lifting.fcs.0.*`, `geoembed.mlp.{0,2}.*`,
`recovery.fcs.0.*`, `projection.fcs.{0,1}.*`, `scale_weighting.*`) so a reference checkpoint loads
with strict=True.  Graph build, GNO and geometric statistics run on the CUDA kernels of this package.
"""
from dataclasses import dataclass, field
from typing import Any, Optional

import torch
import torch.nn as nn

from .. import ops
import torch.nn.functional as F

from .geoembed import GeometricEmbedding
from .integral_transform import IntegralTransform
from .mlp import ChannelMLP, LinearChannelMLP
from ..graph import (apply_neighbor_sampling, get_neighbor_strategy, parse_geoembed_strategy,
                     parse_neighbor_strategy)


@dataclass
class MAGNOConfig:
    use_gno: bool = True
    gno_coord_dim: int = 2
    gno_radius: float = 0.033
    lifting_channels: int = 16
    encoder_feature_attr: Any = "x"
    in_gno_channel_mlp_hidden_layers: list = field(default_factory=lambda: [64, 64, 64])
    in_gno_transform_type: str = "linear"
    projection_channels: int = 256
    out_gno_channel_mlp_hidden_layers: list = field(default_factory=lambda: [64, 64])
    out_gno_transform_type: str = "linear"
    mlp_type: str = "channel"
    scales: list = field(default_factory=lambda: [1.0])
    use_scale_weights: bool = False
    use_graph_cache: bool = True
    gno_use_torch_cluster: bool = False
    gno_use_torch_scatter: str = True
    node_embedding: bool = False
    use_attn: Optional[bool] = None
    attention_type: str = "cosine"
    use_geoembed: Any = field(default_factory=lambda: [True, True])
    embedding_method: str = "statistical"
    pooling: str = "max"
    sampling_strategy: Optional[str] = None
    max_neighbors: Optional[int] = None
    sample_ratio: Optional[float] = None
    neighbor_strategy: Any = "radius"
    k_neighbors: int = 1
    precompute_edges: bool = True
    asynchronous_graph_building: bool = False


def _node_mlp(mlp_type, cin, cout, hidden=None):
    """1- or 2-layer node MLP in the flavour selected by `mlp_type` (reference magno.py:421-430,650-661)."""
    if mlp_type == "linear":
        return LinearChannelMLP(layers=[cin, cout] if hidden is None else [cin, hidden, cout])
    if hidden is None:
        return ChannelMLP(in_channels=cin, out_channels=cout, n_layers=1)
    return ChannelMLP(in_channels=cin, out_channels=cout, hidden_channels=hidden, n_layers=2, n_dim=1)


def _apply_node_mlp(mlp, mlp_type, x):
    # two-layer GELU MLPs of the projection-head shape run as ONE fused tensor-core kernel when the process opted into the
    # mixed-precision tier (ops.set_node_mlp_mode("fused")); Linear and kernel-size-1 Conv1d hold the same [out, in] matrices
    if ops.node_mlp_mode() == "fused" and x.is_cuda and getattr(mlp, "n_layers", 0) == 2 and mlp.dropout is None \
            and mlp.non_linearity is F.gelu:
        w1, w2 = mlp.fcs[0].weight, mlp.fcs[1].weight
        w1, w2 = w1.reshape(w1.shape[0], -1), w2.reshape(w2.shape[0], -1)
        if mlp.fcs[0].bias is not None and mlp.fcs[1].bias is not None and x.dim() == 2 and \
                ops.node_mlp2_supported(w1.shape[1], w1.shape[0], w2.shape[0]):
            return ops.node_mlp2(x, w1, mlp.fcs[0].bias, w2, mlp.fcs[1].bias)
    # one-layer MLPs with a tiny input width (the lifting layer: raw point features -> lifting channels over every physical
    # point) stream through a dedicated fp32 kernel in every mode: as a library GEMM the shape is all padding and its weight
    # gradient a 10^6-row reduction
    if x.is_cuda and getattr(mlp, "n_layers", 0) == 1 and mlp.dropout is None and x.dim() == 2 and x.dtype == torch.float32:
        w = mlp.fcs[0].weight
        w = w.reshape(w.shape[0], -1)
        if ops.node_linear_supported(w.shape[1], w.shape[0]):
            return ops.node_linear(x, w, mlp.fcs[0].bias)
    return mlp(x) if mlp_type == "linear" else mlp(x.transpose(0, 1)).transpose(0, 1)


class _ScaleMixin:
    def _init_scale_weights(self, cfg):
        self.use_scale_weights = cfg.use_scale_weights
        if self.use_scale_weights:
            self.num_scales = len(self.scales)
            self.scale_weighting = nn.Sequential(nn.Linear(self.coord_dim, 16), nn.ReLU(), nn.Linear(16, self.num_scales))
            self.scale_weight_activation = nn.Softmax(dim=-1)

    def _aggregate(self, per_scale, pos):
        if len(per_scale) == 1:
            return per_scale[0]
        stack = torch.stack(per_scale, dim=0)
        if self.use_scale_weights:
            w = self.scale_weight_activation(self.scale_weighting(pos))      # [N, num_scales]
            return (stack * w.permute(1, 0).unsqueeze(-1)).sum(dim=0)
        return stack.sum(dim=0)

    def _sampling(self, edge_index, nq, device):
        out = apply_neighbor_sampling(edge_index, nq, device, self.sampling_strategy, self.max_neighbors,
                                      self.sample_ratio, self.training)
        return out


class MAGNOEncoder(nn.Module, _ScaleMixin):
    def __init__(self, in_channels, out_channels, gno_config: MAGNOConfig):
        super().__init__()
        c = gno_config
        self.gno_radius, self.scales = c.gno_radius, c.scales
        self.lifting_channels, self.coord_dim = c.lifting_channels, c.gno_coord_dim
        self.feature_attr_name = c.encoder_feature_attr
        self.precompute_edges, self.mlp_type = c.precompute_edges, c.mlp_type
        self.encoder_strategy, self.decoder_strategy = parse_neighbor_strategy(c.neighbor_strategy)
        self.k_neighbors = c.k_neighbors
        self.sampling_strategy, self.max_neighbors, self.sample_ratio = c.sampling_strategy, c.max_neighbors, c.sample_ratio
        self.use_gno = c.use_gno
        if self.use_gno:
            kin = self.coord_dim * 2 + (in_channels if c.in_gno_transform_type in ("nonlinear", "nonlinear_kernelonly") else 0)
            self.gno = IntegralTransform(channel_mlp_layers=[kin, *c.in_gno_channel_mlp_hidden_layers, self.lifting_channels],
                                         transform_type=c.in_gno_transform_type, use_attn=c.use_attn,
                                         coord_dim=self.coord_dim, attention_type=c.attention_type)
            self.lifting = _node_mlp(c.mlp_type, in_channels, self.lifting_channels)
        else:
            self.gno, self.lifting = None, None
        self.use_geoembed = parse_geoembed_strategy(c.use_geoembed)[0]
        if self.use_geoembed:
            self.geoembed = GeometricEmbedding(input_dim=self.coord_dim, output_dim=self.lifting_channels,
                                               method=c.embedding_method, pooling=c.pooling)
            self.recovery = _node_mlp(c.mlp_type, 2 * self.lifting_channels, self.lifting_channels)
        self._init_scale_weights(c)

    def _features(self, batch):
        names = self.feature_attr_name if isinstance(self.feature_attr_name, (list, tuple)) else [self.feature_attr_name]
        feats = []
        for n in names:
            f = getattr(batch, n, None)
            if f is None:
                if self.use_gno:
                    raise AttributeError(f"MAGNOEncoder requires feature attribute '{n}' but it was not found in the batch.")
            else:
                feats.append(f)
        if not feats:
            return None
        return feats[0] if len(feats) == 1 else torch.cat(feats, dim=-1)

    def forward(self, batch, latent_tokens_pos: torch.Tensor, latent_tokens_batch_idx: torch.Tensor) -> torch.Tensor:
        phys_pos, batch_idx_phys = batch.pos, batch.batch
        device = phys_pos.device
        num_graphs = batch.num_graphs
        m_per_graph = latent_tokens_pos.shape[0] // num_graphs
        phys_feat = self._features(batch)
        lifted = None
        encoded_scales = []
        for si, scale in enumerate(self.scales):
            if self.precompute_edges:
                name = f"encoder_edge_index_s{si}"
                if not hasattr(batch, name):
                    raise AttributeError(f"Batch object missing pre-computed '{name}'")
                edge_index = getattr(batch, name).to(device)
            else:
                edge_index = get_neighbor_strategy(self.encoder_strategy, phys_pos, batch_idx_phys, latent_tokens_pos,
                                                   latent_tokens_batch_idx, self.gno_radius * scale, self.k_neighbors,
                                                   is_decoder=False)
            edge_index = self._sampling(edge_index, latent_tokens_pos.shape[0], device)
            enc_gno = geo = None
            if self.use_gno:
                if lifted is None:       # same value for every scale (reference recomputes it per scale)
                    lifted = _apply_node_mlp(self.lifting, self.mlp_type, phys_feat)
                enc_gno = self.gno(y_pos=phys_pos, x_pos=latent_tokens_pos, edge_index=edge_index, f_y=lifted,
                                   batch_y=batch_idx_phys, batch_x=latent_tokens_batch_idx)
            if self.use_geoembed:
                geo = self.geoembed(source_pos=phys_pos, query_pos=latent_tokens_pos, edge_index=edge_index,
                                    batch_source=batch_idx_phys, batch_query=latent_tokens_batch_idx)
            if self.use_gno and self.use_geoembed:
                enc = _apply_node_mlp(self.recovery, self.mlp_type, torch.cat([enc_gno, geo], dim=-1))
            elif self.use_gno:
                enc = enc_gno
            elif self.use_geoembed:
                enc = geo
            else:
                raise ValueError("GNO and GeoEmbed are both disabled. No encoding will be performed.")
            encoded_scales.append(enc)
        out = self._aggregate(encoded_scales, latent_tokens_pos)
        return out.view(num_graphs, m_per_graph, self.lifting_channels)


class MAGNODecoder(nn.Module, _ScaleMixin):
    def __init__(self, in_channels, out_channels, gno_config: MAGNOConfig):
        super().__init__()
        c = gno_config
        self.gno_radius, self.scales, self.coord_dim = c.gno_radius, c.scales, c.gno_coord_dim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.use_geoembed = parse_geoembed_strategy(c.use_geoembed)[1]
        self.precompute_edges, self.mlp_type = c.precompute_edges, c.mlp_type
        self.encoder_strategy, self.decoder_strategy = parse_neighbor_strategy(c.neighbor_strategy)
        self.k_neighbors = c.k_neighbors
        self.sampling_strategy, self.max_neighbors, self.sample_ratio = c.sampling_strategy, c.max_neighbors, c.sample_ratio
        kin = self.coord_dim * 2 + (in_channels if c.out_gno_transform_type in ("nonlinear", "nonlinear_kernelonly") else 0)
        self.gno = IntegralTransform(channel_mlp_layers=[kin, *c.out_gno_channel_mlp_hidden_layers, in_channels],
                                     transform_type=c.out_gno_transform_type, use_attn=c.use_attn,
                                     coord_dim=self.coord_dim, attention_type=c.attention_type)
        self.projection = _node_mlp(c.mlp_type, in_channels, out_channels, hidden=c.projection_channels)
        if self.use_geoembed:
            self.geoembed = GeometricEmbedding(input_dim=self.coord_dim, output_dim=in_channels,
                                               method=c.embedding_method, pooling=c.pooling)
            self.recovery = _node_mlp(c.mlp_type, 2 * in_channels, in_channels)
        self._init_scale_weights(c)

    def forward(self, rndata_flat: torch.Tensor, phys_pos_query: torch.Tensor, batch_idx_phys_query: torch.Tensor,
                latent_tokens_pos: torch.Tensor, latent_tokens_batch_idx: torch.Tensor, batch=None) -> torch.Tensor:
        device = rndata_flat.device
        decoded_scales = []
        for si, scale in enumerate(self.scales):
            if self.precompute_edges:
                name = f"decoder_edge_index_s{si}"
                if not hasattr(batch, name):
                    raise AttributeError(f"Batch object missing pre-computed '{name}'")
                edge_index = getattr(batch, name).to(device)
            else:
                edge_index = get_neighbor_strategy(self.decoder_strategy, phys_pos_query, batch_idx_phys_query,
                                                   latent_tokens_pos, latent_tokens_batch_idx, self.gno_radius * scale,
                                                   self.k_neighbors, is_decoder=True)
            edge_index = self._sampling(edge_index, phys_pos_query.shape[0], device)
            dec = self.gno(y_pos=latent_tokens_pos, x_pos=phys_pos_query, edge_index=edge_index, f_y=rndata_flat,
                           batch_y=latent_tokens_batch_idx, batch_x=batch_idx_phys_query)
            if self.use_geoembed:
                geo = self.geoembed(source_pos=latent_tokens_pos, query_pos=phys_pos_query, edge_index=edge_index,
                                    batch_source=latent_tokens_batch_idx, batch_query=batch_idx_phys_query)
                dec = _apply_node_mlp(self.recovery, self.mlp_type, torch.cat([dec, geo], dim=-1))
            decoded_scales.append(dec)
        out = self._aggregate(decoded_scales, phys_pos_query)
        return _apply_node_mlp(self.projection, self.mlp_type, out)
