"""gaot_3d_b200 -- B200-native (sm_100a) hot path of GAOT-3D behind the reference's Python API.

(The directory name normalises the hyphen of "gaot-3d_b200" to an underscore so that it is an
importable Python package.)
"""
from . import ops
from .graph import apply_neighbor_sampling, get_neighbor_strategy, parse_neighbor_strategy
from .layers import (AttentionConfig, FFNConfig, GeometricEmbedding, IntegralTransform, LinearChannelMLP,
                     MAGNOConfig, MAGNODecoder, MAGNOEncoder, Transformer, TransformerConfig)
from .model import GAOT3D, Batch, init_model
from .ops import (set_gno_precision, get_gno_precision, set_node_mlp_tf32, set_node_mlp_mode, set_transformer_precision,
                  transformer_precision)

__all__ = [
    "ops", "get_neighbor_strategy", "parse_neighbor_strategy", "apply_neighbor_sampling", "MAGNOConfig",
    "MAGNOEncoder", "MAGNODecoder", "IntegralTransform", "GeometricEmbedding", "LinearChannelMLP",
    "AttentionConfig", "FFNConfig", "TransformerConfig", "Transformer", "GAOT3D", "init_model", "Batch",
    "set_gno_precision", "get_gno_precision", "set_node_mlp_tf32", "set_node_mlp_mode", "set_transformer_precision",
    "transformer_precision",
]
