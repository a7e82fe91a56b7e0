from .mlp import LinearChannelMLP, ChannelMLP
from .integral_transform import IntegralTransform
from .geoembed import GeometricEmbedding
from .magno import MAGNOConfig, MAGNOEncoder, MAGNODecoder
from .attn import (AttentionConfig, FFNConfig, TransformerConfig, GroupQueryFlashAttention, FFN, RMSNorm,
                   TransformerBlock, Transformer, RotaryEmbedding)
