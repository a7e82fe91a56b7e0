"""GAOT3D model assembly -- drop-in for reference src/model/gaot_3d.py and src/model/__init__.py.

Same constructor, forward signature, buffer (`latent_tokens`) and parameter names
(`encoder.*`, `patch_linear.*`, `processor.*`, `decoder.*`).
"""
from typing import Optional

import torch
import torch.nn as nn

from .layers.attn import Transformer, TransformerConfig, _lin
from .layers.magno import MAGNOConfig, MAGNODecoder, MAGNOEncoder


class GAOT3D(nn.Module):
    def __init__(self, input_size: int, output_size: int, magno_config: MAGNOConfig = None,
                 attn_config: TransformerConfig = None, latent_tokens: tuple = (32, 32, 32),
                 norm_domin: list = ((-1, -1, -1), (1, 1, 1))):
        super().__init__()
        magno_config = magno_config or MAGNOConfig()
        attn_config = attn_config or TransformerConfig()
        self.input_size, self.output_size = input_size, output_size
        self.node_latent_size = magno_config.lifting_channels
        self.patch_size = attn_config.patch_size
        self.D, self.H, self.W = latent_tokens
        self.num_latent_tokens = self.D * self.H * self.W
        self.coord_dim = magno_config.gno_coord_dim
        lo, hi = norm_domin
        axes = [torch.linspace(lo[a], hi[a], n) for a, n in enumerate((self.D, self.H, self.W))]
        grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, self.coord_dim)
        self.register_buffer("latent_tokens", grid)
        self.encoder = self.init_encoder(input_size, self.node_latent_size, magno_config)
        self.processor = self.init_processor(self.node_latent_size, attn_config)
        self.decoder = self.init_decoder(self.node_latent_size, output_size, magno_config)

    def init_encoder(self, input_size, node_latent_size, magno_config):
        return MAGNOEncoder(in_channels=input_size, out_channels=node_latent_size, gno_config=magno_config)

    def init_processor(self, node_latent_size, config):
        width = self.patch_size ** 3 * node_latent_size
        self.patch_linear = nn.Linear(width, width)
        self.positional_embedding_name = config.positional_embedding
        self.positions = self._get_patch_positions()
        for name in ("D", "H", "W"):            # the reference decorates the attention config the same way
            setattr(config.attn_config, name, getattr(self, name))
        return Transformer(input_size=width, output_size=width, config=config)

    def init_decoder(self, node_latent_size, output_size, magno_config):
        return MAGNODecoder(in_channels=node_latent_size, out_channels=output_size, gno_config=magno_config)

    def _get_patch_positions(self):
        P = self.patch_size
        ax = [torch.arange(n // P, dtype=torch.float32) for n in (self.D, self.H, self.W)]
        return torch.stack(torch.meshgrid(*ax, indexing="ij"), dim=-1).reshape(-1, 3)

    def _compute_absolute_embeddings(self, positions, embed_dim):
        """PE[:, 0::2] = sum_axes sin(pos * w_k), PE[:, 1::2] = sum_axes cos(pos * w_k) (gaot_3d.py:102-144)."""
        half = embed_dim // 2
        freq = 1 / 10000 ** (2 * torch.arange(0, half, dtype=torch.float32, device=positions.device) / embed_dim)
        ang = positions[:, :, None] * freq[None, None, :]
        pe = torch.zeros(positions.shape[0], embed_dim, device=positions.device)
        pe[:, 0::2] = torch.sin(ang).sum(dim=1)
        pe[:, 1::2] = torch.cos(ang).sum(dim=1)
        return pe

    def process(self, rndata: Optional[torch.Tensor] = None, condition: Optional[float] = None, slab=(0, 1)) -> torch.Tensor:
        """rndata [B, n, C] -> [B, n, C] (reference gaot_3d.py:166-222).  `slab = (r, R)` (extension, shard.py): `rndata`
        holds only the r-th of R equal D-slabs of the latent grid (a contiguous token range, and a contiguous patch
        range: both orders are D-major); the blocks then run in tblock's sequence-parallel mode."""
        from . import tgraph
        B, n, C = rndata.shape
        r, R = slab
        D, H, W, P = self.D, self.H, self.W, self.patch_size
        assert n * R == D * H * W, f"n_regional_nodes ({n * R}) is not equal to D*H*W ({D * H * W})"
        assert D % P == 0 and H % P == 0 and W % P == 0, "Dimensions must be divisible by patch size"
        assert (D // P) % R == 0, "latent slabs must hold whole patch planes"
        if condition is None:                       # the whole processor as two CUDA graphs (tgraph.py) when it fits
            y = tgraph.process(self, rndata, slab)
            if y is not None:
                return y
        nd, nh, nw = D // P // R, H // P, W // P
        x = rndata.view(B, nd, P, nh, P, nw, P, C).permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous()
        x = _lin(self.patch_linear, x.view(B, nd * nh * nw, P * P * P * C))
        S_l = nd * nh * nw
        pos = self.positions.to(x.device)
        relative_positions = None
        if self.positional_embedding_name == "absolute":
            x = x + self._compute_absolute_embeddings(pos, P * P * P * self.node_latent_size)[r * S_l:(r + 1) * S_l]
        elif self.positional_embedding_name == "rope":
            relative_positions = pos        # only a flag: RoPE is 1-D over the flattened patch index
        x = self.processor(x, condition=condition, relative_positions=relative_positions)
        x = x.view(B, nd, nh, nw, P, P, P, C).permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous()
        return x.view(B, n, C)

    def forward(self, batch, tokens_pos: Optional[torch.Tensor] = None, tokens_batch_idx: Optional[torch.Tensor] = None,
                query_coord_pos: Optional[torch.Tensor] = None, query_coord_batch_idx: Optional[torch.Tensor] = None,
                condition: Optional[float] = None) -> torch.Tensor:
        num_graphs = batch.num_graphs
        device = batch.pos.device
        if tokens_pos is None:
            assert tokens_batch_idx is None, "tokens_batch_idx should be None if tokens_pos is None"
            tokens_pos = self.latent_tokens
        if tokens_batch_idx is None:
            latent_pos = tokens_pos.to(device).repeat(num_graphs, 1)
            latent_batch = torch.arange(num_graphs, device=device).repeat_interleave(self.num_latent_tokens)
        else:
            assert tokens_pos.shape[0] == tokens_batch_idx.shape[0], "tokens_pos and tokens_batch_idx must have same length"
            assert tokens_batch_idx.max() == num_graphs - 1, "tokens_batch_idx does not match batch size"
            latent_pos, latent_batch = tokens_pos.to(device), tokens_batch_idx.to(device)
        if query_coord_pos is None:
            query_pos, query_batch = batch.pos, batch.batch
        else:
            assert query_coord_batch_idx is not None, "query_coord_batch_idx is required if query_coord_pos is provided"
            assert query_coord_pos.shape[0] == query_coord_batch_idx.shape[0], \
                "query_coord_pos and query_coord_batch_idx must have same length"
            assert query_coord_batch_idx.max() == num_graphs - 1, "query_coord_batch_idx does not match batch size"
            query_pos, query_batch = query_coord_pos.to(device), query_coord_batch_idx.to(device)
        from .graph import sample_scope
        with sample_scope():       # searches repeated inside this forward (reverse decoder, knn on both sides, scales) run once
            rndata = self.encoder(batch=batch, latent_tokens_pos=latent_pos, latent_tokens_batch_idx=latent_batch)
            rndata = self.process(rndata=rndata, condition=condition)
            flat = rndata.reshape(-1, self.node_latent_size)
            return self.decoder(rndata_flat=flat, phys_pos_query=query_pos, batch_idx_phys_query=query_batch,
                                latent_tokens_pos=latent_pos, latent_tokens_batch_idx=latent_batch, batch=batch)


def init_model(input_size: int, output_size: int, model: str, config=None):
    if model.lower() == "gaot_3d":
        return GAOT3D(input_size=input_size, output_size=output_size, magno_config=config.magno,
                      attn_config=config.transformer, latent_tokens=config.latent_tokens)
    raise ValueError(f"model {model} not supported currently!")


class Batch:
    """Duck-typed stand-in for torch_geometric.data.Batch: `.pos`, `.batch`, `.num_graphs` and feature
    attributes by name (all the reference's model code touches, magno.py:480-499)."""

    def __init__(self, pos, batch=None, num_graphs=1, **attrs):
        self.pos = pos
        self.batch = batch if batch is not None else torch.zeros(pos.shape[0], dtype=torch.long, device=pos.device)
        self.num_graphs = num_graphs
        for k, v in attrs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self
