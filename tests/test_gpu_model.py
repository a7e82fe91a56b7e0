"""GPU parity of the whole drop-in GAOT3D against the reference's own output (golden) and the
travelling oracle: strict state_dict load, identical edge sets, output within the BF16-attention tolerance."""
import os

import pytest
import torch

from oracle import model as omodel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def build(g):
    import gaot_3d_b200 as G
    mc = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=32, neighbor_strategy=g["strategy"], gno_radius=g["radius"],
                       mlp_type="linear", precompute_edges=False, use_geoembed=g["use_geoembed"],
                       encoder_feature_attr=["pos", "c"], k_neighbors=g["k"])
    tc = G.TransformerConfig(patch_size=2, hidden_size=128, num_layers=3, positional_embedding="rope")
    tc.attn_config.hidden_size = 128
    tc.attn_config.num_heads = 4
    tc.attn_config.num_kv_heads = g["state"]["processor.encoder_layers.0.attn.k_proj.weight"].shape[0] // 32
    tc.attn_config.atten_dropout = 0.0
    tc.ffn_config.hidden_size = 128
    m = G.GAOT3D(6, 4, mc, tc, latent_tokens=tuple(g["latent_tokens"]))
    m.load_state_dict({k: v.float() for k, v in g["state"].items()}, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("tag", ["radius_reverse", "knn", "bidirectional"])
def test_model_golden(tag):
    import gaot_3d_b200 as G
    g = torch.load(os.path.join(GOLD, "model_golden.pt"))[tag]
    m = build(g).eval()
    batch = G.Batch(pos=g["pos"].to(DEV), c=g["c"].to(DEV))
    with torch.no_grad():
        y = m(batch, tokens_pos=g["tokens_pos"].to(DEV))
    ref = g["out"]
    err = (y.cpu() - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item(), f"{tag}: max abs err {err:.3e} vs max |ref| {ref.abs().max().item():.3e}"
    # training step: gradients flow to every trainable parameter; each tensor is held to rtol 2e-2 (relative L2) against
    # the bf16-operand yardstick of oracle.model (q/k projections included -- no widened bars), see
    # tests/test_gpu_model_variants.py::check_against_oracle
    from tests.test_gpu_model_variants import check_against_oracle
    es, ds = G.parse_neighbor_strategy(g["strategy"])
    cfg = dict(latent_tokens=tuple(g["latent_tokens"]), patch_size=2, lifting_channels=32, radius=g["radius"], k=g["k"],
               enc_strategy=es, dec_strategy=ds, use_geoembed=g["use_geoembed"], num_layers=3, num_heads=4,
               num_kv_heads=m.processor.encoder_layers[0].attn.num_kv_heads, norm_eps=1e-6, positional_embedding="rope")
    check_against_oracle(m, lambda: m(batch, tokens_pos=g["tokens_pos"].to(DEV)),
                         dict(pos=g["pos"], feats=[g["pos"], g["c"]], latent_pos=g["tokens_pos"]), cfg, f"golden {tag}")
