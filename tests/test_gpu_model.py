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


@pytest.mark.parametrize("tag", ["radius_reverse", "knn", "bidirectional"])
def test_model_golden_fp32_tier(tag):
    """The strict tier of the north star (rtol 1e-5 in FP32): GNO / geometric embedding / lifting on this library's fp32
    kernels, the transformer through the reference's own fp32 library calls (set_transformer_precision('fp32')).  Truth is the
    oracle in fp64; the bar is elementwise |ours - truth| <= 1e-5 * max|truth| + what the reference's own fp32 evaluation
    (the golden output, computed by the unmodified reference modules on CPU) is away from that truth -- gradients likewise,
    per parameter tensor in relative L2: <= 1e-4 (sums over 1e3..1e4 points in a different order) + twice what the SAME
    arithmetic in fp32 on the CPU (the oracle in fp32) is away from the fp64 truth for that tensor (the q / k projection
    gradients cancel heavily: measured r02g 1.0e-4 .. 1.2e-4 there, everything else < 1e-4)."""
    import gaot_3d_b200 as G
    from tests.test_gpu_model_variants import _leaf_state
    g = torch.load(os.path.join(GOLD, "model_golden.pt"))[tag]
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    G.set_gno_precision("fp32")
    G.set_node_mlp_mode("torch")
    G.set_node_mlp_tf32(False)
    G.set_transformer_precision("fp32")
    try:
        m = build(g).train()
        batch = G.Batch(pos=g["pos"].to(DEV), c=g["c"].to(DEV))
        y = m(batch, tokens_pos=g["tokens_pos"].to(DEV))
        y.pow(2).mean().backward()
    finally:
        G.set_transformer_precision("bf16")
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    es, ds = G.parse_neighbor_strategy(g["strategy"])
    cfg = dict(latent_tokens=tuple(g["latent_tokens"]), patch_size=2, lifting_channels=32, radius=g["radius"], k=g["k"],
               enc_strategy=es, dec_strategy=ds, use_geoembed=g["use_geoembed"], num_layers=3, num_heads=4,
               num_kv_heads=m.processor.encoder_layers[0].attn.num_kv_heads, norm_eps=1e-6, positional_embedding="rope")
    sd = _leaf_state(m)
    y64 = omodel.gaot3d_forward(sd, cfg, keep_graph=True, dtype=torch.float64, pos=g["pos"], feats=[g["pos"], g["c"]],
                                latent_pos=g["tokens_pos"])
    y64.pow(2).mean().backward()
    sd32 = _leaf_state(m)
    omodel.gaot3d_forward(sd32, cfg, keep_graph=True, pos=g["pos"], feats=[g["pos"], g["c"]], latent_pos=g["tokens_pos"]).pow(2).mean().backward()
    scale = y64.detach().abs().max().item()
    ref_err = (g["out"].double() - y64.detach()).abs().max().item()
    err = (y.detach().cpu().double() - y64.detach()).abs().max().item()
    print(f"{tag}: ours-fp64 {err / scale:.2e}  reference(fp32 CPU)-fp64 {ref_err / scale:.2e}")
    assert err <= 1e-5 * scale + 2.0 * ref_err, f"{tag}: |ours - fp64| {err:.3e}, reference's own {ref_err:.3e}, scale {scale:.3e}"
    bad = []
    for n, p in m.named_parameters():
        if not p.requires_grad:
            continue
        gr = sd[n].grad.double()
        rel = ((p.grad.detach().cpu().double() - gr).norm() / gr.norm().clamp(min=1e-30)).item()
        own = ((sd32[n].grad.double() - gr).norm() / gr.norm().clamp(min=1e-30)).item()
        if rel > 1e-4 + 2.0 * own:
            bad.append(f"{n}: {rel:.2e} (fp32 CPU evaluation: {own:.2e})")
    assert not bad, f"{tag} fp32-tier gradient parity: " + "; ".join(bad)
