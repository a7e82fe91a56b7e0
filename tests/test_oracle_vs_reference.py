"""Pins the travelling torch restatements (oracle/gno.py, oracle/attn.py, oracle/model.py) against
(a) the reference's own modules imported from /root/reference (dev container only) and
(b) the committed golden vectors those modules produced (everywhere).  Pure CPU."""
import os

import pytest
import torch

from oracle import attn as oattn, gno as ogno, model as omodel, ref_loader

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


def test_gno_golden():
    for tag, g in torch.load(os.path.join(GOLD, "gno_golden.pt")).items():
        ws = [w.clone().requires_grad_(True) for w in g["weights"]]
        bs = [b.clone().requires_grad_(True) for b in g["biases"]]
        f = g["f_y"].clone().requires_grad_(True)
        out = ogno.integral_transform(g["y_pos"], g["x_pos"], g["edge_index"], f, ws, bs)
        assert torch.allclose(out, g["out"], rtol=1e-5, atol=1e-7), tag
        out.backward(g["d_out"])
        assert torch.allclose(f.grad, g["d_f"], rtol=1e-4, atol=1e-6)
        for a, b in zip(ws, g["d_weights"]):
            assert torch.allclose(a.grad, b, rtol=1e-3, atol=1e-5 * b.abs().max().item())


def test_geo_golden():
    g = torch.load(os.path.join(GOLD, "geo_golden.pt"))
    feats = ogno.geo_statistical_features(g["source_pos"], g["query_pos"], g["edge_index"])
    assert torch.allclose(feats, g["features"], rtol=1e-4, atol=1e-5)
    s = g["state"]
    emb = ogno.geo_embedding(g["source_pos"], g["query_pos"], g["edge_index"], s["mlp.0.weight"], s["mlp.0.bias"],
                             s["mlp.2.weight"], s["mlp.2.bias"])
    assert torch.allclose(emb, g["embedding"], rtol=1e-4, atol=1e-5)


def test_attention_golden():
    for tag, g in torch.load(os.path.join(GOLD, "attn_golden.pt")).items():
        s = g["state"]
        x = g["x"]
        lin = torch.nn.functional.linear
        o = oattn.attention_core(lin(x, s["q_proj.weight"]), lin(x, s["k_proj.weight"]), lin(x, s["v_proj.weight"]),
                                 g["num_heads"], g["num_kv_heads"], g["rope"])
        assert torch.allclose(lin(o, s["o_proj.weight"]), g["out"], rtol=1e-4, atol=1e-6), tag


def test_model_golden():
    for tag, g in torch.load(os.path.join(GOLD, "model_golden.pt")).items():
        strat = g["strategy"]
        es, ds = (strat, strat) if isinstance(strat, str) else strat
        nkv = g["state"]["processor.encoder_layers.0.attn.k_proj.weight"].shape[0] // 32
        cfg = dict(latent_tokens=tuple(g["latent_tokens"]), patch_size=2, lifting_channels=32, radius=g["radius"], k=g["k"],
                   enc_strategy=es, dec_strategy=ds, use_geoembed=g["use_geoembed"], num_layers=3, num_heads=4,
                   num_kv_heads=nkv, norm_eps=1e-6, positional_embedding="rope")
        y = omodel.gaot3d_forward(g["state"], cfg, g["pos"], [g["pos"], g["c"]], latent_pos=g["tokens_pos"])
        assert torch.allclose(y, g["out"], rtol=1e-4, atol=1e-6), tag


@needs_ref
def test_restatements_match_reference_modules():
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    y, x = torch.rand(800, 3) * 2 - 1, torch.rand(200, 3) * 2 - 1
    ei = torch.stack([torch.randint(0, 800, (5000,)), torch.randint(0, 200, (5000,))])
    for tt in ("linear", "nonlinear", "nonlinear_kernelonly"):
        kin = 6 + (16 if tt != "linear" else 0)
        it = ref.integral_transform.IntegralTransform(channel_mlp_layers=[kin, 32, 16], transform_type=tt)
        f = torch.randn(800, 16)
        a = it(y, x, ei, f)
        b = ogno.integral_transform(y, x, ei, f, [l.weight for l in it.channel_mlp.fcs], [l.bias for l in it.channel_mlp.fcs], tt)
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), tt
    # attentional integral transform (use_attn), both score types
    for at in ("cosine", "dot_product"):
        it = ref.integral_transform.IntegralTransform(channel_mlp_layers=[6, 32, 16], use_attn=True, coord_dim=3, attention_type=at)
        f = torch.randn(800, 16)
        a = it(y, x, ei, f)
        attn = dict(type=at, coord_dim=3)
        if at == "dot_product":
            attn.update(wq=it.query_proj.weight, bq=it.query_proj.bias, wk=it.key_proj.weight, bk=it.key_proj.bias)
        b = ogno.integral_transform(y, x, ei, f, [l.weight for l in it.channel_mlp.fcs], [l.bias for l in it.channel_mlp.fcs], attn=attn)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), at
    ge = ref.geoembed.GeometricEmbedding(3, 8)
    a = ge._compute_statistical_features_pyg(y, x, ei)
    assert torch.allclose(a, ogno.geo_statistical_features(y, x, ei), rtol=1e-5, atol=1e-6)
    # PointNet embedding, both poolings (queries without edges exist: 200 queries, some indices never drawn at E = 300)
    ei2 = ei[:, :300]
    for pooling in ("max", "mean"):
        gp = ref.geoembed.GeometricEmbedding(3, 8, method="pointnet", pooling=pooling)
        st = gp.state_dict()
        a = gp(y, x, ei2)
        b = ogno.geo_pointnet_embedding(y, x, ei2, st["pointnet_mlp.0.weight"], st["pointnet_mlp.0.bias"], st["pointnet_mlp.2.weight"],
                                        st["pointnet_mlp.2.bias"], st["fc.0.weight"], st["fc.0.bias"], pooling)
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), pooling
        assert float(b[torch.bincount(ei2[1], minlength=200) == 0].abs().max()) == 0.0
    # 2-D coordinates (MAGNOConfig.gno_coord_dim defaults to 2, reference magno.py:23): the restatements are dimension-generic
    y2, x2 = y[:, :2].contiguous(), x[:, :2].contiguous()
    it = ref.integral_transform.IntegralTransform(channel_mlp_layers=[4, 32, 16])
    f = torch.randn(800, 16)
    b = ogno.integral_transform(y2, x2, ei, f, [l.weight for l in it.channel_mlp.fcs], [l.bias for l in it.channel_mlp.fcs])
    assert torch.allclose(it(y2, x2, ei, f), b, rtol=1e-6, atol=1e-7), "2-D integral transform"
    ge2 = ref.geoembed.GeometricEmbedding(2, 8)
    a = ge2._compute_statistical_features_pyg(y2, x2, ei)
    assert a.shape[1] == 7 and torch.allclose(a, ogno.geo_statistical_features(y2, x2, ei), rtol=1e-5, atol=1e-6), "2-D statistics"
    gp2 = ref.geoembed.GeometricEmbedding(2, 8, method="pointnet")
    st = gp2.state_dict()
    b = ogno.geo_pointnet_embedding(y2, x2, ei2, st["pointnet_mlp.0.weight"], st["pointnet_mlp.0.bias"], st["pointnet_mlp.2.weight"],
                                    st["pointnet_mlp.2.bias"], st["fc.0.weight"], st["fc.0.bias"], "max")
    assert torch.allclose(gp2(y2, x2, ei2), b, rtol=1e-6, atol=1e-7), "2-D pointnet"
    # scatter_native mean with empty segments
    src, idx = torch.randn(50, 4), torch.randint(0, 9, (50,))
    assert torch.allclose(ref.scatter_native.scatter_native(src, idx, dim=0, dim_size=12, reduce="mean"), ogno.scatter_mean(src, idx, 12))


@needs_ref
def test_config_schema_and_state_dict_keys_match_reference():
    """Drop-in boundary: dataclass fields/defaults and parameter names are the schema."""
    import dataclasses
    import gaot_3d_b200 as G
    ref = ref_loader.load_reference()
    for ours, theirs in ((G.MAGNOConfig, ref.magno.MAGNOConfig), (G.TransformerConfig, ref.attn.TransformerConfig),
                         (G.AttentionConfig, ref.attn.AttentionConfig), (G.FFNConfig, ref.attn.FFNConfig)):
        fo = {f.name: f for f in dataclasses.fields(ours)}
        ft = {f.name: f for f in dataclasses.fields(theirs)}
        assert list(fo) == list(ft), ours.__name__
        a, b = ours(), theirs()
        for n in fo:
            va, vb = getattr(a, n), getattr(b, n)
            if dataclasses.is_dataclass(va):
                continue
            assert va == vb, (ours.__name__, n, va, vb)
    for strat, geo, mlp in ((["radius", "reverse"], [True, False], "linear"), ("bidirectional", [True, True], "channel")):
        kw = dict(gno_coord_dim=3, lifting_channels=16, neighbor_strategy=strat, use_geoembed=geo, mlp_type=mlp,
                  encoder_feature_attr=["pos", "c"], use_scale_weights=True, scales=[1.0, 2.0])
        for pe, nl in (("rope", 3), ("absolute", 4)):
            tr, to = ref.attn.TransformerConfig(patch_size=2, hidden_size=64, num_layers=nl, positional_embedding=pe), \
                     G.TransformerConfig(patch_size=2, hidden_size=64, num_layers=nl, positional_embedding=pe)
            for t in (tr, to):
                t.attn_config.hidden_size, t.ffn_config.hidden_size = 64, 96
            mr = ref.gaot_3d.GAOT3D(6, 4, ref.magno.MAGNOConfig(**kw), tr, latent_tokens=(4, 4, 4))
            mo = G.GAOT3D(6, 4, G.MAGNOConfig(**kw), to, latent_tokens=(4, 4, 4))
            sr, so = mr.state_dict(), mo.state_dict()
            assert list(sr) == list(so)
            assert all(sr[k].shape == so[k].shape for k in sr)
            mo.load_state_dict(sr, strict=True)


def test_pointnet_golden_matches_oracle():
    """The committed PointNet fixture (made by the reference module) against the CPU restatement: runs without /root/reference."""
    g = torch.load(os.path.join(GOLD, "pointnet_golden.pt"))
    for pooling, d in g.items():
        st = d["state"]
        out = ogno.geo_pointnet_embedding(d["source_pos"], d["query_pos"], d["edge_index"], st["pointnet_mlp.0.weight"],
                                          st["pointnet_mlp.0.bias"], st["pointnet_mlp.2.weight"], st["pointnet_mlp.2.bias"],
                                          st["fc.0.weight"], st["fc.0.bias"], pooling)
        assert torch.allclose(out, d["out"], rtol=1e-6, atol=1e-7), pooling


VARIANTS = {
    # tag: (strategy, geoembed, mlp_type, scales, use_scale_weights, positional_embedding, num_layers, num_graphs)
    "multiscale_weighted": (["radius", "radius"], [True, True], "linear", [1.0, 1.5], True, "rope", 2, 1),
    "multiscale_sum_channel": ("bidirectional", [True, False], "channel", [0.8, 1.3], False, "rope", 3, 1),
    "absolute_pe": ("knn", [False, False], "linear", [1.0], False, "absolute", 2, 1),
    "batched": (["radius", "reverse"], [True, False], "linear", [1.0], False, "rope", 2, 2),
}


def _variant_inputs(num_graphs):
    from tests import synth
    G = (6, 4, 4)
    n_per = [700, 500][:num_graphs]
    pos = torch.cat([torch.from_numpy(synth.surface_cloud(n, seed=21 + i)) for i, n in enumerate(n_per)])
    c = torch.from_numpy(synth.unit_normals(pos.shape[0], seed=9))
    bidx = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(n_per)])
    lat = torch.from_numpy(synth.latent_grid(G))
    return G, pos, c, bidx, lat


@needs_ref
@pytest.mark.parametrize("tag", list(VARIANTS))
def test_model_restatement_variants_match_reference(tag):
    """oracle.model.gaot3d_forward against the reference's own GAOT3D for the option surface the golden models do not
    reach: multi-scale (+ learned scale weights), mlp_type='channel', absolute PE, a batch of two examples --
    output and parameter gradients (reference magno.py:502-596, :711-798; gaot_3d.py:102-144, :278-290)."""
    ref = ref_loader.load_reference()
    strat, geo, mlp, scales, usw, pe, nl, B = VARIANTS[tag]
    torch.manual_seed(7)
    G, pos, c, bidx, lat = _variant_inputs(B)
    r, k = 0.4, 2
    MC = ref.magno.MAGNOConfig(gno_coord_dim=3, lifting_channels=16, neighbor_strategy=strat, gno_radius=r, mlp_type=mlp,
                               precompute_edges=False, use_geoembed=geo, encoder_feature_attr=["pos", "c"], k_neighbors=k,
                               scales=scales, use_scale_weights=usw)
    TC = ref.attn.TransformerConfig(patch_size=2, hidden_size=128, num_layers=nl, positional_embedding=pe)
    TC.attn_config.hidden_size, TC.attn_config.num_heads, TC.attn_config.num_kv_heads = 128, 4, 4
    TC.ffn_config.hidden_size = 128
    m = ref.gaot_3d.GAOT3D(6, 4, MC, TC, latent_tokens=G).eval()
    y = m(ref_loader.SimpleBatch(pos=pos, batch=bidx, num_graphs=B, c=c), tokens_pos=lat)
    y.pow(2).mean().backward()
    es, ds = (strat, strat) if isinstance(strat, str) else strat
    cfg = dict(latent_tokens=G, patch_size=2, lifting_channels=16, radius=r, k=k, enc_strategy=es, dec_strategy=ds,
               use_geoembed=geo, num_layers=nl, num_heads=4, num_kv_heads=4, norm_eps=1e-6, positional_embedding=pe,
               scales=scales, use_scale_weights=usw)
    sd = {n: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "freqs" not in n and n != "latent_tokens")
          for n, v in m.state_dict().items()}
    yo = omodel.gaot3d_forward(sd, cfg, pos, [pos, c], latent_pos=lat, keep_graph=True,
                               batch_idx=bidx if B > 1 else None, num_graphs=B)
    assert torch.allclose(yo, y, rtol=1e-4, atol=1e-6), tag
    yo.pow(2).mean().backward()
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        assert sd[n].grad is not None, n
        rel = ((sd[n].grad - p.grad).norm() / p.grad.norm().clamp(min=1e-30)).item()
        assert rel < 1e-3, (tag, n, rel)


@needs_ref
def test_model_restatement_explicit_edges_match_reference():
    """Precomputed-edge path (magno.py:506-516, :715-725): int32 edges in arbitrary order handed over on the batch."""
    ref = ref_loader.load_reference()
    torch.manual_seed(3)
    G, pos, c, bidx, lat = _variant_inputs(1)
    r, k = 0.4, 2
    MC = ref.magno.MAGNOConfig(gno_coord_dim=3, lifting_channels=16, neighbor_strategy="radius", gno_radius=r, mlp_type="linear",
                               precompute_edges=True, use_geoembed=[True, False], encoder_feature_attr=["pos", "c"])
    TC = ref.attn.TransformerConfig(patch_size=2, hidden_size=128, num_layers=2, positional_embedding="rope")
    TC.attn_config.hidden_size, TC.attn_config.num_heads, TC.attn_config.num_kv_heads = 128, 4, 4
    TC.ffn_config.hidden_size = 128
    m = ref.gaot_3d.GAOT3D(6, 4, MC, TC, latent_tokens=G).eval()
    from oracle import graph as og
    ee = torch.from_numpy(og.get_neighbor_strategy_np("bidirectional", pos.numpy(), None, lat.numpy(), None, r, k, False))
    de = torch.from_numpy(og.get_neighbor_strategy_np("radius", pos.numpy(), None, lat.numpy(), None, r, k, True))
    ee = ee[:, torch.randperm(ee.shape[1])].to(torch.int32)
    de = de[:, torch.randperm(de.shape[1])].to(torch.int32)
    y = m(ref_loader.SimpleBatch(pos=pos, c=c, encoder_edge_index_s0=ee, decoder_edge_index_s0=de), tokens_pos=lat)
    cfg = dict(latent_tokens=G, patch_size=2, lifting_channels=16, radius=r, k=k, enc_strategy="radius", dec_strategy="radius",
               use_geoembed=[True, False], num_layers=2, num_heads=4, num_kv_heads=4, norm_eps=1e-6, positional_embedding="rope")
    yo = omodel.gaot3d_forward(m.state_dict(), cfg, pos, [pos, c], latent_pos=lat, enc_edges=[ee], dec_edges=[de])
    assert torch.allclose(yo, y, rtol=1e-4, atol=1e-6)


def test_bf16_yardstick_is_close_to_fp32_oracle():
    """The bf16-operand yardstick (oracle.model emulate_bf16) is the same function up to operand rounding: its output is
    within the mixed-precision tier of the fp32 restatement on a golden model."""
    g = torch.load(os.path.join(GOLD, "model_golden.pt"))["knn"]
    cfg = dict(latent_tokens=tuple(g["latent_tokens"]), patch_size=2, lifting_channels=32, radius=g["radius"], k=g["k"],
               enc_strategy="knn", dec_strategy="knn", use_geoembed=g["use_geoembed"], num_layers=3, num_heads=4,
               num_kv_heads=2, norm_eps=1e-6, positional_embedding="rope")
    y64 = omodel.gaot3d_forward(g["state"], cfg, g["pos"], [g["pos"], g["c"]], latent_pos=g["tokens_pos"], dtype=torch.float64)
    assert torch.allclose(y64.float(), g["out"], rtol=1e-4, atol=1e-6)
    yb = omodel.gaot3d_forward(g["state"], cfg, g["pos"], [g["pos"], g["c"]], latent_pos=g["tokens_pos"], dtype=torch.float64,
                               emulate_bf16=True)
    err = (yb - y64).abs().max().item() / y64.abs().max().item()
    assert 0 < err < 2e-2, err
