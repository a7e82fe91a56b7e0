"""Whole-model GPU parity over the option surface and the precision tiers (VERDICT r01 "what's weak"):

* the three golden models and four option-surface variants (multi-scale + learned scale weights, mlp_type='channel' +
  multi-scale sum, 'absolute' PE, a batch of two examples), forward AND parameter gradients, in BOTH tiers:
  'default' (fp32 GNO kernels, torch node MLPs) and 'bench' (exactly what bench.py times: set_gno_precision("bf16") +
  set_node_mlp_mode("fused"));
* 'ratio' edge sampling through the model with the mask made explicit (the edges the model actually used are captured
  and handed to the oracle), precomputed int32 edges in arbitrary order through MAGNOEncoder / MAGNODecoder.

Oracle: oracle.model.gaot3d_forward (pinned against the reference's GAOT3D for these very variants in
tests/test_oracle_vs_reference.py).  Bars: output max-abs error <= 2e-2 * max|ref| (north star's BF16 tier; the
transformer is bf16 in both tiers).  Gradients: relative L2 <= 2e-2 per parameter tensor against the fp32 oracle or
the bf16-operand YARDSTICK (the same model in fp64 with every tensor-core operand rounded where an ideal bf16
implementation rounds it); the distance yardstick <-> fp32 oracle is printed next to it -- that is what bf16 costs.
Only where that cost itself exceeds 2e-2 (q/k projection gradients of some layers) the bar becomes "no farther from
fp32 than twice the ideal bf16 evaluation".
"""
import os

import pytest
import torch

from oracle import model as omodel
from tests import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"

VARIANTS = {
    # tag: (strategy, geoembed, mlp_type, scales, use_scale_weights, positional_embedding, num_layers, num_graphs)
    "multiscale_weighted": (["radius", "radius"], [True, True], "linear", [1.0, 1.5], True, "rope", 2, 1),
    "multiscale_sum_channel": ("bidirectional", [True, False], "channel", [0.8, 1.3], False, "rope", 3, 1),
    "absolute_pe": ("knn", [False, False], "linear", [1.0], False, "absolute", 2, 1),
    "batched": (["radius", "reverse"], [True, False], "linear", [1.0], False, "rope", 2, 2),
}
GRID = (8, 6, 4)
R, K, C = 0.4, 2, 32


@pytest.fixture(params=["default", "bench"])
def tier(request):
    import gaot_3d_b200 as G
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    if request.param == "bench":
        G.set_gno_precision("bf16")
        G.set_node_mlp_mode("fused")
    else:
        G.set_gno_precision("fp32")
        G.set_node_mlp_mode("torch")
    yield request.param
    G.set_gno_precision("fp32")
    G.set_node_mlp_mode("torch")
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def _inputs(num_graphs):
    n_per = [1500, 1100][:num_graphs]
    pos = torch.cat([torch.from_numpy(synth.surface_cloud(n, seed=21 + i)) for i, n in enumerate(n_per)])
    c = torch.from_numpy(synth.unit_normals(pos.shape[0], seed=9))
    bidx = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(n_per)])
    lat = torch.from_numpy(synth.latent_grid(GRID))
    return pos, c, bidx, lat


def _build(strat, geo, mlp, scales, usw, pe, nl, **extra):
    import gaot_3d_b200 as G
    mc = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=C, neighbor_strategy=strat, gno_radius=R, mlp_type=mlp,
                       precompute_edges=False, use_geoembed=geo, encoder_feature_attr=["pos", "c"], k_neighbors=K,
                       scales=scales, use_scale_weights=usw, **extra)
    tc = G.TransformerConfig(patch_size=2, hidden_size=128, num_layers=nl, positional_embedding=pe)
    tc.attn_config.hidden_size, tc.attn_config.num_heads, tc.attn_config.num_kv_heads = 128, 4, 4
    tc.attn_config.atten_dropout = 0.0
    tc.ffn_config.hidden_size = 128
    torch.manual_seed(11)
    return G.GAOT3D(6, 4, mc, tc, latent_tokens=GRID).to(DEV)


def _cfg(strat, geo, scales, usw, pe, nl, nkv=4):
    es, ds = (strat, strat) if isinstance(strat, str) else strat
    return dict(latent_tokens=GRID, patch_size=2, lifting_channels=C, radius=R, k=K, enc_strategy=es, dec_strategy=ds,
                use_geoembed=geo, num_layers=nl, num_heads=4, num_kv_heads=nkv, norm_eps=1e-6, positional_embedding=pe,
                scales=scales, use_scale_weights=usw)


def _leaf_state(m):
    return {n: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and "freqs" not in n and n != "latent_tokens")
            for n, v in m.state_dict().items()}


def check_against_oracle(m, run_model, oracle_kwargs, cfg, label, out_tol=2e-2, grad_tol=2e-2):
    """forward + gradients of `m` (loss = mean(y^2)) against the fp32 oracle (output) and the bf16 yardstick (gradients)."""
    m.train()
    y = run_model()
    y.pow(2).mean().backward()
    torch.cuda.synchronize()
    missing = [n for n, p in m.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing
    sd32 = _leaf_state(m)
    y32 = omodel.gaot3d_forward(sd32, cfg, keep_graph=True, **oracle_kwargs)
    y32.pow(2).mean().backward()
    err = (y.detach().cpu() - y32.detach()).abs().max().item()
    scale = y32.detach().abs().max().item()
    assert err <= out_tol * scale, f"{label}: output max abs err {err:.3e} vs max|ref| {scale:.3e}"
    sdy = _leaf_state(m)
    yy = omodel.gaot3d_forward(sdy, cfg, keep_graph=True, dtype=torch.float64, emulate_bf16=True, **oracle_kwargs)
    yy.pow(2).mean().backward()
    bad, report = [], []
    for n, p in m.named_parameters():
        if not p.requires_grad:
            continue
        g, g32, gy = p.grad.detach().cpu().double(), sd32[n].grad.double(), sdy[n].grad.double()
        rel_y = ((g - gy).norm() / gy.norm().clamp(min=1e-30)).item()
        rel_32 = ((g - g32).norm() / g32.norm().clamp(min=1e-30)).item()
        cost = ((gy - g32).norm() / g32.norm().clamp(min=1e-30)).item()
        report.append(f"{n}: ours-yardstick {rel_y:.2e}  ours-fp32 {rel_32:.2e}  yardstick-fp32 {cost:.2e}")
        # bar: rtol 2e-2 against the fp32 oracle or the yardstick; where bf16 ITSELF moves the fp32 gradient by more than
        # that (q/k projections: dS = P (dP - D) cancels), ours may be no farther from fp32 than twice what the ideal
        # bf16-operand evaluation is (measured r02: yardstick-fp32 1.6e-2..2.9e-2 there, ours-fp32 2.0e-2..2.8e-2)
        if min(rel_y, rel_32) > grad_tol and rel_32 > 2.0 * cost:
            bad.append(report[-1])
    print(f"--- {label}\n" + "\n".join(report))
    assert not bad, f"{label} gradient parity: " + "; ".join(bad)
    return err / scale


@pytest.mark.parametrize("tag", list(VARIANTS))
def test_model_variants(tag, tier):
    import gaot_3d_b200 as G
    strat, geo, mlp, scales, usw, pe, nl, B = VARIANTS[tag]
    pos, c, bidx, lat = _inputs(B)
    m = _build(strat, geo, mlp, scales, usw, pe, nl)
    batch = G.Batch(pos=pos.to(DEV), batch=bidx.to(DEV), num_graphs=B, c=c.to(DEV))
    check_against_oracle(m, lambda: m(batch, tokens_pos=lat.to(DEV)),
                         dict(pos=pos, feats=[pos, c], latent_pos=lat, batch_idx=bidx if B > 1 else None, num_graphs=B),
                         _cfg(strat, geo, scales, usw, pe, nl), f"{tag}/{tier}")


@pytest.mark.parametrize("tag", ["radius_reverse", "knn", "bidirectional"])
def test_golden_models_in_both_tiers(tag, tier):
    """The reference's own golden models (fixtures written by the reference GAOT3D) in the exact bench configuration as
    well as the default one: output against the golden output, gradients against the yardstick."""
    import gaot_3d_b200 as G
    from tests.test_gpu_model import build
    g = torch.load(os.path.join(GOLD, "model_golden.pt"))[tag]
    m = build(g)
    batch = G.Batch(pos=g["pos"].to(DEV), c=g["c"].to(DEV))
    m.eval()
    with torch.no_grad():
        y = m(batch, tokens_pos=g["tokens_pos"].to(DEV))
    err = (y.cpu() - g["out"]).abs().max().item()
    assert err <= 2e-2 * g["out"].abs().max().item(), f"{tag}/{tier}: {err:.3e}"
    strat = g["strategy"]
    nkv = m.processor.encoder_layers[0].attn.num_kv_heads
    cfg = _cfg(strat, g["use_geoembed"], [1.0], False, "rope", 3, nkv)
    cfg.update(latent_tokens=tuple(g["latent_tokens"]), radius=g["radius"], k=g["k"])
    check_against_oracle(m, lambda: m(batch, tokens_pos=g["tokens_pos"].to(DEV)),
                         dict(pos=g["pos"], feats=[g["pos"], g["c"]], latent_pos=g["tokens_pos"]), cfg, f"golden {tag}/{tier}")


def test_ratio_sampling_through_the_model_with_explicit_mask(tier):
    """sampling_strategy='ratio' (reference magno.py:360-368): the Philox mask cannot match torch.rand bit for bit, so the
    edges the model actually used are captured and the oracle runs on exactly those; the mask itself is checked for what
    dropout_edge guarantees (order-preserving subset of the full graph, keep rate ~ sample_ratio, eval() = passthrough)."""
    import gaot_3d_b200 as G
    from gaot_3d_b200 import graph as gg
    strat, geo = ["radius", "radius"], [True, False]
    pos, c, _, lat = _inputs(1)
    m = _build(strat, geo, "linear", [1.0], False, "rope", 2, sampling_strategy="ratio", sample_ratio=0.6)
    used = {}
    for name, mod in (("enc", m.encoder), ("dec", m.decoder)):
        orig = mod._sampling

        def wrapped(ei, nq, device, _orig=orig, _name=name):
            out = _orig(ei, nq, device)
            used[_name] = (ei.detach().clone(), out.detach().clone())
            return out
        mod._sampling = wrapped
    batch = G.Batch(pos=pos.to(DEV), c=c.to(DEV))
    cfg = _cfg(strat, geo, [1.0], False, "rope", 2)

    def run():
        return m(batch, tokens_pos=lat.to(DEV))

    m.train()
    run()
    first = used["enc"][1].cpu()
    for name in ("enc", "dec"):
        full, kept = (t.cpu() for t in used[name])
        E, Ek = full.shape[1], kept.shape[1]
        assert 0.5 * E < Ek < 0.7 * E, f"{name}: kept {Ek} of {E} at ratio 0.6"
        mult = int(full.max()) + 1
        idx = {int(v): i for i, v in enumerate((full[0] * mult + full[1]).tolist())}
        kk = (kept[0] * mult + kept[1]).tolist()
        assert all(int(v) in idx for v in kk), f"{name}: kept edges are not a subset of the graph"
        seq = [idx[int(v)] for v in kk]
        assert all(a_ < b_ for a_, b_ in zip(seq, seq[1:])), f"{name}: edge order not preserved by the mask"
    run()                                          # a second training forward draws a different mask
    assert not torch.equal(used["enc"][1].cpu(), first), "two training forwards drew the same edge mask"
    enc_e, dec_e = used["enc"][1].cpu(), used["dec"][1].cpu()
    # oracle on exactly the edges of the LAST forward; re-run the model with the sampling pinned to those edges
    m.encoder._sampling = lambda ei, nq, device: used["enc"][1]
    m.decoder._sampling = lambda ei, nq, device: used["dec"][1]
    check_against_oracle(m, run, dict(pos=pos, feats=[pos, c], latent_pos=lat, enc_edges=[enc_e], dec_edges=[dec_e]), cfg,
                         f"ratio/{tier}")
    # eval(): passthrough (dropout_edge(training=False))
    m.eval()
    captured = {}
    m.encoder._sampling = lambda ei, nq, device: captured.setdefault("e", gg.apply_neighbor_sampling(
        ei, nq, device, "ratio", None, 0.6, False))
    with torch.no_grad():
        m(batch, tokens_pos=lat.to(DEV))
    assert captured["e"].shape[1] == used["enc"][0].shape[1]


def test_precomputed_int32_edges_any_order(tier):
    """precompute_edges=True (reference magno.py:506-516, :715-725; int32 on disk, stat.py:191,208): edges handed over on
    the batch as int32 in ARBITRARY order must give the result of the same graph built online."""
    import gaot_3d_b200 as G
    from oracle import graph as og
    pos, c, _, lat = _inputs(1)
    m = _build("radius", [True, False], "linear", [1.0], False, "rope", 2)
    m.encoder.precompute_edges = m.decoder.precompute_edges = True
    ee = torch.from_numpy(og.get_neighbor_strategy_np("bidirectional", pos.numpy(), None, lat.numpy(), None, R, K, False))
    de = torch.from_numpy(og.get_neighbor_strategy_np("radius", pos.numpy(), None, lat.numpy(), None, R, K, True))
    gen = torch.Generator().manual_seed(5)
    ee = ee[:, torch.randperm(ee.shape[1], generator=gen)].to(torch.int32)
    de = de[:, torch.randperm(de.shape[1], generator=gen)].to(torch.int32)
    batch = G.Batch(pos=pos.to(DEV), c=c.to(DEV), encoder_edge_index_s0=ee.to(DEV), decoder_edge_index_s0=de.to(DEV))
    cfg = _cfg("radius", [True, False], [1.0], False, "rope", 2)
    check_against_oracle(m, lambda: m(batch, tokens_pos=lat.to(DEV)),
                         dict(pos=pos, feats=[pos, c], latent_pos=lat, enc_edges=[ee], dec_edges=[de]), cfg, f"precomputed/{tier}")
    with pytest.raises(AttributeError):
        m(G.Batch(pos=pos.to(DEV), c=c.to(DEV)), tokens_pos=lat.to(DEV))


def test_bad_precomputed_edges_raise():
    """an out-of-range index in a precomputed edge list must raise, not read or write out of bounds
    (the reference's torch indexing raises an IndexError there)."""
    import gaot_3d_b200 as G
    pos, c, _, lat = _inputs(1)
    m = _build("radius", [False, False], "linear", [1.0], False, "rope", 2)
    m.encoder.precompute_edges = m.decoder.precompute_edges = True
    n, M = pos.shape[0], lat.shape[0]
    good_e = torch.stack([torch.randint(0, n, (4000,)), torch.randint(0, M, (4000,))]).to(torch.int32)
    good_d = good_e.flip(0).contiguous()
    for bad_row, bad_val in ((0, n), (1, M), (0, -1)):
        bad = good_e.clone()
        bad[bad_row, 17] = bad_val
        batch = G.Batch(pos=pos.to(DEV), c=c.to(DEV), encoder_edge_index_s0=bad.to(DEV), decoder_edge_index_s0=good_d.to(DEV))
        with pytest.raises((ValueError, IndexError)):
            m(batch, tokens_pos=lat.to(DEV))
            torch.cuda.synchronize()
