"""GPU parity: fused GNO forward/backward and geometric-embedding kernels vs the CPU oracle and the
golden vectors produced by the reference's own IntegralTransform / GeometricEmbedding.
Tolerance (north star): rtol 1e-5 in FP32 (+ atol 1e-6*max|ref| for near-zero entries)."""
import os

import numpy as np
import pytest
import torch

from oracle import gno as ogno, graph as og
from tests import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def close(a, b, rtol=1e-5, atol_rel=2e-6, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    atol = atol_rel * max(b.abs().max().item(), 1e-30)
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max abs err {(a - b).abs().max().item():.3e} (ref max {b.abs().max().item():.3e})"


@pytest.mark.parametrize("tag", ["enc", "dec"])
def test_gno_golden_forward_backward(tag):
    from gaot_3d_b200.layers import IntegralTransform
    gld = torch.load(os.path.join(GOLD, "gno_golden.pt"))[tag]
    it = IntegralTransform(channel_mlp_layers=gld["layers"]).to(DEV)
    with torch.no_grad():
        for fc, w, b in zip(it.channel_mlp.fcs, gld["weights"], gld["biases"]):
            fc.weight.copy_(w); fc.bias.copy_(b)
    f = gld["f_y"].to(DEV).requires_grad_(True)
    out = it(gld["y_pos"].to(DEV), gld["x_pos"].to(DEV), gld["edge_index"].to(DEV), f)
    close(out, gld["out"], what=f"{tag} forward")
    out.backward(gld["d_out"].to(DEV))
    # gradients accumulate over ~1e4 edges: compare at 2e-5 of the gradient scale
    close(f.grad, gld["d_f"], rtol=2e-5, atol_rel=1e-5, what=f"{tag} d_f")
    for i, fc in enumerate(it.channel_mlp.fcs):
        close(fc.weight.grad, gld["d_weights"][i], rtol=1e-4, atol_rel=2e-5, what=f"{tag} dW{i}")
        close(fc.bias.grad, gld["d_biases"][i], rtol=1e-4, atol_rel=2e-5, what=f"{tag} db{i}")


@pytest.mark.parametrize("strategy,dec", [("knn", False), ("radius", False), ("bidirectional", True)])
def test_gno_vs_oracle_larger(strategy, dec):
    """32K-point cloud / 16^3 tokens: segments that span tiles, empty queries, long segments."""
    from gaot_3d_b200 import ops
    torch.manual_seed(0)
    N, C = 32768, 32
    phys, lat = synth.surface_cloud(N, seed=1), synth.latent_grid((16, 16, 16))
    ei = torch.from_numpy(og.get_neighbor_strategy_np(strategy, phys, None, lat, None, 0.15, 2, dec, workers=-1))
    ypos, xpos = (torch.from_numpy(lat), torch.from_numpy(phys)) if dec else (torch.from_numpy(phys), torch.from_numpy(lat))
    layers = [6, 64, 64, C] if dec else [6, 64, 64, 64, C]
    ws = [torch.randn(layers[i + 1], layers[i]) / np.sqrt(layers[i]) for i in range(len(layers) - 1)]
    bs = [torch.randn(layers[i + 1]) * 0.1 for i in range(len(layers) - 1)]
    f = torch.randn(ypos.shape[0], C)
    ref = ogno.integral_transform(ypos.double(), xpos.double(), ei, f.double(), [w.double() for w in ws], [b.double() for b in bs])
    ref32 = ogno.integral_transform(ypos, xpos, ei, f, ws, bs)
    csr = ops.build_csr(ei[0].to(DEV), ei[1].to(DEV), ypos.shape[0], xpos.shape[0])
    wd = [w.to(DEV).requires_grad_(True) for w in ws]
    bd = [b.to(DEV).requires_grad_(True) for b in bs]
    fd = f.to(DEV).requires_grad_(True)
    out = ops.gno(ypos.to(DEV), xpos.to(DEV), fd, csr, wd, bd, precision="fp32")
    # we must be as close to the fp64 truth as the reference's own fp32 evaluation is (x4 slack)
    err_ours = (out.double().cpu() - ref).abs().max().item()
    err_ref32 = (ref32.double() - ref).abs().max().item()
    assert err_ours <= 4 * err_ref32 + 1e-7, (err_ours, err_ref32)
    close(out, ref32, rtol=1e-5, atol_rel=4e-6, what="forward vs fp32 oracle")
    # determinism: two runs bit-identical (no atomics in the forward)
    out2 = ops.gno(ypos.to(DEV), xpos.to(DEV), fd, csr, wd, bd, precision="fp32")
    assert torch.equal(out, out2)
    # sum reduction (sharded-encoder partials) = mean * count
    outs = ops.gno(ypos.to(DEV), xpos.to(DEV), fd, csr, wd, bd, reduce="sum", precision="fp32")
    cnt = torch.bincount(ei[1], minlength=xpos.shape[0]).clamp(min=1).to(DEV)
    close(outs / cnt[:, None], out, rtol=1e-6, what="sum vs mean")
    # backward against fp64 autograd of the oracle
    g = torch.randn_like(ref)
    wr = [w.double().requires_grad_(True) for w in ws]
    br = [b.double().requires_grad_(True) for b in bs]
    fr = f.double().requires_grad_(True)
    ogno.integral_transform(ypos.double(), xpos.double(), ei, fr, wr, br).backward(g)
    out.backward(g.float().to(DEV))
    close(fd.grad, fr.grad, rtol=2e-5, atol_rel=1e-5, what="d_f")
    for i in range(len(ws)):
        close(wd[i].grad, wr[i].grad, rtol=1e-4, atol_rel=3e-5, what=f"dW{i}")
        close(bd[i].grad, br[i].grad, rtol=1e-4, atol_rel=3e-5, what=f"db{i}")


def test_gno_edge_cases():
    from gaot_3d_b200.layers import IntegralTransform
    it = IntegralTransform(channel_mlp_layers=[6, 64, 32]).to(DEV)
    y, x = torch.rand(100, 3, device=DEV), torch.rand(50, 3, device=DEV)
    f = torch.randn(100, 32, device=DEV)
    out = it(y, x, torch.empty(2, 0, dtype=torch.long, device=DEV), f)
    assert out.shape == (50, 32) and float(out.abs().max()) == 0.0
    # int32 edges in arbitrary order (precomputed-edge path, stat.py:191) + one edge + ragged segments
    ei = torch.tensor([[5, 7, 7, 99, 0], [49, 0, 49, 3, 0]], dtype=torch.int32, device=DEV)
    out = it(y, x, ei, f)
    ref = ogno.integral_transform(y.cpu(), x.cpu(), ei.long().cpu(), f.cpu(), [fc.weight.detach().cpu() for fc in it.channel_mlp.fcs],
                                  [fc.bias.detach().cpu() for fc in it.channel_mlp.fcs])
    close(out, ref, what="tiny ragged")
    # kernel-only transform and f_y=None
    it2 = IntegralTransform(channel_mlp_layers=[6 + 32, 64, 16], transform_type="nonlinear_kernelonly").to(DEV)
    out = it2(y, x, ei, f)
    ref = ogno.integral_transform(y.cpu(), x.cpu(), ei.long().cpu(), f.cpu(), [fc.weight.detach().cpu() for fc in it2.channel_mlp.fcs],
                                  [fc.bias.detach().cpu() for fc in it2.channel_mlp.fcs], transform_type="nonlinear_kernelonly")
    close(out, ref, what="nonlinear_kernelonly")
    it3 = IntegralTransform(channel_mlp_layers=[6 + 32, 64, 32], transform_type="nonlinear").to(DEV)
    fg = f.clone().requires_grad_(True)
    out = it3(y, x, ei, fg)
    fr = f.detach().cpu().double().requires_grad_(True)
    ref = ogno.integral_transform(y.cpu().double(), x.cpu().double(), ei.long().cpu(), fr,
                                  [fc.weight.detach().cpu().double() for fc in it3.channel_mlp.fcs],
                                  [fc.bias.detach().cpu().double() for fc in it3.channel_mlp.fcs], transform_type="nonlinear")
    close(out, ref, what="nonlinear")
    g = torch.randn(50, 32)
    ref.backward(g.double()); out.backward(g.to(DEV))
    close(fg.grad, fr.grad, rtol=2e-5, atol_rel=1e-5, what="nonlinear d_f")
    out = it(y, x, ei, None) if False else None     # f_y=None needs an MLP with matching width; covered below
    it4 = IntegralTransform(channel_mlp_layers=[6, 32, 8]).to(DEV)
    out = it4(y, x, ei, None)
    ref = ogno.integral_transform(y.cpu(), x.cpu(), ei.long().cpu(), None, [fc.weight.detach().cpu() for fc in it4.channel_mlp.fcs],
                                  [fc.bias.detach().cpu() for fc in it4.channel_mlp.fcs])
    close(out, ref, what="f_y=None")


def test_geo_golden_and_oracle():
    from gaot_3d_b200.layers import GeometricEmbedding
    gld = torch.load(os.path.join(GOLD, "geo_golden.pt"))
    ge = GeometricEmbedding(3, 32).to(DEV)
    ge.load_state_dict(gld["state"])
    src, qry, ei = gld["source_pos"].to(DEV), gld["query_pos"].to(DEV), gld["edge_index"].to(DEV)
    feats = ge.statistical_features(src, qry, ei)
    # the smallest covariance eigenvalue is ill-conditioned in the reference's own fp32 eigvalsh; compare
    # against the fp64 evaluation of the same formulas with the fp32 reference's error as the yardstick
    ref64 = ogno.geo_statistical_features(gld["source_pos"].double(), gld["query_pos"].double(), gld["edge_index"])
    err_ours = (feats.double().cpu() - ref64).abs().max(dim=0).values
    err_ref = (gld["features"].double() - ref64).abs().max(dim=0).values
    assert bool((err_ours <= 4 * err_ref + 1e-5).all()), (err_ours, err_ref)
    emb = ge(src, qry, ei)
    close(emb, gld["embedding"], rtol=1e-4, atol_rel=1e-4, what="embedding vs reference fp32")
    # un-normalised features, bigger cloud, includes empty queries
    phys, lat = synth.surface_cloud(32768, seed=4), synth.latent_grid((16, 16, 16))
    ei = torch.from_numpy(og.radius_np(phys, lat, 0.15, workers=-1)[::-1].copy())
    raw = ge.statistical_features(torch.from_numpy(phys).to(DEV), torch.from_numpy(lat).to(DEV), ei.to(DEV), normalize=False)
    ref = ogno.geo_statistical_features(torch.from_numpy(phys).double(), torch.from_numpy(lat).double(), ei, normalize=False)
    close(raw[:, :6], ref[:, :6], rtol=1e-4, atol_rel=1e-5, what="moments")
    close(raw[:, 6:], ref[:, 6:], rtol=1e-3, atol_rel=1e-4, what="eigenvalues")
    assert float(raw[(ref[:, 0] == 0).to(DEV)].abs().max()) == 0.0


@pytest.mark.parametrize("strategy,dec", [("knn", False), ("radius", True), ("bidirectional", False)])
def test_gno_bf16_tensor_core_path(strategy, dec):
    """BF16 tcgen05 MLP path: rtol 2e-2 (north star) against the fp64 oracle; deterministic."""
    from gaot_3d_b200 import ops
    torch.manual_seed(1)
    N, C = 32768, 32
    phys, lat = synth.surface_cloud(N, seed=2), synth.latent_grid((16, 16, 16))
    ei = torch.from_numpy(og.get_neighbor_strategy_np(strategy, phys, None, lat, None, 0.15, 2, dec, workers=-1))
    ypos, xpos = (torch.from_numpy(lat), torch.from_numpy(phys)) if dec else (torch.from_numpy(phys), torch.from_numpy(lat))
    layers = [6, 64, 64, C] if dec else [6, 64, 64, 64, C]
    ws = [torch.randn(layers[i + 1], layers[i]) / np.sqrt(layers[i]) for i in range(len(layers) - 1)]
    bs = [torch.randn(layers[i + 1]) * 0.1 for i in range(len(layers) - 1)]
    f = torch.randn(ypos.shape[0], C)
    ref = ogno.integral_transform(ypos.double(), xpos.double(), ei, f.double(), [w.double() for w in ws], [b.double() for b in bs])
    csr = ops.build_csr(ei[0].to(DEV), ei[1].to(DEV), ypos.shape[0], xpos.shape[0])
    wd = [w.to(DEV).requires_grad_(True) for w in ws]
    bd = [b.to(DEV).requires_grad_(True) for b in bs]
    fd = f.to(DEV).requires_grad_(True)
    out = ops.gno(ypos.to(DEV), xpos.to(DEV), fd, csr, wd, bd, precision="bf16")
    o = out.double().cpu()
    rel_l2 = ((o - ref).norm() / ref.norm()).item()
    rel_max = ((o - ref).abs().max() / ref.abs().max()).item()
    assert rel_l2 < 1e-2 and rel_max < 2e-2, (rel_l2, rel_max)
    assert torch.equal(out, ops.gno(ypos.to(DEV), xpos.to(DEV), fd, csr, wd, bd, precision="bf16"))
    # spatial sensitivity survives the bf16 operands (hi/lo coordinate split): moving the queries by r/10
    # must change the output like the oracle says
    xs = xpos + 0.015
    ref2 = ogno.integral_transform(ypos.double(), xs.double(), ei, f.double(), [w.double() for w in ws], [b.double() for b in bs])
    out2 = ops.gno(ypos.to(DEV), xs.to(DEV), fd, csr, wd, bd, precision="bf16").double().cpu()
    d_ref, d_out = ref2 - ref, out2 - o
    assert ((d_out - d_ref).norm() / d_ref.norm()).item() < 0.5      # plain bf16 coordinates would give O(1) here
    # backward (fp32 recompute) still matches the oracle's gradients at the mixed-precision tolerance
    g = torch.randn_like(ref)
    wr = [w.double().requires_grad_(True) for w in ws]
    fr = f.double().requires_grad_(True)
    ogno.integral_transform(ypos.double(), xpos.double(), ei, fr, wr, [b.double() for b in bs]).backward(g)
    out.backward(g.float().to(DEV))
    br = None
    e = ((fd.grad.double().cpu() - fr.grad).norm() / fr.grad.norm()).item()
    assert e < 2e-2, ("d_f", e)
    for i in range(len(ws)):
        e = ((wd[i].grad.double().cpu() - wr[i].grad).norm() / wr[i].grad.norm()).item()
        assert e < 2e-2, (f"dW{i}", e)
    # biases: compare against fp64 autograd as well
    wr2 = [w.double().requires_grad_(True) for w in ws]
    br2 = [b.double().requires_grad_(True) for b in bs]
    ogno.integral_transform(ypos.double(), xpos.double(), ei, f.double(), wr2, br2).backward(g)
    for i in range(len(bs)):
        e = ((bd[i].grad.double().cpu() - br2[i].grad).norm() / br2[i].grad.norm()).item()
        assert e < 2e-2, (f"db{i}", e)


def test_geo_moments_compose_over_source_shards():
    """Sharded-encoder statistics: moment sums of two disjoint physical-point shards add up to
    the moments of the whole cloud, and finishing them gives the features of the fused single-pass kernel."""
    from gaot_3d_b200 import ops
    phys, lat = synth.surface_cloud(20000, seed=5), synth.latent_grid((12, 12, 12))
    ei = torch.from_numpy(og.radius_np(phys, lat, 0.2, workers=-1)[::-1].copy())            # [phys, latent]
    P, L = torch.from_numpy(phys).to(DEV), torch.from_numpy(lat).to(DEV)
    csr = ops.build_csr(ei[0].to(DEV), ei[1].to(DEV), P.shape[0], L.shape[0])
    whole = ops.geo_stats(P, L, csr, normalize=False)
    mom = ops.geo_moments(P, L, csr)
    assert torch.equal(ops.geo_from_moments(mom, normalize=False), whole)                  # same code, same bits
    cut = 9000
    parts = []
    for lo, hi in ((0, cut), (cut, P.shape[0])):
        m = (ei[0] >= lo) & (ei[0] < hi)
        e = ei[:, m].clone()
        e[0] -= lo
        c = ops.build_csr(e[0].to(DEV), e[1].to(DEV), hi - lo, L.shape[0])
        parts.append(ops.geo_moments(P[lo:hi].contiguous(), L, c))
    both = ops.geo_from_moments(parts[0] + parts[1], normalize=False)
    ref = ogno.geo_statistical_features(torch.from_numpy(phys).double(), torch.from_numpy(lat).double(), ei, normalize=False)
    assert torch.equal(both[:, 0], whole[:, 0])                                              # counts are exact
    close(both[:, :6], ref[:, :6], rtol=1e-4, atol_rel=1e-5, what="sharded moments")
    close(both[:, 6:], ref[:, 6:], rtol=1e-3, atol_rel=1e-4, what="sharded eigenvalues")
    # z-scored features, then the decoder-side global z-score of row shards
    from gaot_3d_b200 import shard
    import torch.distributed as dist
    z = ops.geo_from_moments(parts[0] + parts[1], normalize=True)
    zr = ogno.geo_statistical_features(torch.from_numpy(phys).double(), torch.from_numpy(lat).double(), ei, normalize=True)
    close(z, zr, rtol=2e-3, atol_rel=2e-4, what="z-scored features")


@pytest.mark.parametrize("pooling", ["max", "mean"])
def test_pointnet_embedding_golden(pooling):
    """'pointnet' geometric embedding (reference geoembed.py:184-222): fused per-edge MLP + pool kernel against the
    reference module's own output and parameter gradients (tests/golden/make_pointnet_golden.py) and the CPU oracle."""
    from gaot_3d_b200.layers import GeometricEmbedding
    gld = torch.load(os.path.join(GOLD, "pointnet_golden.pt"))[pooling]
    ge = GeometricEmbedding(3, 16, method="pointnet", pooling=pooling).to(DEV)
    ge.load_state_dict(gld["state"])                       # same parameter names as the reference
    src, qry, ei = gld["source_pos"].to(DEV), gld["query_pos"].to(DEV), gld["edge_index"].to(DEV)
    out = ge(src, qry, ei)
    close(out, gld["out"], rtol=1e-5, atol_rel=2e-6, what="pointnet forward")
    empty = (torch.bincount(gld["edge_index"][1], minlength=qry.shape[0]) == 0).to(DEV)
    assert bool(empty.any()) and float(out.detach()[empty].abs().max()) == 0.0
    out.backward(gld["d_out"].to(DEV))
    for name, p in ge.named_parameters():
        close(p.grad, gld["grads"][name], rtol=1e-4, atol_rel=2e-5, what=f"pointnet d {name}")
    st = gld["state"]
    ref = ogno.geo_pointnet_embedding(gld["source_pos"], gld["query_pos"], gld["edge_index"], st["pointnet_mlp.0.weight"],
                                      st["pointnet_mlp.0.bias"], st["pointnet_mlp.2.weight"], st["pointnet_mlp.2.bias"],
                                      st["fc.0.weight"], st["fc.0.bias"], pooling)
    close(out, ref, rtol=1e-5, atol_rel=2e-6, what="pointnet vs oracle")
    # no edges at all -> zeros, as the reference
    z = ge(src, qry, torch.empty(2, 0, dtype=torch.long, device=DEV))
    assert z.shape == out.shape and float(z.abs().max()) == 0.0


@pytest.mark.parametrize("attention_type", ["cosine", "dot_product"])
def test_gno_attentional_transform(attention_type):
    """use_attn (reference integral_transform.py:128-141,:161-165): segment-softmax edge weights, SUM reduction; the
    fused kernel takes the weights per CSR edge.  Against the fp64 oracle: output, d f_y, kernel-MLP gradients and (for
    'dot_product') the gradients of the learnable score projections."""
    from gaot_3d_b200.layers import IntegralTransform
    torch.manual_seed(3)
    phys, lat = synth.surface_cloud(6000, seed=6), synth.latent_grid((8, 8, 8))
    ei = torch.from_numpy(og.radius_np(phys, lat, 0.25, workers=-1)[::-1].copy())            # [phys, latent]
    ypos, xpos = torch.from_numpy(phys), torch.from_numpy(lat)
    it = IntegralTransform(channel_mlp_layers=[6, 64, 64, 32], use_attn=True, coord_dim=3, attention_type=attention_type).to(DEV)
    f = torch.randn(ypos.shape[0], 32)
    fd = f.to(DEV).requires_grad_(True)
    out = it(ypos.to(DEV), xpos.to(DEV), ei.to(DEV), fd)
    params = {n: p.detach().cpu().double().requires_grad_(True) for n, p in it.named_parameters()}
    attn = dict(type=attention_type, coord_dim=3)
    if attention_type == "dot_product":
        attn.update(wq=params["query_proj.weight"], bq=params["query_proj.bias"], wk=params["key_proj.weight"], bk=params["key_proj.bias"])
    fr = f.double().requires_grad_(True)
    ws = [params[f"channel_mlp.fcs.{i}.weight"] for i in range(3)]
    bs = [params[f"channel_mlp.fcs.{i}.bias"] for i in range(3)]
    ref = ogno.integral_transform(ypos.double(), xpos.double(), ei, fr, ws, bs, attn=attn)
    close(out, ref, rtol=1e-5, atol_rel=4e-6, what="attentional forward")
    g = torch.randn_like(ref)
    ref.backward(g)
    out.backward(g.float().to(DEV))
    close(fd.grad, fr.grad, rtol=2e-5, atol_rel=1e-5, what="d_f")
    for n, p in it.named_parameters():
        if n == "key_proj.bias":
            # a bias on the keys shifts every score of a segment by the same amount: softmax-invariant, gradient == 0
            assert float(params[n].grad.abs().max()) < 1e-12 and float(p.grad.abs().max()) < 1e-5
            continue
        close(p.grad, params[n].grad, rtol=1e-4, atol_rel=5e-5, what=f"d {n}")


@pytest.mark.parametrize("attention_type", ["cosine", "dot_product"])
def test_gno_attentional_transform_tensor_core(attention_type):
    """The attentional variant on the second-generation tensor-core kernels (per-edge weight in the epilogue, d weight out
    of the backward): rtol 2e-2 tier against the fp64 oracle, relative L2 like the other mixed-precision GNO tests."""
    from gaot_3d_b200 import ops
    from gaot_3d_b200.layers import IntegralTransform
    torch.manual_seed(4)
    phys, lat = synth.surface_cloud(6000, seed=7), synth.latent_grid((8, 8, 8))
    ei = torch.from_numpy(og.radius_np(phys, lat, 0.25, workers=-1)[::-1].copy())
    ypos, xpos = torch.from_numpy(phys), torch.from_numpy(lat)
    it = IntegralTransform(channel_mlp_layers=[6, 64, 64, 32], use_attn=True, coord_dim=3, attention_type=attention_type).to(DEV)
    f = torch.randn(ypos.shape[0], 32)
    fd = f.to(DEV).requires_grad_(True)
    prev = ops.get_gno_precision()
    ops.set_gno_precision("bf16")
    try:
        out = it(ypos.to(DEV), xpos.to(DEV), ei.to(DEV), fd)
        g = torch.randn(xpos.shape[0], 32)
        out.backward(g.to(DEV))
    finally:
        ops.set_gno_precision(prev)
    params = {n: p.detach().cpu().double().requires_grad_(True) for n, p in it.named_parameters()}
    attn = dict(type=attention_type, coord_dim=3)
    if attention_type == "dot_product":
        attn.update(wq=params["query_proj.weight"], bq=params["query_proj.bias"], wk=params["key_proj.weight"], bk=params["key_proj.bias"])
    fr = f.double().requires_grad_(True)
    ref = ogno.integral_transform(ypos.double(), xpos.double(), ei, fr, [params[f"channel_mlp.fcs.{i}.weight"] for i in range(3)],
                                  [params[f"channel_mlp.fcs.{i}.bias"] for i in range(3)], attn=attn)
    ref.backward(g.double())
    rl2 = lambda a, b: ((a.detach().double().cpu() - b).norm() / b.norm().clamp(min=1e-30)).item()
    assert rl2(out, ref) < 1e-2, ("out", rl2(out, ref))
    assert rl2(fd.grad, fr.grad) < 2e-2, ("d_f", rl2(fd.grad, fr.grad))
    for n, p in it.named_parameters():
        if n == "key_proj.bias":
            continue                                            # identically zero (softmax shift invariance)
        assert rl2(p.grad, params[n].grad) < 2e-2, (n, rl2(p.grad, params[n].grad))


def test_two_dimensional_coordinates():
    """gno_coord_dim = 2 (the MAGNOConfig default, reference magno.py:23): graph build, integral transform, statistical and
    PointNet geometric embedding on [N, 2] coordinates against the (dimension-generic, reference-pinned) oracle."""
    import gaot_3d_b200 as G
    from gaot_3d_b200.layers import GeometricEmbedding, IntegralTransform
    torch.manual_seed(3)
    rng = np.random.default_rng(5)
    phys = rng.uniform(-1, 1, (6000, 2)).astype(np.float32)
    gx = np.linspace(-1, 1, 24, dtype=np.float32)
    lat = np.stack(np.meshgrid(gx, gx, indexing="ij"), -1).reshape(-1, 2)
    r, C = 0.11, 16
    P, L = torch.from_numpy(phys), torch.from_numpy(lat)
    for strategy, dec in (("radius", False), ("knn", True), ("bidirectional", False)):
        ref_e = og.get_neighbor_strategy_np(strategy, phys, None, lat, None, r, 2, dec)
        e = G.get_neighbor_strategy(strategy, P.to(DEV), None, L.to(DEV), None, r, 2, dec)
        assert np.array_equal(e.cpu().numpy(), ref_e), f"2-D {strategy} edges"
    ei = torch.from_numpy(og.get_neighbor_strategy_np("radius", phys, None, lat, None, r, 2, False))    # [phys; latent]
    it = IntegralTransform(channel_mlp_layers=[4, 32, C]).to(DEV)
    f = torch.randn(len(phys), C)
    fd = f.to(DEV).requires_grad_(True)
    out = it(P.to(DEV), L.to(DEV), ei.to(DEV), fd)
    ws = [l.weight.detach().cpu().clone().requires_grad_(True) for l in it.channel_mlp.fcs]
    bs = [l.bias.detach().cpu().clone().requires_grad_(True) for l in it.channel_mlp.fcs]
    fr = f.clone().requires_grad_(True)
    ref = ogno.integral_transform(P, L, ei, fr, ws, bs)
    close(out, ref, what="2-D gno forward")
    go = torch.randn_like(ref)
    out.backward(go.to(DEV))
    ref.backward(go)
    close(fd.grad, fr.grad, rtol=2e-5, atol_rel=1e-5, what="2-D d_f")
    for i, l in enumerate(it.channel_mlp.fcs):
        assert l.weight.grad.shape == ws[i].shape
        close(l.weight.grad, ws[i].grad, rtol=1e-4, atol_rel=2e-5, what=f"2-D dW{i}")
    ge = GeometricEmbedding(2, 8).to(DEV)
    feat = ge.statistical_features(P.to(DEV), L.to(DEV), ei.to(DEV))
    reff = ogno.geo_statistical_features(P, L, ei)
    assert feat.shape == reff.shape == (len(lat), 7)
    close(feat, reff, rtol=1e-3, atol_rel=1e-4, what="2-D statistical features")
    gp = GeometricEmbedding(2, 8, method="pointnet").to(DEV)
    st = {k: v.detach().cpu() for k, v in gp.state_dict().items()}
    refp = ogno.geo_pointnet_embedding(P, L, ei, st["pointnet_mlp.0.weight"], st["pointnet_mlp.0.bias"], st["pointnet_mlp.2.weight"],
                                       st["pointnet_mlp.2.bias"], st["fc.0.weight"], st["fc.0.bias"], "max")
    close(gp(P.to(DEV), L.to(DEV), ei.to(DEV)), refp, rtol=1e-4, atol_rel=1e-5, what="2-D pointnet")
