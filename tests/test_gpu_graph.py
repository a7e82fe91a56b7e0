"""GPU parity: cell-list radius / kNN / coalesce / strategy composition vs the CPU oracle.
Bit-exact (integer index work): identical edge lists in identical order."""
import os

import numpy as np
import pytest
import torch

from oracle import graph as og
from tests import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def dev():
    return torch.device("cuda:0")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def test_golden_strategies():
    import gaot_3d_b200 as g
    z = np.load(os.path.join(GOLD, "graph_golden.npz"))
    phys, lat, r, k = t(z["phys"]), t(z["lat"]), float(z["r"]), int(z["k"])
    bp = torch.zeros(phys.shape[0], dtype=torch.long, device=dev())
    bl = torch.zeros(lat.shape[0], dtype=torch.long, device=dev())
    for name, dec in (("knn", False), ("radius", False), ("bidirectional", False), ("knn", True), ("radius", True),
                      ("bidirectional", True), ("reverse", True)):
        e = g.get_neighbor_strategy(name, phys, bp, lat, bl, r, k, dec)
        ref = z[f"{'dec' if dec else 'enc'}_{name}"]
        assert e.dtype == torch.long and e.shape[0] == 2
        assert np.array_equal(e.cpu().numpy(), ref), f"{name} dec={dec}"
    with pytest.raises(ValueError):
        g.get_neighbor_strategy("nope", phys, bp, lat, bl, r, k, False)
    with pytest.raises(ValueError):
        g.get_neighbor_strategy("reverse", phys, bp, lat, bl, r, k, False)   # encoder has no 'reverse'


@pytest.mark.parametrize("N,G,r", [(32768, (16, 16, 16), 0.15), (50000, (32, 32, 16), 0.066), (777, (4, 4, 4), 0.5)])
def test_radius_knn_vs_oracle(N, G, r):
    from gaot_3d_b200 import ops
    phys, lat = synth.surface_cloud(N, seed=N), synth.latent_grid(G)
    for xs, ys in ((phys, lat), (lat, phys)):
        ry, cx = ops.radius(t(xs), t(ys), r)
        ref = og.radius_np(xs, ys, r, workers=-1)
        assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), ref)
        ry, cx = ops.radius(t(xs), t(ys), r, max_num_neighbors=7)
        assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.radius_np(xs, ys, r, max_num_neighbors=7, workers=-1))
    for k in (1, 4, 20):
        ry, cx = ops.knn(t(lat), t(phys), k)
        assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.knn_np(lat, phys, k, workers=-1))


def test_adversarial_and_empty():
    from gaot_3d_b200 import ops
    rng = np.random.default_rng(1)
    x = np.repeat(rng.uniform(-1, 1, (50, 3)).astype(np.float32), 4, 0)
    y = rng.uniform(-3, 3, (500, 3)).astype(np.float32)
    for k in (1, 2, 7):
        ry, cx = ops.knn(t(x), t(y), k)
        assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.knn_bruteforce(x, y, k))
    for r in (0.3, 1.0, 5.0):
        ry, cx = ops.radius(t(x), t(y), r)
        assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.radius_bruteforce(x, y, r))
    g = np.linspace(-1, 1, 9).astype(np.float32)
    L = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    yq = (L[rng.integers(0, len(L), 400)] + np.float32(0.125) * rng.integers(-1, 2, (400, 3))).astype(np.float32)
    for k in (1, 2, 4, 9):
        ry, cx = ops.knn(t(L), t(yq), k)
        assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.knn_bruteforce(L, yq, k))
    ry, cx = ops.radius(t(L), t(yq), 0.25)
    assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.radius_bruteforce(L, yq, 0.25))
    # empty inputs / no matches / k > #sources
    e = torch.empty(0, 3, device=dev())
    assert ops.radius(e, t(y), 0.5)[0].numel() == 0 and ops.radius(t(x), e, 0.5)[0].numel() == 0
    assert ops.knn(e, t(y), 3)[0].numel() == 0
    assert ops.radius(t(x), t(y + 100.0), 0.5)[0].numel() == 0
    ry, cx = ops.knn(t(x[:3]), t(y), 8)
    assert np.array_equal(torch.stack([ry, cx]).cpu().numpy(), og.knn_bruteforce(x[:3], y, 8))


def test_batched_examples():
    import gaot_3d_b200 as g
    rng = np.random.default_rng(5)
    phys = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    lat = np.concatenate([synth.latent_grid((6, 6, 6), "unit")] * 3)
    bp = np.sort(rng.integers(0, 3, 3000))
    bl = np.repeat(np.arange(3), 216)
    for name, dec in (("knn", False), ("radius", True), ("bidirectional", False), ("reverse", True)):
        e = g.get_neighbor_strategy(name, t(phys), t(bp), t(lat), t(bl), 0.4, 2, dec)
        ref = og.get_neighbor_strategy_np(name, phys, bp, lat, bl, 0.4, 2, dec)
        assert np.array_equal(e.cpu().numpy(), ref), name


def test_coalesce_csr_mask_properties():
    from gaot_3d_b200 import ops
    rng = np.random.default_rng(7)
    E, n0, n1 = 200000, 5000, 700
    r0, r1 = rng.integers(0, n0, E), rng.integers(0, n1, E)
    o0, o1 = ops.coalesce(t(r0), t(r1), n0 - 1, n1 - 1)
    ref = og.coalesce_np(np.stack([r0, r1]))
    assert np.array_equal(torch.stack([o0, o1]).cpu().numpy(), ref)
    # idempotent
    p0, p1 = ops.coalesce(o0, o1, n0 - 1, n1 - 1)
    assert torch.equal(p0, o0) and torch.equal(p1, o1)
    # CSR side-band: stable by query, row sums = bincount
    csr = ops.build_csr(t(r0), t(r1), n0, n1)
    cnt = np.bincount(r1, minlength=n1)
    assert np.array_equal(np.diff(csr.rowptr.cpu().numpy()), cnt)
    order = np.argsort(r1, kind="stable")
    assert np.array_equal(csr.perm.cpu().numpy(), order)
    assert np.array_equal(csr.src.cpu().numpy(), r0[order]) and np.array_equal(csr.qry.cpu().numpy(), r1[order])
    # edge mask: order preserved, deterministic for a seed, keep ratio ~ 1-p
    m0, m1 = ops.edge_mask(t(r0), t(r1), 0.3, seed=42)
    n0_, n1_ = ops.edge_mask(t(r0), t(r1), 0.3, seed=42)
    assert torch.equal(m0, n0_) and torch.equal(m1, n1_)
    frac = m0.numel() / E
    assert abs(frac - 0.7) < 0.01
    key = m0.cpu().numpy() * n1 + m1.cpu().numpy()
    full = r0 * n1 + r1
    it = iter(full)                      # subsequence check
    assert all(any(kk == f for f in it) for kk in key[:2000])


def test_host_buffer_entry_points():
    """The C ABI with HOST buffers (the e2e arm): same edges as the device path."""
    import ctypes
    from gaot_3d_b200 import _lib
    lib = _lib.load()
    phys, lat = synth.surface_cloud(20000, seed=2), synth.latent_grid((16, 16, 8))
    cap = 32
    oy = np.zeros(len(lat) * cap, np.int64); ox = np.zeros_like(oy); E = ctypes.c_int64(0)
    rc = lib.gaot_radius_host(phys.ctypes.data, len(phys), lat.ctypes.data, len(lat), 0.1, cap, oy.ctypes.data, ox.ctypes.data, ctypes.byref(E))
    assert rc == 0, lib.gaot_last_error()
    assert np.array_equal(np.stack([oy[:E.value], ox[:E.value]]), og.radius_np(phys, lat, 0.1, workers=-1))
    oy = np.zeros(len(phys) * 2, np.int64); ox = np.zeros_like(oy)
    rc = lib.gaot_knn_host(lat.ctypes.data, len(lat), phys.ctypes.data, len(phys), 2, oy.ctypes.data, ox.ctypes.data, ctypes.byref(E))
    assert rc == 0, lib.gaot_last_error()
    assert np.array_equal(np.stack([oy[:E.value], ox[:E.value]]), og.knn_np(lat, phys, 2, workers=-1))


def test_sample_scope_cache_reuses_searches_without_changing_results():
    """Inside one forward (graph.sample_scope) the reverse decoder graph reuses the bidirectional encoder build and repeated knn
    searches run once; outside a scope nothing is cached; an in-place change of the coordinates misses."""
    import gaot_3d_b200 as G
    from gaot_3d_b200 import graph, ops
    DEV = torch.device("cuda:0")
    phys = torch.from_numpy(synth.surface_cloud(20000, seed=4)).to(DEV)
    lat = torch.from_numpy(synth.latent_grid((12, 12, 12))).to(DEV)
    r, k = 0.2, 2
    ref_enc = G.get_neighbor_strategy("bidirectional", phys, None, lat, None, r, k, False)
    ref_dec = G.get_neighbor_strategy("reverse", phys, None, lat, None, r, k, True)
    ops.reset_launch_count()
    G.get_neighbor_strategy("reverse", phys, None, lat, None, r, k, True)
    cold = ops.launch_count()
    with graph.sample_scope():
        e = G.get_neighbor_strategy("bidirectional", phys, None, lat, None, r, k, False)
        ops.reset_launch_count()
        d = G.get_neighbor_strategy("reverse", phys, None, lat, None, r, k, True)
        warm = ops.launch_count()
        e_knn = G.get_neighbor_strategy("knn", phys, None, lat, None, r, k, False)
        ops.reset_launch_count()
        d_knn = G.get_neighbor_strategy("knn", phys, None, lat, None, r, k, True)
        assert ops.launch_count() == 0, "the decoder's knn search must come from the cache"
        phys.add_(0.0)                                     # bumps the version counter: the next search is a miss
        ops.reset_launch_count()
        G.get_neighbor_strategy("knn", phys, None, lat, None, r, k, True)
        assert ops.launch_count() > 0
    assert torch.equal(e, ref_enc) and torch.equal(d, ref_dec) and torch.equal(d_knn, e_knn.flip(0))
    assert warm == 0 and cold > 0, (warm, cold)
    assert not graph._CACHE["entries"], "the scope's exit must drop every entry"
