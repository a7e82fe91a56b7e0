"""Pins the graph oracle (oracle/graph.py): KD-tree path == brute-force evaluation of the definition,
golden vectors, and the property tests of SURVEY.md §4 (pure CPU)."""
import os

import numpy as np
import pytest

from oracle import graph as og
from tests import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_kdtree_equals_bruteforce(seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (1500, 3)).astype(np.float32)
    y = rng.uniform(-1.2, 1.2, (300, 3)).astype(np.float32)
    for r in (0.05, 0.2, 0.7):
        assert np.array_equal(og.radius_np(x, y, r), og.radius_bruteforce(x, y, r))
        assert np.array_equal(og.radius_np(x, y, r, max_num_neighbors=3), og.radius_bruteforce(x, y, r, max_num_neighbors=3))
    for k in (1, 5, 40):
        assert np.array_equal(og.knn_np(x, y, k), og.knn_bruteforce(x, y, k))
    bx, by = np.sort(rng.integers(0, 4, 1500)), np.sort(rng.integers(0, 4, 300))
    assert np.array_equal(og.radius_np(x, y, 0.3, bx, by), og.radius_bruteforce(x, y, 0.3, bx, by))
    assert np.array_equal(og.knn_np(x, y, 3, bx, by), og.knn_bruteforce(x, y, 3, bx, by))


def test_golden_graph():
    z = np.load(os.path.join(GOLD, "graph_golden.npz"))
    phys, lat, r, k = z["phys"], z["lat"], float(z["r"]), int(z["k"])
    assert np.array_equal(og.radius_np(phys, lat, r), z["enc_radius_raw"])
    assert np.array_equal(og.radius_np(lat, phys, r), z["dec_radius_raw"])
    assert np.array_equal(og.knn_np(lat, phys, k), z["knn_raw"])
    for name, dec in (("knn", False), ("radius", False), ("bidirectional", False), ("knn", True), ("radius", True),
                      ("bidirectional", True), ("reverse", True)):
        e = og.get_neighbor_strategy_np(name, phys, None, lat, None, r, k, dec)
        assert np.array_equal(e, z[f"{'dec' if dec else 'enc'}_{name}"]), name


def test_semantics_and_properties():
    rng = np.random.default_rng(3)
    phys, lat = synth.surface_cloud(4000, seed=9), synth.latent_grid((8, 8, 8))
    r, k = 0.3, 2
    # the 32-cap keeps the FIRST 32 by ascending source index
    e = og.radius_np(phys, lat, r)
    full = og.radius_np(phys, lat, r, max_num_neighbors=10 ** 6)
    cnt = np.bincount(e[0], minlength=len(lat))
    assert cnt.max() == 32 and np.bincount(full[0], minlength=len(lat)).max() > 32
    for q in np.nonzero(cnt == 32)[0][:20]:
        assert np.array_equal(e[1][e[0] == q], np.sort(full[1][full[0] == q])[:32])
    # strict inequality: a point at distance exactly r is excluded
    x = np.array([[0, 0, 0], [0.5, 0, 0]], np.float32)
    assert og.radius_np(x, np.array([[0, 0, 0]], np.float32), 0.5).shape[1] == 1
    # knn ties -> lower index
    x = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0]], np.float32)
    assert og.knn_np(x, np.zeros((1, 3), np.float32), 2)[1].tolist() == [0, 1]
    # reverse(dec) == flip(bidirectional(enc)); coalesce idempotent and sorted
    enc = og.get_neighbor_strategy_np("bidirectional", phys, None, lat, None, r, k, False)
    dec = og.get_neighbor_strategy_np("reverse", phys, None, lat, None, r, k, True)
    assert np.array_equal(dec, enc[::-1])
    assert np.array_equal(og.coalesce_np(enc), enc)
    key = enc[0] * (enc.max() + 1) + enc[1]
    assert np.all(np.diff(key) > 0)
    # invariance to a permutation of the physical points (after index remap), uncapped case
    perm = rng.permutation(len(phys))
    inv = np.argsort(perm)
    a = og.radius_np(lat, phys, r)                      # decoder radius: cap not hit
    b = og.radius_np(lat, phys[perm], r)
    b = np.stack([inv[b[0]] * 0 + perm[b[0]], b[1]])
    assert np.array_equal(og.sort_edges(a), og.sort_edges(b))
    with pytest.raises(ValueError):
        og.get_neighbor_strategy_np("nope", phys, None, lat, None, r, k, False)
    with pytest.raises(ValueError):
        og.radius_np(phys[:10], lat[:10], r, np.array([1, 0] * 5), None)    # unsorted batch vector
