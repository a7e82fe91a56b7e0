"""CPU check of the per-query bodies in gaot_3d_b200/csrc/graph_core.cuh (the code the CUDA
kernels run one-thread-per-query) through the host driver tests/emu/graph_emu.cpp, against
the oracle.  Bit-exact: integer / index work."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import graph
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ctypes.c_void_p


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "graph_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests/emu/graph_emu.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.emu_radius.restype = ctypes.c_int64
    lib.emu_knn.restype = ctypes.c_int64
    lib.emu_radius.argtypes = [P, ctypes.c_int64, P, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, P, P]
    lib.emu_knn.argtypes = [P, ctypes.c_int64, P, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P]
    return lib


def emu_radius(lib, x, y, r, cap=32, mc=1 << 21, md=1024):
    x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
    oy = np.zeros(max(1, len(y) * cap), np.int64)
    ox = np.zeros_like(oy)
    E = lib.emu_radius(x.ctypes.data, len(x), y.ctypes.data, len(y), r, cap, mc, md, oy.ctypes.data, ox.ctypes.data)
    assert E >= 0, "count pass and emit pass disagree"
    return np.stack([oy[:E], ox[:E]])


def emu_knn(lib, x, y, k, mc=1 << 21, md=1024):
    x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
    oy = np.zeros(max(1, len(y) * k), np.int64)
    ox = np.zeros_like(oy)
    E = lib.emu_knn(x.ctypes.data, len(x), y.ctypes.data, len(y), k, mc, md, oy.ctypes.data, ox.ctypes.data)
    return np.stack([oy[:E], ox[:E]])


CASES = [(5000, (16, 16, 8), 0.15), (20000, (32, 32, 16), 0.07), (3000, (8, 8, 8), 0.6), (100, (4, 4, 4), 0.05)]


@pytest.mark.parametrize("N,G,r", CASES)
def test_radius_surface(emu, N, G, r):
    phys, lat = synth.surface_cloud(N, seed=N), synth.latent_grid(G)
    for xs, ys in ((phys, lat), (lat, phys)):
        assert np.array_equal(graph.radius_np(xs, ys, r), emu_radius(emu, xs, ys, r))
        for cap in (1, 5):
            assert np.array_equal(graph.radius_np(xs, ys, r, max_num_neighbors=cap), emu_radius(emu, xs, ys, r, cap))
        for mc, md in ((64, 8), (1000, 16)):   # bounded grids: h grows past r, reach stays valid
            assert np.array_equal(graph.radius_np(xs, ys, r), emu_radius(emu, xs, ys, r, 32, mc, md))


@pytest.mark.parametrize("N,G,r", CASES)
@pytest.mark.parametrize("k", [1, 3, 17])
def test_knn_surface(emu, N, G, r, k):
    phys, lat = synth.surface_cloud(N, seed=N), synth.latent_grid(G)
    ref = graph.knn_np(lat, phys, k)
    assert np.array_equal(ref, emu_knn(emu, lat, phys, k))
    assert np.array_equal(ref, emu_knn(emu, lat, phys, k, 50, 6))
    assert np.array_equal(graph.knn_np(phys, lat, k), emu_knn(emu, phys, lat, k))


def test_adversarial(emu):
    rng = np.random.default_rng(1)
    x = np.repeat(rng.uniform(-1, 1, (50, 3)).astype(np.float32), 4, 0)      # duplicates
    y = rng.uniform(-3, 3, (500, 3)).astype(np.float32)                      # queries outside the grid
    for k in (1, 2, 7):
        assert np.array_equal(graph.knn_bruteforce(x, y, k), emu_knn(emu, x, y, k))
    for r in (0.3, 1.0, 5.0):
        assert np.array_equal(graph.radius_bruteforce(x, y, r), emu_radius(emu, x, y, r))
    g = np.linspace(-1, 1, 9).astype(np.float32)
    L = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    yq = (L[rng.integers(0, len(L), 400)] + np.float32(0.125) * rng.integers(-1, 2, (400, 3))).astype(np.float32)
    for k in (1, 2, 4, 9):                                                   # exact distance ties
        assert np.array_equal(graph.knn_bruteforce(L, yq, k), emu_knn(emu, L, yq, k))
    assert np.array_equal(graph.radius_bruteforce(L, yq, 0.25), emu_radius(emu, L, yq, 0.25))  # d == r excluded
    x1 = np.zeros((10, 3), np.float32)
    assert np.array_equal(graph.knn_bruteforce(x1, y, 3), emu_knn(emu, x1, y, 3))
    xp = rng.uniform(-1, 1, (2000, 3)).astype(np.float32)
    xp[:, 2] = 0.5                                                           # flat source set
    assert np.array_equal(graph.knn_bruteforce(xp, y, 2), emu_knn(emu, xp, y, 2))
    assert np.array_equal(graph.radius_bruteforce(xp, y, 0.4), emu_radius(emu, xp, y, 0.4))
    assert np.array_equal(graph.knn_bruteforce(x[:1], y, 1), emu_knn(emu, x[:1], y, 1))
    assert emu_radius(emu, x, y, 0.0).shape == (2, 0)
