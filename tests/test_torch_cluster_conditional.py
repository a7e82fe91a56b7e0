"""SURVEY.md §8(c) parity protocol (2): when a REAL torch_cluster is importable (site-packages, or a wheel the driver
installed under baseline/_ref/), the restated oracle (oracle/graph.py) and the CUDA search are checked against it:

  * queries the 32-cap does not touch -> identical neighbour SETS vs torch_cluster's CPU path (and its CUDA path on a GPU);
  * capped queries -> identical to torch_cluster's CUDA path (first 32 by ascending source index); against its CPU path
    (nanoflann traversal order) only "a 32-subset of the true ball";
  * kNN -> identical (distance, index) order where distances are distinct.

torch_cluster is an un-vendored, unpinned wheel (reference requirements / README.md:112-113) and is ABSENT from this
image, so these tests skip here; they exist so that the pin tightens automatically wherever the wheel is present.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)

torch_cluster = pytest.importorskip("torch_cluster", reason="torch_cluster is not installed in this image (parity unpinned for the search)")

import torch  # noqa: E402

from oracle import graph as og  # noqa: E402
from tests import synth  # noqa: E402

N, GRID, R, K = 20000, (16, 16, 8), 0.12, 3


def _clouds():
    return synth.surface_cloud(N, seed=4), synth.latent_grid(GRID)


def _groups(row, col, n):
    order = np.lexsort((col, row))
    row, col = row[order], col[order]
    starts = np.searchsorted(row, np.arange(n + 1))
    return [col[starts[i]:starts[i + 1]] for i in range(n)]


def test_oracle_radius_vs_torch_cluster_cpu():
    phys, lat = _clouds()
    ours = og.radius_np(phys, lat, R)                                   # [latent (y), phys (x)]
    tc = torch_cluster.radius(torch.from_numpy(phys), torch.from_numpy(lat), R, max_num_neighbors=32).numpy()
    true_ball = og.radius_np(phys, lat, R, max_num_neighbors=10 ** 9)
    go, gt, gb = (_groups(e[0], e[1], len(lat)) for e in (ours, tc, true_ball))
    for q in range(len(lat)):
        if len(gb[q]) <= 32:
            assert np.array_equal(np.sort(gt[q]), go[q]), f"uncapped query {q}"
        else:
            assert len(gt[q]) == 32 and np.isin(gt[q], gb[q]).all(), f"capped query {q}: not a 32-subset of the ball"
            assert np.array_equal(go[q], gb[q][:32])


def test_oracle_knn_vs_torch_cluster_cpu():
    phys, lat = _clouds()
    ours = og.knn_np(lat, phys, K)                                      # [phys (y), latent (x)]
    tc = torch_cluster.knn(torch.from_numpy(lat), torch.from_numpy(phys), K).numpy()
    d2 = og.dist2_f32(lat[ours[1]], phys[ours[0]]).reshape(N, K)
    distinct = (np.diff(d2, axis=1) > 0).all(axis=1)                    # rows without distance ties: order is forced
    o, t = ours[1].reshape(N, K), tc[1].reshape(N, K)
    assert np.array_equal(o[distinct], t[distinct])


@pytest.mark.gpu
def test_cuda_search_vs_torch_cluster_cuda():
    from gaot_3d_b200 import ops
    phys, lat = _clouds()
    P, L = torch.from_numpy(phys).cuda(), torch.from_numpy(lat).cuda()
    ry, cx = ops.radius(P, L, R)
    tc = torch_cluster.radius(P, L, R, max_num_neighbors=32)
    assert torch.equal(torch.stack([ry, cx]), tc), "capped + uncapped queries: bit-exact vs torch_cluster CUDA"
    ky, kx = ops.knn(L, P, K)
    tk = torch_cluster.knn(L, P, K)
    d2 = og.dist2_f32(lat[kx.cpu().numpy()], phys[ky.cpu().numpy()]).reshape(N, K)
    distinct = torch.from_numpy((np.diff(d2, axis=1) > 0).all(axis=1)).cuda()
    assert torch.equal(kx.view(N, K)[distinct], tk[1].view(N, K)[distinct])
