"""World-size-2 gloo test (CPU) of the intra-sample sharding host logic: partitioning,
the global radius-cap exchange, and the autograd collectives, with the CPU oracle standing in for
the local CUDA ops.  Sharded == unsharded: bit-exact edges, fp tolerance on sums and gradients."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaot_3d_b200 import shard
        from oracle import gno as ogno, graph as og
        from tests import synth
        torch.manual_seed(0)
        N, G, r, C = 3001, (6, 6, 6), 0.3, 8
        phys, lat = synth.surface_cloud(N, seed=3), synth.latent_grid(G)
        M = len(lat)
        lo, hi = shard.shard_range(N, rank, world)
        # --- radius cap composes over contiguous shards: bit-exact with the unsharded capped graph
        full = og.radius_np(phys, lat, r)                                    # [latent, phys], capped at 32 globally
        loc = og.radius_np(phys[lo:hi], lat, r)                              # locally capped
        li, pi = shard.apply_global_radius_cap(torch.from_numpy(loc[0]), torch.from_numpy(loc[1]), M, 32)
        mine = full[:, (full[1] >= lo) & (full[1] < hi)]
        assert np.array_equal(np.stack([li.numpy(), pi.numpy() + lo]), mine), "cap exchange"
        assert (np.bincount(full[0], minlength=M) == 32).any(), "test must exercise the cap"
        # --- encoder partial sums + counts all-reduced == unsharded mean; decoder shard-local;
        #     gradients: d latent all-reduced, parameter grads summed
        dims = [6, 16, C]
        ws = [torch.nn.Parameter(torch.randn(dims[i + 1], dims[i]) * 0.3) for i in range(2)]
        bs = [torch.nn.Parameter(torch.randn(dims[i + 1]) * 0.1) for i in range(2)]
        mix = torch.nn.Parameter(torch.randn(C, C) * 0.3)                    # stands in for the replicated transformer
        f = torch.randn(N, C)
        P, L = torch.from_numpy(phys), torch.from_numpy(lat)
        e_loc = torch.stack([pi, li])                                        # [local phys, latent]
        part = ogno.scatter_sum(ogno.mlp_forward(torch.cat([P[lo:hi][e_loc[0]], L[e_loc[1]]], -1), ws, bs) * f[lo:hi][e_loc[0]], e_loc[1], M)
        cnt = torch.bincount(e_loc[1], minlength=M).float()
        dist.all_reduce(cnt)
        latent = shard.all_reduce_forward(part) / cnt.clamp(min=1)[:, None]
        z = shard.all_reduce_backward(latent @ mix)
        dec_e = e_loc.flip(0)
        out = ogno.integral_transform(L, P[lo:hi], dec_e, z, ws, bs)
        loss = out.pow(2).sum() / N                                          # local share of the global mean
        loss.backward()
        pg = torch.cat([p.grad.reshape(-1) for p in ws + bs])
        dist.all_reduce(pg)                                                  # GNO-side params: partial -> SUM
        tot = loss.detach().clone()
        dist.all_reduce(tot)
        # --- unsharded reference on every rank
        ws2 = [w.detach().clone().requires_grad_(True) for w in ws]
        bs2 = [b.detach().clone().requires_grad_(True) for b in bs]
        mix2 = mix.detach().clone().requires_grad_(True)
        e_full = torch.from_numpy(full[::-1].copy())                         # [phys, latent]
        lat_full = ogno.integral_transform(P, L, e_full, f, ws2, bs2)
        out_full = ogno.integral_transform(L, P, e_full.flip(0), lat_full @ mix2, ws2, bs2)
        loss_full = out_full.pow(2).sum() / N
        loss_full.backward()
        assert torch.allclose(latent, lat_full, rtol=1e-5, atol=1e-6), "latent"
        assert torch.allclose(out, out_full[lo:hi], rtol=1e-4, atol=1e-6), "decoder rows"
        assert torch.allclose(tot, loss_full, rtol=1e-5), "loss"
        pg2 = torch.cat([p.grad.reshape(-1) for p in ws2 + bs2])
        assert torch.allclose(pg, pg2, rtol=1e-3, atol=1e-6), "param grads"
        assert torch.allclose(mix.grad, mix2.grad, rtol=1e-3, atol=1e-6), "replicated-part grad is already total"
        # --- decoder-side geometric embedding: z-score of row-sharded features over ALL rows (geoembed.py:177-180)
        feats = torch.randn(N, 9, generator=torch.Generator().manual_seed(5)) * torch.arange(1, 10) + 3.0
        feats[:, 4] = 2.5                                                    # constant column: std < 1e-6 -> 1
        zl = shard.global_zscore(feats[lo:hi], N)
        std = feats.std(dim=0)
        zf = (feats - feats.mean(dim=0)) / torch.where(std < 1e-6, torch.ones_like(std), std)
        assert torch.allclose(zl, zf[lo:hi], rtol=1e-4, atol=1e-5), "global z-score"
        # --- the token-sharded ("sp") mode's collectives as autograd functions: reduce-scatter forward <-> all-gather backward,
        #     all-gather forward <-> reduce-scatter backward (gloo has neither natively: shard._reduce_scatter / _all_gather)
        Mtok = 2 * world * 3
        xp = (torch.arange(Mtok * 4, dtype=torch.float32).view(Mtok, 4) * (rank + 1)).requires_grad_(True)
        slab = shard.reduce_scatter_forward(xp)                              # [Mtok / world, 4] = (sum_r x_r)[my slab]
        n_s = Mtok // world
        tot_ref = torch.arange(Mtok * 4, dtype=torch.float32).view(Mtok, 4) * sum(range(1, world + 1))
        assert torch.equal(slab, tot_ref[rank * n_s:(rank + 1) * n_s]), "reduce-scatter forward"
        full = shard.all_gather_forward(slab * 2.0)                          # every rank: the whole field
        assert torch.equal(full, tot_ref * 2.0), "all-gather forward"
        wgt = torch.arange(Mtok * 4, dtype=torch.float32).view(Mtok, 4) / 7.0 + rank    # a different consumer per rank
        (full * wgt).sum().backward()
        # d full is a per-rank partial -> reduce-scatter sums it over ranks for my slab; then x2; then the reduce-scatter's
        # backward all-gathers the slab gradients: every rank's x receives the gradient of every slab
        wsum = sum(torch.arange(Mtok * 4, dtype=torch.float32).view(Mtok, 4) / 7.0 + r for r in range(world))
        assert torch.allclose(xp.grad, 2.0 * wsum), "sp collectives backward"
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_shard_range():
    from gaot_3d_b200.shard import shard_range
    for n in (0, 1, 7, 500000, 8000001):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
