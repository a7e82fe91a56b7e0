"""torchrun worker: intra-sample sharded GAOT3D forward/backward on R GPUs == single-GPU result.
Launched by tests/test_gpu_shard.py (needs >= 2 GPUs) or by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/shard_worker.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gaot_3d_b200 as G  # noqa: E402
from gaot_3d_b200 import shard  # noqa: E402
from tests import synth  # noqa: E402


CASES = ((["radius", "reverse"], [False, False]), ("knn", [False, False]), (["bidirectional", "radius"], [False, False]),
         ("bidirectional", [True, False]), (["radius", "knn"], [True, True]))


def parity_cases(dev, rank, world, mode="sp", verbose=True, cases=CASES):
    """Sharded (this process group) vs unsharded (every rank, replicated) forward / loss / parameter gradients of small
    GAOT3D models.  Returns {"ok": bool, "cases": [...]}; also used by bench.py's `shard_parity` key."""
    out, ok = [], True
    for strat, geo in cases:
        torch.manual_seed(0)
        N, grid, r, k = 40000, (16, 16, 8), 0.12, 2
        pos = torch.from_numpy(synth.surface_cloud(N, seed=1)).to(dev)
        c = torch.from_numpy(synth.unit_normals(N)).to(dev)
        tgt = torch.randn(N, 4, device=dev)
        lat = torch.from_numpy(synth.latent_grid(grid)).to(dev)
        mc = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=32, neighbor_strategy=strat, gno_radius=r, mlp_type="linear",
                           precompute_edges=False, use_geoembed=geo, encoder_feature_attr=["pos", "c"], k_neighbors=k)
        heads = 8 if world > 4 else 4                                     # head_dim 32; the heads split across the ranks
        tc = G.TransformerConfig(patch_size=2, hidden_size=32 * heads, num_layers=2, positional_embedding="rope")
        tc.attn_config.hidden_size, tc.attn_config.num_heads, tc.attn_config.num_kv_heads = 32 * heads, heads, heads
        tc.attn_config.atten_dropout = 0.0
        tc.ffn_config.hidden_size = 256
        model = G.GAOT3D(6, 4, mc, tc, latent_tokens=grid).to(dev)        # same seed -> identical replicas
        # ---- unsharded reference on every rank
        y_full = model(G.Batch(pos=pos, c=c), tokens_pos=lat)
        loss_full = torch.nn.functional.mse_loss(y_full, tgt)
        loss_full.backward()
        g_full = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        model.zero_grad(set_to_none=True)
        # ---- sharded
        lo, hi = shard.shard_range(N, rank, world)
        y_loc = shard.sharded_forward(model, G.Batch(pos=pos[lo:hi].contiguous(), c=c[lo:hi].contiguous()), lat, N, mode=mode)
        loss_loc = ((y_loc - tgt[lo:hi]) ** 2).sum() / (N * 4)            # local share of the global mean
        loss_loc.backward()
        shard.allreduce_partial_grads(model, mode=mode)
        tot = loss_loc.detach().clone()
        dist.all_reduce(tot)
        e_out = (y_loc - y_full[lo:hi]).abs().max().item() / y_full.abs().max().item()
        e_loss = abs(tot.item() - loss_full.item()) / abs(loss_full.item())
        worst, missing = 0.0, 0
        for n, p in model.named_parameters():
            if n not in g_full:
                continue
            if p.grad is None:
                missing += 1
                continue
            rel = ((p.grad - g_full[n]).norm() / g_full[n].norm().clamp(min=1e-12)).item()
            worst = max(worst, rel)
        stat = torch.tensor([e_out, e_loss, worst, float(missing)], device=dev)
        dist.all_reduce(stat, op=dist.ReduceOp.MAX)                       # every rank checks its own rows: report the worst
        e_out, e_loss, worst, missing = stat.tolist()
        good = e_out < 2e-2 and e_loss < 1e-3 and worst < 5e-2 and missing == 0
        ok &= good
        out.append({"strategy": strat, "geoembed": geo, "out_rel_err": e_out, "loss_rel_err": e_loss, "worst_grad_rel_l2": worst, "ok": good})
        if rank == 0 and verbose:
            print(f"strategy={strat} geoembed={geo}: out rel err {e_out:.2e}, loss rel err {e_loss:.2e}, worst grad rel l2 {worst:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        model.zero_grad(set_to_none=True)
        from gaot_3d_b200 import tgraph
        tgraph.reset()
    return {"ok": bool(ok), "mode": mode, "ranks": world, "bars": {"out": 2e-2, "loss": 1e-3, "grad_rel_l2": 5e-2}, "cases": out}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for mode in ("sp", "hp"):
        if rank == 0:
            print(f"--- mode {mode}", flush=True)
        ok &= parity_cases(dev, rank, world, mode=mode)["ok"]
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() < 1.0:
        sys.exit(1)
    if rank == 0:
        print("SHARD_OK")


if __name__ == "__main__":
    main()
