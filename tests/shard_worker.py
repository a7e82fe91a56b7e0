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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for strat, geo in ((["radius", "reverse"], [False, False]), ("knn", [False, False]), (["bidirectional", "radius"], [False, False]),
                       ("bidirectional", [True, False]), (["radius", "knn"], [True, True])):
        torch.manual_seed(0)
        N, grid, r, k = 40000, (16, 16, 8), 0.12, 2
        pos = torch.from_numpy(synth.surface_cloud(N, seed=1)).to(dev)
        c = torch.from_numpy(synth.unit_normals(N)).to(dev)
        tgt = torch.randn(N, 4, device=dev)
        lat = torch.from_numpy(synth.latent_grid(grid)).to(dev)
        mc = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=32, neighbor_strategy=strat, gno_radius=r, mlp_type="linear",
                           precompute_edges=False, use_geoembed=geo, encoder_feature_attr=["pos", "c"], k_neighbors=k)
        tc = G.TransformerConfig(patch_size=2, hidden_size=128, num_layers=2, positional_embedding="rope")
        tc.attn_config.hidden_size, tc.attn_config.num_heads, tc.attn_config.num_kv_heads = 128, 4, 4
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
        y_loc = shard.sharded_forward(model, G.Batch(pos=pos[lo:hi].contiguous(), c=c[lo:hi].contiguous()), lat, N)
        loss_loc = ((y_loc - tgt[lo:hi]) ** 2).sum() / (N * 4)            # local share of the global mean
        loss_loc.backward()
        shard.allreduce_partial_grads(model)
        tot = loss_loc.detach().clone()
        dist.all_reduce(tot)
        e_out = (y_loc - y_full[lo:hi]).abs().max().item() / y_full.abs().max().item()
        e_loss = abs(tot.item() - loss_full.item()) / abs(loss_full.item())
        worst = 0.0
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            rel = ((p.grad - g_full[n]).norm() / g_full[n].norm().clamp(min=1e-12)).item()
            worst = max(worst, rel)
        good = e_out < 2e-2 and e_loss < 1e-3 and worst < 5e-2
        ok &= good
        if rank == 0:
            print(f"strategy={strat} geoembed={geo}: out rel err {e_out:.2e}, loss rel err {e_loss:.2e}, worst grad rel l2 {worst:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        model.zero_grad(set_to_none=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() < 1.0:
        sys.exit(1)
    if rank == 0:
        print("SHARD_OK")


if __name__ == "__main__":
    main()
