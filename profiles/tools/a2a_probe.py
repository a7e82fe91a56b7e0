"""torchrun tool: the peer-memory all-to-all (csrc/a2a.cu + symmetric-memory barrier) against ncclSend/Recv groups:
equality of the results, eager and CUDA-graph-replayed time per exchange at the sharded transformer's message size.
  python -m torch.distributed.run --nproc-per-node N ... profiles/tools/a2a_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from gaot_3d_b200 import tblock
    S, W = 16384, 768                                   # tokens x (q|k|v) columns of the benchmark's transformer
    send = torch.randn(world, S // world, W // world, device=dev).to(torch.bfloat16)
    ref = torch.empty_like(send)
    dist.all_to_all_single(ref, send)
    out = tblock._all_to_all(send, None).clone()
    ok = torch.equal(out, ref)
    if rank == 0:
        print(f"backend {tblock.a2a_backend()}  setup error {__import__('gaot_3d_b200.p2p', fromlist=['x'])._CTX.get('error')}  equal to NCCL: {ok}", flush=True)
    for i in range(5):                                  # alternating buffers, back to back
        s2 = send + i
        r2 = torch.empty_like(s2)
        dist.all_to_all_single(r2, s2)
        ok &= torch.equal(tblock._all_to_all(s2, None).clone(), r2)

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    recv = torch.empty_like(send)
    t_nccl = timeit(lambda: dist.all_to_all_single(recv, send))
    t_p2p = timeit(lambda: tblock._all_to_all(send, None))
    # graph-replayed (what the sharded step does): 20 exchanges per graph
    res = {}
    for name, fn in (("nccl", lambda: dist.all_to_all_single(recv, send)), ("p2p", lambda: tblock._all_to_all(send, None))):
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(); fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(); dist.barrier()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                for _ in range(20):
                    fn()
            g.replay(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 100 * 1e3
        except Exception as e:
            res[name] = repr(e)[:200]
    # the other collectives of the sharded step: reduce-scatter / all-gather of the latent field, gradient all-reduce
    from gaot_3d_b200 import p2p
    torch.manual_seed(rank)
    part = torch.randn(131072, 33, device=dev)
    ref_rs = torch.empty(131072 // world, 33, device=dev)
    dist.reduce_scatter_tensor(ref_rs, part)
    rs = p2p.reduce_scatter(part)
    ok_rs = rs is not None and torch.allclose(rs, ref_rs, rtol=1e-5, atol=1e-5)
    rs2 = p2p.reduce_scatter(part)
    ok_rs &= rs2 is not None and torch.equal(rs, rs2)                     # deterministic
    slab = torch.randn(131072 // world, 32, device=dev)
    ref_ag = torch.empty(131072, 32, device=dev)
    dist.all_gather_into_tensor(ref_ag, slab)
    ag = p2p.all_gather(slab)
    ok_ag = ag is not None and torch.equal(ag, ref_ag)
    flat = torch.randn(11_000_003, device=dev)
    ref_ar = flat.clone()
    dist.all_reduce(ref_ar)
    mine = flat.clone()
    ok_ar = p2p.all_reduce_(mine) and torch.allclose(mine, ref_ar, rtol=1e-5, atol=1e-5)
    t = {}
    t["rs_nccl"] = timeit(lambda: dist.reduce_scatter_tensor(ref_rs, part), 20)
    t["rs_p2p"] = timeit(lambda: p2p.reduce_scatter(part), 20)
    t["ag_nccl"] = timeit(lambda: dist.all_gather_into_tensor(ref_ag, slab), 20)
    t["ag_p2p"] = timeit(lambda: p2p.all_gather(slab), 20)
    t["ar_nccl"] = timeit(lambda: dist.all_reduce(ref_ar), 20)
    t["ar_p2p"] = timeit(lambda: p2p.all_reduce_(mine), 20)
    if rank == 0:
        print(f"reduce-scatter ok {ok_rs}  all-gather ok {ok_ag}  all-reduce ok {ok_ar};  us per call " + "  ".join(f"{k} {v:.1f}" for k, v in t.items()), flush=True)
    ok &= bool(ok_rs and ok_ag and ok_ar)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world}: block {send[0].numel() * 2} B per peer; eager us/exchange nccl {t_nccl:.1f} p2p {t_p2p:.1f}; graph-replayed {res}; all equal {bool(flag.item())}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
