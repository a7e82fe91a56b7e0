"""torchrun tool: kernel-level trace of sharded 8M-point steps on EVERY rank (torch.profiler / CUPTI; no nsys in the image).
Each rank writes <out>_<world>_rank<r>.txt: device busy vs span, the NCCL kernels in issue order with their durations (a long
one = this rank waited for a peer), the largest idle gaps with the kernels either side, and the kernels by total device time.
  python -m torch.distributed.run --nproc-per-node N ... profiles/tools/shard_trace.py [--workload drivaerml8m] [--no-graph]
Exits through os._exit after writing (the profiler's teardown with live NCCL communicators hung a round-2 run).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="drivaerml8m")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/shard_trace")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gaot_3d_b200 import tgraph
    if args.no_graph:
        tgraph.set_enabled(False)
    job = bench.Job(args.workload, "shard" if world > 1 else "single", dev, rank, world)
    job.warm(4)
    ms = job.timed(5) / 5
    from torch.profiler import profile, ProfilerActivity
    job.barrier()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            job.step(job.resident[i % job.nsamp])
        torch.cuda.synchronize()
    n = float(args.steps)
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and not e.name.startswith("nccl:")]
    ev.sort(key=lambda e: e.time_range.start)
    t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
    busy, last_end, last_name, gaps = 0.0, None, "", []
    for e in ev:                                       # union of the device intervals (streams overlap)
        s, t = e.time_range.start, e.time_range.end
        if last_end is None or s > last_end:
            if last_end is not None and s - last_end > 30:
                gaps.append((s - last_end, (last_end - t0) / 1e3, last_name[:60], e.name[:60]))
            busy += t - s
            last_end, last_name = t, e.name
        elif t > last_end:
            busy += t - last_end
            last_end, last_name = t, e.name
    agg = {}
    for e in ev:
        k = e.name[:100]
        c, tot = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, tot + (e.time_range.end - e.time_range.start))
    lines = [f"rank {rank} of {world}  workload {args.workload}  graphs {'off' if args.no_graph else 'on'}  timed {ms:.2f} ms/step",
             f"{args.steps} profiled steps: span {(t1 - t0) / 1e3:.2f} ms, device busy (union over streams) {busy / 1e3:.2f} ms, {len(ev)} device events, "
             f"idle gaps > 30 us: {len(gaps)} totalling {sum(g[0] for g in gaps) / 1e3:.2f} ms"]
    lines.append("--- NCCL kernels in issue order (start ms, duration us)")
    for e in ev:
        if "nccl" in e.name.lower():
            d = e.time_range.end - e.time_range.start
            if d > 60:
                lines.append(f"  {(e.time_range.start - t0) / 1e3:8.3f}  {d:9.1f}  {e.name[:60]}")
    lines.append("--- largest idle gaps (us, at ms, after kernel, before kernel)")
    for g in sorted(gaps, reverse=True)[:25]:
        lines.append(f"  {g[0]:8.1f}  {g[1]:8.3f}  {g[2]}  ->  {g[3]}")
    lines.append("--- kernels by device time")
    for k, (c, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        lines.append(f"{tot / n / 1e3:9.3f} ms/step  {c / n:7.1f} calls/step  {k}")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out + f"_{world}_rank{rank}.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    if rank == 0:
        print("\n".join(lines[:40]), flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
