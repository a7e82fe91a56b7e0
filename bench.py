#!/usr/bin/env python
"""bench.py -- GAOT-3D hot path on B200: fwd+bwd samples/s (with GNO edges/s and per-kernel roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload drivaernet500k|small]

Workload (BASELINE.json configs[1]): DrivAerNet++-shaped 500K-point surface cloud, latent 64x64x32,
kNN encoder / decoder (k=1), C=32, pos+normals in, 4 output channels (pressure + WSS), 10-layer
transformer (H=256, 8 heads, FFN 1024, RoPE, patch 2 -> S=16384), one sample per step, fwd + bwd +
AdamW step, online graph build inside the step.  N>1: one process per GPU, one sample per rank per
step (the reference's DDP mode), gradients all-reduced over NCCL -> weak scaling.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "drivaernet500k": dict(n_points=500_000, latent=(64, 64, 32), box="drivaernet", k=1, layers=10, hidden=256, heads=8, ffn=1024),
    # BASELINE configs[2]: NASA-CRM-shaped sample, bidirectional knn+radius graph, geometric embedding in the encoder only
    "crm500k": dict(n_points=500_000, latent=(64, 64, 32), box="crm", k=1, layers=10, hidden=256, heads=8, ffn=1024,
                    strategy="bidirectional", radius=0.033, geoembed=[True, False], features="mach_aoa"),
    "small": dict(n_points=32_768, latent=(16, 16, 16), box="drivaernet", k=1, layers=4, hidden=256, heads=8, ffn=1024),
    # BASELINE configs[3]: DrivaerML-shaped full-resolution sample; with --shard the physical points of ONE
    # sample are split across the ranks (encoder partial sums all-reduced, decoder query-sharded)
    "drivaerml8m": dict(n_points=8_000_000, latent=(64, 64, 32), box="drivaerml", k=1, layers=10, hidden=256, heads=8, ffn=1024,
                        strategy=["bidirectional", "reverse"], radius=0.033),
}
C_LIFT, C_IN, C_OUT, PATCH = 32, 6, 4, 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def make_sample(wl, seed):
    from tests import synth
    pos = synth.surface_cloud(wl["n_points"], wl["box"], seed=seed)
    nrm = synth.mach_aoa(wl["n_points"], seed=seed) if wl.get("features") == "mach_aoa" else synth.unit_normals(wl["n_points"], seed=seed)
    rng = np.random.default_rng(seed + 7)
    tgt = rng.standard_normal((wl["n_points"], C_OUT)).astype(np.float32)
    return pos, nrm, tgt


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = str(gpu_index), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == self.idx:
                self.rows.append(f)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = []
        for j, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(r[j].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_cfg(wl):
    return dict(k=wl["k"], C=C_LIFT, hidden=wl["hidden"], heads=wl["heads"], ffn=wl["ffn"], num_layers=wl["layers"], patch=PATCH,
                latent_tokens=wl["latent"], enc_mlp=[6, 64, 64, 64, C_LIFT], dec_mlp=[6, 64, 64, C_LIFT])


def run_cpu_arm(wl, steps, warmup, name):
    """The reference's CPU path (oracle port) on the host cores: bounded sample per step."""
    from oracle import cpu_step
    from tests import synth
    lat = synth.latent_grid(wl["latent"], wl["box"])
    cores = torch.get_num_threads()
    times = []
    for i in range(warmup + steps):
        pos, nrm, tgt = make_sample(wl, seed=i % 4)
        r = cpu_step.cpu_step_seconds(pos, nrm, tgt, lat, cpu_cfg(wl), layers_timed=1, search_workers=-1, seed=i)
        if i >= warmup:
            times.append(r["seconds"])
        last = r
    sec = float(np.mean(times))
    sample = (f"per step: full kNN graph build (scipy cKDTree, all cores), full GNO encoder+decoder fwd+bwd at E={wl['n_points'] * wl['k']}, "
              f"1 of {wl['layers']} transformer blocks fwd+bwd at S={(wl['latent'][0] // PATCH) * (wl['latent'][1] // PATCH) * (wl['latent'][2] // PATCH)} "
              f"scaled x{wl['layers']}; torch CPU fp32, SDPA as in reference attn.py:126")
    return sec, cores, sample, last["parts"]


def gno_large_graph(lib, G, pos, lat, dev, pk_hbm, precision, radius=0.033):
    """GNO edges/s where the kernel, not the launch, is what is timed: the radius graphs of the same cloud
    (decoder: every point -> latents within r, ~12 edges per point; encoder: capped at 32 per latent), fused
    GNO forward and backward timed by the in-library CUDA events (outside the step timing)."""
    import ctypes
    from gaot_3d_b200 import ops
    out = {}
    for name, dec, layers in (("decoder_radius", True, [6, 64, 64, C_LIFT]), ("encoder_radius", False, [6, 64, 64, 64, C_LIFT])):
        ei = G.get_neighbor_strategy("radius", pos, None, lat, None, radius, 1, dec)
        ypos, xpos = (lat, pos) if dec else (pos, lat)
        torch.manual_seed(1)
        ws = [(torch.randn(layers[i + 1], layers[i], device=dev) / layers[i] ** 0.5).requires_grad_(True) for i in range(len(layers) - 1)]
        bs = [torch.zeros(layers[i + 1], device=dev, requires_grad=True) for i in range(len(layers) - 1)]
        f_y = torch.randn(ypos.shape[0], C_LIFT, device=dev, requires_grad=True)
        csr = ops.csr_of(ei, ypos.shape[0], xpos.shape[0])
        go = torch.randn(xpos.shape[0], C_LIFT, device=dev)
        for _ in range(3):
            ops.gno(ypos, xpos, f_y, csr, ws, bs, precision=precision).backward(go)
        lib.gaot_profile_enable(1)
        for _ in range(5):
            ops.gno(ypos, xpos, f_y, csr, ws, bs, precision=precision).backward(go)
        buf = ctypes.create_string_buffer(1 << 14)
        lib.gaot_profile_summary(buf, len(buf))
        lib.gaot_profile_enable(0)
        E = int(ei.shape[1])
        rec = {"edges": E, "mlp": layers}
        for line in buf.value.decode().strip().splitlines():
            nm, calls, ms, work = line.split()
            if nm in ("gno_fwd", "gno_bwd"):
                t = float(ms) / int(calls) * 1e-3
                rec[nm + "_ms"] = t * 1e3
                rec[nm + "_edges_per_s"] = E / t
                rec[nm + "_hbm_frac"] = float(work) / int(calls) / t / 1e9 / pk_hbm       # algorithmic bytes / measured copy peak
        if "gno_fwd_ms" in rec and "gno_bwd_ms" in rec:
            rec["fwd_bwd_edges_per_s"] = E / ((rec["gno_fwd_ms"] + rec["gno_bwd_ms"]) * 1e-3)
        out[name] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="drivaernet500k", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gno-precision", default="bf16", choices=["fp32", "bf16"],
                    help="per-edge kernel MLP operands: bf16 tcgen05 (rtol 2e-2 tier of the north star, default) or fp32 CUDA cores (rtol 1e-5 tier)")
    ap.add_argument("--node-mlp", default="fused", choices=["fp32", "tf32", "fused"],
                    help="node-level MLPs (lifting / projection / recovery): strict fp32 torch GEMMs, TF32 torch GEMMs (what the reference's "
                         "default Conv1d node MLPs get from cuDNN), or TF32 + the fused f16/bf16 tensor-core kernel for the projection head")
    ap.add_argument("--shard", action="store_true", help="intra-sample sharding: all ranks cooperate on ONE sample (strong scaling)")
    ap.add_argument("--profile-step", action="store_true", help="warm up, then run ONE step between cudaProfilerStart/Stop (for ncu --profile-from-start off) and exit")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that print to the C-level stdout (NCCL prints its version line
    # there) are sent to stderr; the JSON line goes to the saved descriptor
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    S = (wl["latent"][0] // PATCH) * (wl["latent"][1] // PATCH) * (wl["latent"][2] // PATCH)
    strat = wl.get("strategy", "knn")
    strat_txt = (f"{strat[0]} encoder / {strat[1]} decoder" if isinstance(strat, (list, tuple)) else f"{strat} enc+dec") + \
                (f" (k={wl['k']}, r={wl.get('radius', 0.033)})" if strat != "knn" else f" k={wl['k']}")
    shape = {"drivaernet": "DrivAerNet++", "drivaerml": "DrivaerML", "crm": "NASA-CRM"}[wl["box"]]
    c_in = 5 if wl.get("features") == "mach_aoa" else C_IN
    config = {"workload": f"{shape}-shaped {wl['n_points']}-point surface cloud, latent {wl['latent']}, {strat_txt}, "
                          f"C={C_LIFT}, in {c_in} ({'pos+Mach/AOA' if c_in == 5 else 'pos+normals'}), out 4 (pressure+WSS), "
                          f"geoembed {wl.get('geoembed', [False, False])}, {wl['layers']}-layer transformer H={wl['hidden']} S={S}, "
                          f"fwd+bwd+AdamW, online graph build, batch 1/GPU, atten_dropout 0",
              "n_points": wl["n_points"], "latent_tokens": list(wl["latent"]), "seq_len": S,
              "parallelism": (f"shard{world} (one sample split across ranks)" if args.shard else f"dp{world}") if world > 1 else "single",
              "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; 4 distinct samples cycled"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sec, cores, sample, parts = run_cpu_arm(wl, args.steps, min(args.warmup, 1), "reference")
        v = 1.0 / sec
        emit({"impl": "reference", "metric": "fwd+bwd samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                                           "parts_s": parts},
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ our arm (B200)
    import gaot_3d_b200 as G
    from gaot_3d_b200 import _lib, ops
    from tests import synth
    import ctypes
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    G.set_gno_precision(args.gno_precision)
    G.set_node_mlp_mode({"fp32": "torch", "tf32": "tf32", "fused": "fused"}[args.node_mlp])
    torch.manual_seed(0)
    mc = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=C_LIFT, neighbor_strategy=wl.get("strategy", "knn"), k_neighbors=wl["k"],
                       gno_radius=wl.get("radius", 0.033),
                       mlp_type="linear", precompute_edges=False, use_geoembed=wl.get("geoembed", [False, False]), encoder_feature_attr=["pos", "c"],
                       in_gno_channel_mlp_hidden_layers=[64, 64, 64], out_gno_channel_mlp_hidden_layers=[64, 64], projection_channels=256)
    tc = G.TransformerConfig(patch_size=PATCH, hidden_size=wl["hidden"], num_layers=wl["layers"], positional_embedding="rope")
    tc.attn_config.hidden_size, tc.attn_config.num_heads, tc.attn_config.num_kv_heads = wl["hidden"], wl["heads"], wl["heads"]
    tc.attn_config.atten_dropout = 0.0
    tc.ffn_config.hidden_size = wl["ffn"]
    model = G.GAOT3D(c_in, C_OUT, mc, tc, latent_tokens=wl["latent"]).to(dev).train()
    if world > 1 and not args.shard:
        model_step = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    else:
        model_step = model
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=1e-5, fused=True)
    lat = torch.from_numpy(synth.latent_grid(wl["latent"], wl["box"])).to(dev)
    host = []
    n_total = wl["n_points"]
    nsamp = 2 if n_total > 2_000_000 else 4
    for s in range(nsamp):
        if args.shard and world > 1:
            from gaot_3d_b200 import shard as _shard
            pos, nrm, tgt = make_sample(wl, seed=s)                 # every rank generates the same sample, keeps its range
            lo, hi = _shard.shard_range(n_total, rank, world)
            pos, nrm, tgt = pos[lo:hi].copy(), nrm[lo:hi].copy(), tgt[lo:hi].copy()
        else:
            pos, nrm, tgt = make_sample(wl, seed=s + 100 * rank)
        host.append(tuple(torch.from_numpy(a).pin_memory() for a in (pos, nrm, tgt)))
    resident = [tuple(t.to(dev) for t in h) for h in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def step(sample):
        pos, nrm, tgt = sample
        opt.zero_grad(set_to_none=True)
        if args.shard and world > 1:
            y = _shard.sharded_forward(model, G.Batch(pos=pos, c=nrm), lat, n_total)
            loss = ((y - tgt) ** 2).sum() / (n_total * C_OUT)       # local share of the global mean
            loss.backward()
            _shard.allreduce_partial_grads(model)
        else:
            y = model_step(G.Batch(pos=pos, c=nrm), tokens_pos=lat)
            loss = torch.nn.functional.mse_loss(y, tgt)
            loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(n):
            if e2e:
                smp = tuple(t.to(dev, non_blocking=True) for t in host[i % nsamp])
                float(step(smp).item())                     # D2H read of the step's loss
            else:
                step(resident[i % nsamp])
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step(resident[i % nsamp])
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    clocks = ClockSampler(local_rank)
    clocks.start()
    lib.gaot_profile_enable(1)
    lib.gaot_launch_count_reset()
    ms_total = timed(args.steps, e2e=False)
    launches = int(lib.gaot_launch_count())
    buf = ctypes.create_string_buffer(1 << 16)
    lib.gaot_profile_summary(buf, len(buf))
    lib.gaot_profile_enable(0)
    clk = clocks.stop()
    ms_e2e = timed(args.steps, e2e=True)
    gno_large = gno_large_graph(lib, G, resident[0][0], lat, dev, pk_hbm=peaks()["hbm"], precision=args.gno_precision) \
        if (world == 1 and wl["n_points"] <= 1_000_000) else None

    ms_step = ms_total / args.steps
    samples_per_step = 1 if (args.shard and world > 1) else world
    value = samples_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = samples_per_step * args.steps / (ms_e2e * 1e-3)
    pk = peaks()
    kernels = {}
    for line in buf.value.decode().strip().splitlines():
        name, calls, ms, work = line.split()
        calls, ms, work = int(calls), float(ms), float(work)
        kernels[name] = {"calls_per_step": calls / args.steps, "ms_per_step": ms / args.steps, "share_of_step": ms / ms_total,
                         "avg_launch_ms": ms / calls, "work_per_launch": work / calls}
    tensor_bound = {"attn_fwd", "attn_bwd", "linear_fwd", "linear_bwd_x", "linear_bwd_w"}
    for name, kd in kernels.items():
        rate = kd["work_per_launch"] / (kd["avg_launch_ms"] * 1e-3)
        if name in tensor_bound:
            kd.update(bound="tensor", achieved=rate / 1e12, peak=pk["tf_sust"], unit="TFLOP/s", frac=rate / 1e12 / pk["tf_sust"])
        else:
            kd.update(bound="hbm", achieved=rate / 1e9, peak=pk["hbm"], unit="GB/s", frac=rate / 1e9 / pk["hbm"])
    dom = max(kernels, key=lambda n: kernels[n]["ms_per_step"]) if kernels else None
    roofline = None
    if dom:
        kd = kernels[dom]
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json,
        # written by profiles/summarize.py from dram__bytes_read.sum + dram__bytes_write.sum); null if never captured
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        roofline = {"kernel": dom, "bound": kd["bound"], "achieved": kd["achieved"], "peak": kd["peak"], "unit": kd["unit"],
                    "frac": kd["frac"], "traffic": traffic, "peak_source": f"{pk['src']} (MEASURED_PEAKS.json, sustained bf16 / copy bandwidth)",
                    "share_of_step": kd["share_of_step"]}
    E = wl["n_points"] * wl["k"]
    gno = {}
    if "gno_fwd" in kernels and "gno_bwd" in kernels and wl.get("strategy", "knn") == "knn":     # E is known only for the knn graphs
        f, b = kernels["gno_fwd"], kernels["gno_bwd"]
        gno = {"edges_per_launch": E, "fwd_edges_per_s": E / (f["avg_launch_ms"] * 1e-3),
               "fwd_bwd_edges_per_s": E / ((f["avg_launch_ms"] + b["avg_launch_ms"]) * 1e-3),
               "note": "encoder and decoder launches averaged (same E for knn k=1)", "precision": args.gno_precision}
    out = {"metric": "fwd+bwd samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.shard else "weak", "vs_baseline": None,
           "dtype": ("bf16 tensor-core operands / f32 accumulate (attention, transformer dense layers" +
                     (", GNO edge MLP)" if args.gno_precision == "bf16" else "); f32 GNO edge MLP") +
                     f"; f32 residual stream, statistics, optimizer; node MLPs {args.node_mlp}"), "data": "synthetic",
           "config": config, "clocks": clk,
           "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "gno_edges_per_s": gno}
    if gno_large:
        out["gno_edges_per_s"]["large_graphs"] = gno_large
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                sec, cores, sample, parts = run_cpu_arm(wl, 1, 1, "cpu_baseline")
                out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample, "parts_s": parts}
            except Exception as e:  # the baseline is reporting only; never lose the GPU line
                out["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": f"failed: {e}"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
