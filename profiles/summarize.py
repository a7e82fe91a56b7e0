"""Turns gpurun_out/ ncu artefacts into the committed text summaries under profiles/.
usage: python profiles/summarize.py <tag>   (reads gpurun_out/launches.csv and gpurun_out/prof_*.ncu-rep)"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]

lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        k = re.sub(r"<.*", "", row["Kernel Name"])[:80]
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    with open(f"profiles/{tag}_launches.txt", "w") as f:
        f.write(f"# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none ; one bench.py step (fwd+bwd+AdamW)\n")
        f.write(f"# total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare SHARES)\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.3f} ms {100 * t / tot:6.2f}%  x{c:5d}  {k}\n")
for rep in sorted(glob.glob(os.path.join(src, "prof_*.ncu-rep"))):
    name = os.path.basename(rep)[5:-8]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(f"profiles/{tag}_{name}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{name} -c 1 (one launch inside a bench.py step)\n")
        for i, h in enumerate(hdr):
            if h in KEYS or h in ("Kernel Name",):
                f.write(f"{h} = {vals[i]} {units[i]}\n")
# DRAM traffic per launch of each captured kernel -> profiles/traffic.json (read by bench.py's roofline.traffic)
import json
alias = {"attn_bwd2_kernel": "attn_bwd", "attn_bwd_kernel": "attn_bwd", "attn_fwd_kernel": "attn_fwd", "attn_fwd2_kernel": "attn_fwd",
         "gno_fwd_tc_kernel": "gno_fwd", "gno_bwd_tc_kernel": "gno_bwd", "gno_fwd_tc2_kernel": "gno_fwd", "gno_bwd_tc2_kernel": "gno_bwd", "gemm_tc_kernel": "linear_bwd_w", "gemm2_kernel": "linear_fwd", "gemm3_kernel": "linear_fwd", "node_mlp2_bwd_kernel": "node_mlp_bwd", "node_mlp2_fwd_kernel": "node_mlp_fwd", "knn_kernel": "knn_search"}
tp = "profiles/traffic.json"
traffic = json.load(open(tp)) if os.path.exists(tp) else {}
for rep in sorted(glob.glob(os.path.join(src, "prof_*.ncu-rep"))):
    name = os.path.basename(rep)[5:-8]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if key in hdr:
            i = hdr.index(key)
            tot += float(vals[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
    if name in alias:
        traffic[alias[name]] = tot
json.dump(traffic, open(tp, "w"), indent=1, sort_keys=True)
print(os.listdir("profiles"))
