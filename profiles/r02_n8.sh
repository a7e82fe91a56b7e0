N=${1:-8}; TAG=${2:-k}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29641 profiles/tools/a2a_probe.py > gpurun_out/r02${TAG}_a2a_probe_$N.log 2>&1; echo "probe rc=$?"; grep -v "^\[\|^\*\|OMP_NUM\|^$" gpurun_out/r02${TAG}_a2a_probe_$N.log | tail -6
timeout 400 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02${TAG}_shard_$N.json 2> gpurun_out/r02${TAG}_shard_$N.err; echo "shard bench rc=$?"; grep -v "^\[\|^\*\|OMP_NUM\|^$" gpurun_out/r02${TAG}_shard_$N.err | tail -4
GAOT_P2P_OPS=a2a timeout 200 $TR --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r02${TAG}_shard_a2aonly_$N.json 2> gpurun_out/r02${TAG}_shard_a2aonly_$N.err; echo "a2a-only rc=$?"
GAOT_A2A=nccl timeout 200 $TR --master-port 29524 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r02${TAG}_shard_nccl_$N.json 2> gpurun_out/r02${TAG}_shard_nccl_$N.err; echo "nccl rc=$?"
python - <<PY
import json
for f in ("shard", "shard_a2aonly", "shard_nccl"):
    try:
        d = json.load(open("gpurun_out/r02${TAG}_%s_$N.json" % f))
        print(f, d["n_gpus"], round(d["value"], 2), round(d["ms_per_step"], 2), d["scaling"], round(d["e2e"]["value"], 2), d["gpu_launches"], d["clocks"])
        for k, v in d.get("extras", {}).items():
            print("   ", k, json.dumps(v)[:260])
    except Exception as e:
        print(f, "unreadable", e)
PY
