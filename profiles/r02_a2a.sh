N=${1:-2}; TAG=${2:-i}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29641 profiles/tools/a2a_probe.py > gpurun_out/r02${TAG}_a2a_probe_$N.log 2>&1; echo "probe rc=$?"; grep -v "^\[\|^\*\|OMP_NUM\|^$" gpurun_out/r02${TAG}_a2a_probe_$N.log | tail -12
timeout 300 $TR --master-port 29611 tests/shard_worker.py > gpurun_out/r02${TAG}_shardtest_$N.log 2>&1; echo "shard test rc=$?"; grep -v "^\[\|^\*\|OMP_NUM\|^$" gpurun_out/r02${TAG}_shardtest_$N.log | tail -14
timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02${TAG}_shard_$N.json 2> gpurun_out/r02${TAG}_shard_$N.err; echo "shard bench rc=$?"; grep -v "^\[\|^\*\|OMP_NUM\|^$" gpurun_out/r02${TAG}_shard_$N.err | tail -4
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02${TAG}_shard_$N.json"))
    print(d["n_gpus"], round(d["value"], 2), round(d["ms_per_step"], 2), d["scaling"], round(d["e2e"]["value"], 2), d["gpu_launches"], d["clocks"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
    for k, v in d.get("extras", {}).items():
        print(k, json.dumps(v)[:300])
except Exception as e:
    print("unreadable", e)
PY
