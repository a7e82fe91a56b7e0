# usage (GPU box, N GPUs): bash profiles/r02_multi.sh N TAG  -> gpurun_out/r02<TAG>_shard_N.json (+ shard test log)
N=${1:-2}; TAG=${2:-c}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$3" != "notest" ]; then
timeout 600 $TR --master-port 29611 tests/shard_worker.py > gpurun_out/r02${TAG}_shardtest_$N.log 2>&1; echo "shard test rc=$?"; grep -v "^\[" gpurun_out/r02${TAG}_shardtest_$N.log | tail -14
fi
timeout 900 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02${TAG}_shard_$N.json 2> gpurun_out/r02${TAG}_shard_$N.err; echo "shard bench rc=$?"; tail -4 gpurun_out/r02${TAG}_shard_$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02${TAG}_shard_$N.json"))
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["scaling"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
    print({k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
    for k, v in d.get("extras", {}).items():
        print(k, json.dumps(v)[:700])
except Exception as e:
    print("unreadable", e)
PY
if [ "$4" == "trace" ]; then
timeout 150 $TR --master-port 29533 profiles/tools/shard_trace.py --out gpurun_out/r02${TAG}_trace > gpurun_out/r02${TAG}_trace_$N.log 2>&1; echo "trace rc=$?"; grep -v "^\[" gpurun_out/r02${TAG}_trace_$N.log | tail -45
fi
