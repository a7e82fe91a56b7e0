# usage (GPU box, N GPUs): bash profiles/multi_gpu.sh N   -> gpurun_out/scale_*.json
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_dp$N.json 2> gpurun_out/scale_dp$N.err; echo "dp rc=$?"; tail -2 gpurun_out/scale_dp$N.err
timeout 600 $TR --master-port 29522 bench.py --gpus $N --workload drivaerml8m --shard --steps 3 --warmup 3 > gpurun_out/scale_shard8m_$N.json 2> gpurun_out/scale_shard8m_$N.err; echo "shard rc=$?"; tail -3 gpurun_out/scale_shard8m_$N.err
python - <<PY
import json
for f in ("gpurun_out/scale_dp$N.json", "gpurun_out/scale_shard8m_$N.json"):
    try:
        d = json.load(open(f)); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["scaling"], d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
