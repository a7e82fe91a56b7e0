N=8; TAG=d
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 170 $TR --master-port 29533 profiles/tools/shard_trace.py --out gpurun_out/r02${TAG}_trace > gpurun_out/r02${TAG}_trace_$N.log 2>&1; echo "trace rc=$?"; grep -v "^\[" gpurun_out/r02${TAG}_trace_$N.log | tail -45
