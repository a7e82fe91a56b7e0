N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload drivaerml8m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_shard8m_1.json 2> gpurun_out/scale_shard8m_1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload drivaerml8m --shard --steps 3 --warmup 3 > gpurun_out/scale_shard8m_$N.json 2> gpurun_out/scale_shard8m_$N.err
fi
tail -2 gpurun_out/scale_shard8m_$N.err
python -c "
import json; d=json.load(open('gpurun_out/scale_shard8m_$N.json')); print('N=$N', d['ms_per_step'], d['value'], d['e2e']['value']); print({k: round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"
