set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_node_mlp.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -3 gpurun_out/bench_f.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_f.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 3), round(v['frac'], 3))
PY
