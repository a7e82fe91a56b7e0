set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_tblock.py tests/test_gpu_model.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -3 gpurun_out/bench_e.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_e.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 3), round(v['frac'], 3))
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/l.log 2>&1
