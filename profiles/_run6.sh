set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_tblock.py -x -q 2>&1 | tail -15
GAOT_GEMM_GEN=2 timeout 300 python tests/prof_ops.py dense 20 2>&1 | grep -v Warning | tee gpurun_out/dense_gen2.txt
timeout 300 python tests/prof_ops.py dense 20 2>&1 | grep -v Warning | tee gpurun_out/dense_gen3.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_d.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 3), round(v['frac'], 3))
PY
