set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_gno.py -x -q 2>&1 | tail -15
GAOT_GNO_GEN=1 timeout 300 python tests/prof_ops.py gno 5 2>&1 | grep -v Warning | tee gpurun_out/gno_gen1.txt
timeout 300 python tests/prof_ops.py gno 5 2>&1 | grep -v Warning | tee gpurun_out/gno_gen2.txt
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gno.py 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_a.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gno_edges_per_s'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 3), round(v['frac'], 3))
PY
