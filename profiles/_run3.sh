set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gno.py -x -q 2>&1 | tail -15
timeout 300 python tests/prof_ops.py gno 5 2>&1 | grep -v Warning | tee gpurun_out/gno_gen2b.txt
for k in gno_fwd_tc2_kernelILi3 gno_bwd_tc2_kernelILi3; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/prof_$k python tests/prof_ops.py gno 2 > gpurun_out/p_$k.log 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_b.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gno_edges_per_s'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 3), round(v['frac'], 3))
PY
