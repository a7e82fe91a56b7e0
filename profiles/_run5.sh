set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err
timeout 600 python bench.py --workload crm500k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_crm.json 2> gpurun_out/bench_crm.err; tail -5 gpurun_out/bench_crm.err
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm2_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_gemm2_kernel python bench.py --profile-step --no-cpu-baseline > gpurun_out/p_gemm2.log 2>&1
python - <<'PY'
import json
for f in ('bench_c', 'bench_crm'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps(d['gno_edges_per_s'])[:1500])
        for k, v in d['kernels'].items(): print('  ', k, round(v['ms_per_step'], 3), round(v['frac'], 3), v['calls_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
