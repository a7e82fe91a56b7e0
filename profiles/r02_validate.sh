mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02g_tests.log; tail -8 gpurun_out/r02g_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02g_bench.json'))
print(round(d['value'], 2), round(d['ms_per_step'], 2), round(d['e2e']['value'], 2), d['gpu_launches'], d['clocks'])
print('   ', {k: round(v['ms_per_step'], 2) for k, v in d['kernels'].items()})
for k, v in d.get('extras', {}).items():
    if k == 'cublas_bars':
        for n, r in v.items():
            print('   ', n, {kk: (round(vv['ours_ms'] * 1e3, 1), round(vv['cublas_ms'] * 1e3, 1)) for kk, vv in r.items() if isinstance(vv, dict) and 'ours_ms' in vv})
    elif k != 'graph_sweep':
        print('   ', k, json.dumps(v)[:400])
print(d.get('cpu_baseline'))
PY
