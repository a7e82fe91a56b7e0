mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -3

timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
for f in ('bench_final',):
    d = json.load(open(f'gpurun_out/{f}.json'))
    print(f, round(d['value'], 2), round(d['ms_per_step'], 2), round(d['e2e']['value'], 2), d['clocks'], d.get('cpu_baseline', {}).get('value'))
    print('   ', {k: round(v['ms_per_step'], 2) for k, v in d['kernels'].items()})
PY
