# round 2 first GPU pass (1 GPU): smoke, GPU tests, bench with CUDA graphs on/off
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_tests.log; tail -5 gpurun_out/r02b_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02b_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-graph --no-extras --no-cpu-baseline > gpurun_out/r02b_bench_nograph.json 2> gpurun_out/r02b_bench_nograph.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r02b_bench', 'r02b_bench_nograph'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(d['value'], 2), round(d['ms_per_step'], 2), round(d['e2e']['value'], 2), d['gpu_launches'], d['clocks'])
        print('   ', {k: round(v['ms_per_step'], 2) for k, v in d['kernels'].items()})
        ex = d.get('extras', {})
        for k, v in ex.items():
            print('   ', k, json.dumps(v)[:600])
    except Exception as e:
        print(f, 'unreadable', e)
PY
