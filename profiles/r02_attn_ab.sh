# attention A/B on one GPU: parity tests, then bench with 2 vs 3 backward score stages
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_tblock.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5
for st in 3 2; do
GAOT_ATTN_BWD_STAGES=$st timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02e_bench_stg$st.json 2> gpurun_out/r02e_bench_stg$st.err; echo "rc=$?"
done
python - <<'PY'
import json
for f in ('r02e_bench_stg3', 'r02e_bench_stg2'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(d['value'], 2), round(d['ms_per_step'], 2), round(d['e2e']['value'], 2), d['clocks'])
        print('   ', {k: round(v['ms_per_step'], 2) for k, v in d['kernels'].items()})
    except Exception as e:
        print(f, 'unreadable', e)
PY
