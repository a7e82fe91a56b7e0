set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gno.py -x -q 2>&1 | tail -15
timeout 300 python tests/prof_ops.py gno 5 2>&1 | grep -v Warning | tee gpurun_out/gno_gen2c.txt
