set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gno.py -x -q -k "moments or geo" 2>&1 | tail -5
for k in gno_fwd_tc2_kernelILi3 gno_bwd_tc2_kernelILi3 gno_bwd_tc2_kernelILi4; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/prof_$k python tests/prof_ops.py gno 2 > gpurun_out/p_$k.log 2>&1
done
ls -la gpurun_out
