set -x
mkdir -p gpurun_out
for skip in 1 3; do
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm3_kernel --launch-skip $skip -c 1 -f -o gpurun_out/prof_gemm3_$skip python bench.py --profile-step --no-cpu-baseline > gpurun_out/p_gemm3.log 2>&1
done
ls -la gpurun_out | grep gemm3
