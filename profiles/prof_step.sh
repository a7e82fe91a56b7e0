set -x
B="python bench.py --profile-step --no-cpu-baseline --gno-precision bf16"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/l.log 2>&1
for k in attn_bwd_kernel attn_fwd_kernel gemm_tc_kernel gno_fwd_tc_kernel gno_bwd_tc_kernel; do
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k $B > gpurun_out/p_$k.log 2>&1
done
python bench.py --steps 10 --warmup 3 --gno-precision bf16 > gpurun_out/r01c_bench_bf16.json 2> gpurun_out/b.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_bench_fp32.json 2>> gpurun_out/b.err
ls -la gpurun_out
