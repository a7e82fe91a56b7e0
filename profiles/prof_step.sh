# usage (GPU box): bash profiles/prof_step.sh   -> gpurun_out/{launches.csv, prof_*.ncu-rep, bench lines}
set -x
mkdir -p gpurun_out
B="python bench.py --profile-step --no-cpu-baseline"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/l.log 2>&1
for k in attn_bwd2_kernel attn_fwd2_kernel gemm3_kernel gno_fwd_tc2_kernel gno_bwd_tc2_kernel node_mlp2_bwd_kernel knn_kernel; do
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k $B > gpurun_out/p_$k.log 2>&1
done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/b.err
python bench.py --steps 10 --warmup 3 --gno-precision fp32 --no-cpu-baseline > gpurun_out/bench_gno_fp32.json 2>> gpurun_out/b.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/b.err
ls -la gpurun_out
