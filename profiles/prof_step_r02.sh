# usage (GPU box): bash profiles/prof_step_r02.sh   -> gpurun_out/{launches.csv, launches_crm.csv, prof_*.ncu-rep}
# ncu launch list of one headline step + one CRM-config step (fp32 GNO: radius / coalesce / geo kernels), then --set full captures
mkdir -p gpurun_out
B="python bench.py --profile-step --no-cpu-baseline --no-extras --no-graph"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/l.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_crm.csv $B --workload crm500k --gno-precision fp32 > gpurun_out/l_crm.log 2>&1; echo "crm launch list rc=$?"
for k in attn_bwd2_kernel attn_fwd2_kernel gemm3_kernel gno_fwd_tc2_kernel gno_bwd_tc2_kernel node_mlp2_bwd_kernel node_linear_bwd_kernel knn_kernel; do
  timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k $B > gpurun_out/p_$k.log 2>&1; echo "$k rc=$?"
done
for k in radius_warp_kernel sort_scatter_kernel geo_stats_kernel gno_fwd_fp32_kernel gno_bwd_fp32_kernel unique_flag_kernel; do
  timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k $B --workload crm500k --gno-precision fp32 > gpurun_out/p_$k.log 2>&1; echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
