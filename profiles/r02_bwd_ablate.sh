# backward attention ablations (timing only): which resource does the step time follow?  Needs a scratch build that instantiates
# attn_bwd2_kernel<16,false,false,0,2,DBG> behind GAOT_ATTN_BWD_DBG (as commit 1001e31 did); the product library no longer does.
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 8 16 11 27 31; do
  echo "DBG=$dbg"; GAOT_ATTN_BWD_DBG=$dbg timeout 120 python profiles/tools/prof_ops.py attn 10 2>&1 | grep "attn_bwd\|attn_fwd"
done | tee gpurun_out/r02e_bwd_ablate.txt
