# attention prefetch / stage variants: timing (prof_ops) then parity tests with the winning env
mkdir -p gpurun_out
run() { echo "$*"; env "$@" timeout 120 python profiles/tools/prof_ops.py attn 10 2>&1 | grep "attn_bwd\|attn_fwd"; }
{
run GAOT_X=0
run GAOT_ATTN_BWD_PREF=1 GAOT_ATTN_FWD_PREF=1
run GAOT_ATTN_BWD_PREF=1 GAOT_ATTN_BWD_STAGES=3 GAOT_ATTN_FWD_PREF=1 GAOT_ATTN_EMU=8
run GAOT_ATTN_BWD_STAGES=3 GAOT_ATTN_FWD_PREF=1 GAOT_ATTN_EMU=0
} | tee gpurun_out/r02f_attn_pref.txt
GAOT_ATTN_BWD_PREF=1 GAOT_ATTN_BWD_STAGES=3 GAOT_ATTN_FWD_PREF=1 timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_tblock.py -x -q 2>&1 | tail -3
