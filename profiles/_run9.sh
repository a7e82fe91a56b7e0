set -x
mkdir -p gpurun_out
for e in 0 4 8 12; do echo "EMU=$e"; GAOT_ATTN_EMU=$e timeout 300 python tests/prof_ops.py attn 5 2>&1 | grep -E "attn_fwd|attn_bwd"; done | tee gpurun_out/attn_emu.txt
timeout 900 python -m pytest tests/test_gpu_attn.py tests/test_gpu_tblock.py -x -q 2>&1 | tail -5
GAOT_ATTN_EMU=12 timeout 900 python -m pytest tests/test_gpu_attn.py -x -q 2>&1 | tail -3
