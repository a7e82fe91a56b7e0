// gno_bf16.cu -- tcgen05 (BF16 operands, FP32 accumulate in TMEM) variant of the fused GNO forward.
#include "gno_common.cuh"
namespace gaot {
int gno_forward_bf16(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st) {
    (void)a; (void)ws; (void)ws_bytes; (void)out; (void)st;
    set_error("gno: precision=bf16 (tcgen05) path not built in this revision");
    return GAOT_ERR_UNSUPPORTED;
}
}  // namespace gaot
