// gno.cu -- C ABI of the fused GNO (IntegralTransform) forward / backward.
#include <cstdlib>
#include "gno_common.cuh"

namespace gaot {
int gno_forward_fp32(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st);
size_t gno_forward_ws_bytes(int64_t E, int Cout);
int gno_backward_fp32(const GnoArgs& a, const float* d_out, void* ws, size_t ws_bytes,
                      float* d_params, float* d_f, cudaStream_t st);
size_t gno_backward_ws_bytes(int n_params);
int gno_forward_bf16(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st);
bool gno_backward_bf16_supported(const GnoArgs& a);
bool gno_forward_bf16_supported(const GnoArgs& a);
int gno_backward_bf16(const GnoArgs& a, const float* d_out, void* ws, size_t ws_bytes, float* d_params, float* d_f,
                      cudaStream_t st);
// second-generation tensor-core kernels (gno_tc2.cu)
bool gno_forward_tc2_supported(const GnoArgs& a);
int gno_forward_tc2(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st);
bool gno_backward_tc2_supported(const GnoArgs& a);
int gno_backward_tc2(const GnoArgs& a, const float* d_out, void* ws, size_t ws_bytes, float* d_params, float* d_f,
                     cudaStream_t st);
}  // namespace gaot

// GAOT_GNO_GEN=1 pins the first-generation tensor-core kernels (A/B timing on the GPU box); default: newest that fits
static bool use_gen2() {
    const char* e = getenv("GAOT_GNO_GEN");
    return !(e && e[0] == '1');
}

using namespace gaot;

static int fill(GnoArgs& a, const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq,
                const float* f_y, int32_t c_f, const int32_t* rowptr, const int32_t* csr_src,
                const int32_t* csr_qry, int64_t E, const gaot_mlp_desc* mlp, const float* params,
                int transform, int reduce) {
    GAOT_CHECK_ARG(E >= 0 && E < ((int64_t)1 << 31) - 256, "gno: bad E");
    GAOT_CHECK_ARG(nq >= 0 && nq < ((int64_t)1 << 31) && n_src >= 0 && n_src < ((int64_t)1 << 31), "gno: bad sizes");
    GAOT_CHECK_ARG(transform >= 0 && transform <= 3, "gno: bad transform %d", transform);
    GAOT_CHECK_ARG(reduce == 0 || reduce == 1, "gno: bad reduce %d", reduce);
    GAOT_CHECK_ARG((transform == 3) == (f_y == nullptr), "gno: f_y must be given exactly when transform != 3");
    GAOT_CHECK_ARG(params != nullptr, "gno: params is null");
    memset(&a, 0, sizeof(a));
    int rc = gno_fill_args(a, mlp, f_y ? c_f : 0, transform);
    if (rc) return rc;
    a.y_pos = y_pos; a.x_pos = x_pos; a.f_y = f_y; a.c_f = f_y ? c_f : 0;
    a.rowptr = rowptr; a.csr_src = csr_src; a.csr_qry = csr_qry; a.params = params;
    a.E = (int32_t)E; a.nq = (int32_t)nq; a.n_src = (int32_t)n_src;
    a.transform = transform; a.reduce = reduce;
    a.ntiles = (int32_t)((E + 127) / 128);
    return GAOT_OK;
}

extern "C" {

size_t gaot_gno_workspace_bytes(int64_t E, int64_t nq, const gaot_mlp_desc* mlp) {
    (void)nq;
    int np = 0, cout = 64;
    if (mlp && mlp->n_layers >= 1 && mlp->n_layers <= GNO_MAX_LAYERS) {
        for (int l = 0; l < mlp->n_layers; ++l) np += mlp->dims[l] * mlp->dims[l + 1] + mlp->dims[l + 1];
        cout = mlp->dims[mlp->n_layers];
    }
    const size_t f = gno_forward_ws_bytes(E, cout), b = gno_backward_ws_bytes(np);
    return (f > b ? f : b) + 1024;
}

int gaot_gno_forward(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq, const float* f_y,
                     int32_t c_f, const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry,
                     int64_t E, const gaot_mlp_desc* mlp, const float* params, int transform, int reduce,
                     int precision, void* ws, size_t ws_bytes, float* out, void* stream) {
    return gaot_gno_forward_weighted(y_pos, n_src, x_pos, nq, f_y, c_f, rowptr, csr_src, csr_qry, E, mlp, params, transform, reduce,
                                     precision, nullptr, ws, ws_bytes, out, stream);
}

int gaot_gno_forward_weighted(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq, const float* f_y,
                              int32_t c_f, const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry,
                              int64_t E, const gaot_mlp_desc* mlp, const float* params, int transform, int reduce,
                              int precision, const float* edge_w, void* ws, size_t ws_bytes, float* out, void* stream) {
    GnoArgs a;
    int rc = fill(a, y_pos, n_src, x_pos, nq, f_y, c_f, rowptr, csr_src, csr_qry, E, mlp, params, transform, reduce);
    if (rc) return rc;
    GAOT_CHECK_ARG(edge_w == nullptr || reduce == 1, "gno: per-edge weights go with reduce = sum (integral_transform.py:165)");
    a.edge_w = edge_w;
    // per-edge weights: FP32 kernels or the second-generation tensor-core kernels (the first generation has no weight path)
    if (precision == 0) return gno_forward_fp32(a, ws, ws_bytes, out, (cudaStream_t)stream);
    if (edge_w && !(use_gen2() && gno_forward_tc2_supported(a))) return gno_forward_fp32(a, ws, ws_bytes, out, (cudaStream_t)stream);
    if (precision == 1) {
        if (use_gen2() && gno_forward_tc2_supported(a)) return gno_forward_tc2(a, ws, ws_bytes, out, (cudaStream_t)stream);
        if (gno_forward_bf16_supported(a)) return gno_forward_bf16(a, ws, ws_bytes, out, (cudaStream_t)stream);
        return gno_forward_fp32(a, ws, ws_bytes, out, (cudaStream_t)stream);     // outside the tcgen05 envelope: FP32 kernel
    }
    set_error("gno: unknown precision %d", precision);
    return GAOT_ERR_INVALID;
}

int gaot_gno_backward(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq, const float* f_y,
                      int32_t c_f, const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry,
                      int64_t E, const gaot_mlp_desc* mlp, const float* params, int transform, int reduce,
                      int precision, const float* d_out, void* ws, size_t ws_bytes, float* d_params,
                      float* d_f_y, void* stream) {
    return gaot_gno_backward_weighted(y_pos, n_src, x_pos, nq, f_y, c_f, rowptr, csr_src, csr_qry, E, mlp, params, transform, reduce,
                                      precision, nullptr, d_out, ws, ws_bytes, d_params, d_f_y, nullptr, stream);
}

int gaot_gno_backward_weighted(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq, const float* f_y,
                               int32_t c_f, const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry,
                               int64_t E, const gaot_mlp_desc* mlp, const float* params, int transform, int reduce,
                               int precision, const float* edge_w, const float* d_out, void* ws, size_t ws_bytes,
                               float* d_params, float* d_f_y, float* d_edge_w, void* stream) {
    GnoArgs a;
    int rc = fill(a, y_pos, n_src, x_pos, nq, f_y, c_f, rowptr, csr_src, csr_qry, E, mlp, params, transform, reduce);
    if (rc) return rc;
    GAOT_CHECK_ARG(d_out != nullptr && d_params != nullptr, "gno_backward: null gradient buffers");
    GAOT_CHECK_ARG(edge_w != nullptr || d_edge_w == nullptr, "gno_backward: d_edge_w without edge_w");
    GAOT_CHECK_ARG(edge_w == nullptr || reduce == 1, "gno: per-edge weights go with reduce = sum");
    a.edge_w = edge_w; a.d_edge_w = d_edge_w;
    if (edge_w && !(precision == 1 && use_gen2() && gno_backward_tc2_supported(a)))
        return gno_backward_fp32(a, d_out, ws, ws_bytes, d_params, d_f_y, (cudaStream_t)stream);
    // precision 1: tensor-core backward when the MLP fits its envelope, otherwise the FP32 recompute
    if (precision == 1 && use_gen2() && gno_backward_tc2_supported(a))
        return gno_backward_tc2(a, d_out, ws, ws_bytes, d_params, d_f_y, (cudaStream_t)stream);
    if (precision == 1 && gno_backward_bf16_supported(a))
        return gno_backward_bf16(a, d_out, ws, ws_bytes, d_params, d_f_y, (cudaStream_t)stream);
    return gno_backward_fp32(a, d_out, ws, ws_bytes, d_params, d_f_y, (cudaStream_t)stream);
}

}  // extern "C"
