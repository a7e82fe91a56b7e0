// gno_common.cuh -- shared definitions of the fused GNO (IntegralTransform) kernels.
#pragma once
#include "common.cuh"

namespace gaot {

constexpr int GNO_MAX_LAYERS = 6;
constexpr int GNO_MAXW = 64;        // max hidden / output width
constexpr int GNO_MAXIN = 72;       // max first-layer input width (2*3 coords + c_f)

struct GnoArgs {
    const float* y_pos; const float* x_pos; const float* f_y;
    const int32_t* rowptr; const int32_t* csr_src; const int32_t* csr_qry;
    const float* params;
    const float* edge_w;                // optional per-edge weight in CSR order (attentional integral: replaces 1/count, reduce = sum)
    float* d_edge_w;                    // optional: d loss / d edge_w (backward; zero-initialised by the launcher)
    int32_t E, nq, n_src, c_f;
    int32_t n_layers;
    int32_t dims[GNO_MAX_LAYERS + 1];
    int32_t w_off[GNO_MAX_LAYERS];      // offset of W_l in the flat params buffer
    int32_t b_off[GNO_MAX_LAYERS];
    int32_t n_params;
    int32_t transform;                  // 0 linear, 1 nonlinear, 2 nonlinear_kernelonly, 3 no f_y
    int32_t reduce;                     // 0 mean, 1 sum
    int32_t ntiles;
};

inline int gno_fill_args(GnoArgs& a, const gaot_mlp_desc* mlp, int c_f, int transform) {
    if (!mlp || mlp->n_layers < 1 || mlp->n_layers > GNO_MAX_LAYERS) {
        set_error("gno: n_layers must be in [1,%d]", GNO_MAX_LAYERS); return GAOT_ERR_UNSUPPORTED;
    }
    a.n_layers = mlp->n_layers;
    int off = 0;
    for (int l = 0; l <= mlp->n_layers; ++l) a.dims[l] = mlp->dims[l];
    for (int l = 0; l < mlp->n_layers; ++l) {
        a.w_off[l] = off; off += a.dims[l] * a.dims[l + 1];
        a.b_off[l] = off; off += a.dims[l + 1];
        if (a.dims[l + 1] < 1 || a.dims[l + 1] > GNO_MAXW) {
            set_error("gno: layer width %d outside [1,%d]", a.dims[l + 1], GNO_MAXW); return GAOT_ERR_UNSUPPORTED;
        }
    }
    a.n_params = off;
    const int want_in = 6 + ((transform == 1 || transform == 2) ? c_f : 0);
    if (a.dims[0] != want_in) {
        set_error("gno: MLP input width %d does not match 2*coord_dim(+c_f) = %d", a.dims[0], want_in);
        return GAOT_ERR_INVALID;
    }
    if (a.dims[0] > GNO_MAXIN) { set_error("gno: input width %d > %d", a.dims[0], GNO_MAXIN); return GAOT_ERR_UNSUPPORTED; }
    if (transform == 0 || transform == 1) {
        if (a.dims[a.n_layers] != c_f) {
            set_error("gno: kernel MLP output width %d must equal feature width %d", a.dims[a.n_layers], c_f);
            return GAOT_ERR_INVALID;
        }
    }
    return GAOT_OK;
}

__device__ __forceinline__ float gelu_exact(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

}  // namespace gaot
