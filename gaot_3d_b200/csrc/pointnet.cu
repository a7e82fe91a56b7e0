// pointnet.cu -- PointNet-style geometric embedding (reference src/model/layers/geoembed.py:184-222, method='pointnet'):
//   pooled[q] = pool_{e : qry(e) = q}  relu(W2 relu(W1 (y[src(e)] - x[q]) + b1) + b2)          pool = max | mean
// The reference materialises [E,3], two [E,32] activations and scatters them; here one warp owns one query and walks
// its CSR row: lane c is hidden channel c, layer 1 is three FMAs per lane, layer 2 a 32-step shuffle broadcast of the
// layer-1 activations against the lane's row of W2 (registers), and the pool lives in a register.  Empty queries give 0
// (the reference's zero-initialised scatter); the Linear(32 -> out) behind the pool stays a node-level GEMM in the host.
// Backward (parameters only: coordinates carry no gradient): recompute per edge, route the pooled gradient (max: to the
// first maximal edge per channel, recorded by the forward; mean: 1/n to every edge), accumulate dW / db per lane in
// registers over all queries of the warp, reduce the warps of a CTA through shared memory in a fixed order, then a
// fixed-order reduction over CTAs: deterministic.
#include "common.cuh"

namespace gaot {

namespace pn { constexpr int H = 32, NPAR = H * 3 + H + H * H + H; constexpr int O_B1 = H * 3, O_W2 = O_B1 + H, O_B2 = O_W2 + H * H; }

__global__ void __launch_bounds__(128)
pointnet_fwd_kernel(const float* __restrict__ src_pos, const float* __restrict__ qry_pos, int64_t nq, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ csr_src, const float* __restrict__ params, int pooling, float* __restrict__ pooled,
                    int32_t* __restrict__ argmax) {
    using namespace pn;
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float w10 = params[lane * 3], w11 = params[lane * 3 + 1], w12 = params[lane * 3 + 2], b1 = params[O_B1 + lane];
    float w2[H];
#pragma unroll
    for (int j = 0; j < H; ++j) w2[j] = params[O_W2 + lane * H + j];
    const float b2 = params[O_B2 + lane];
    for (int64_t q = wid; q < nq; q += nw) {
        const int b = rowptr[q], e = rowptr[q + 1];
        const float qx = qry_pos[q * 3], qy = qry_pos[q * 3 + 1], qz = qry_pos[q * 3 + 2];
        float acc = 0.f;                               // relu outputs are >= 0 and an empty query pools to 0
        int best = -1;
        for (int p = b; p < e; ++p) {
            const int s = csr_src[p];
            const float dx = src_pos[(size_t)s * 3] - qx, dy = src_pos[(size_t)s * 3 + 1] - qy, dz = src_pos[(size_t)s * 3 + 2] - qz;
            const float h1 = fmaxf(fmaf(w12, dz, fmaf(w11, dy, fmaf(w10, dx, b1))), 0.f);
            float z = b2;
#pragma unroll
            for (int j = 0; j < H; ++j) z = fmaf(w2[j], __shfl_sync(0xffffffffu, h1, j), z);
            const float h2 = fmaxf(z, 0.f);
            if (pooling == 0) { if (best < 0 || h2 > acc) { acc = h2; best = p; } }
            else acc += h2;
        }
        if (pooling == 1 && e > b) acc /= (float)(e - b);
        pooled[q * H + lane] = acc;
        if (argmax) argmax[q * H + lane] = best;
    }
}

__global__ void __launch_bounds__(128)
pointnet_bwd_kernel(const float* __restrict__ src_pos, const float* __restrict__ qry_pos, int64_t nq, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ csr_src, const float* __restrict__ params, int pooling, const float* __restrict__ d_pooled,
                    const int32_t* __restrict__ argmax, float* __restrict__ partial) {
    using namespace pn;
    __shared__ float red[4][NPAR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float w10 = params[lane * 3], w11 = params[lane * 3 + 1], w12 = params[lane * 3 + 2], b1 = params[O_B1 + lane];
    float w2r[H], w2c[H];                              // row `lane` of W2 (forward) and column `lane` (d h1)
#pragma unroll
    for (int j = 0; j < H; ++j) { w2r[j] = params[O_W2 + lane * H + j]; w2c[j] = params[O_W2 + j * H + lane]; }
    const float b2 = params[O_B2 + lane];
    float dw2[H], dw1x = 0.f, dw1y = 0.f, dw1z = 0.f, db1 = 0.f, db2 = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) dw2[j] = 0.f;
    for (int64_t q = wid; q < nq; q += nw) {
        const int b = rowptr[q], e = rowptr[q + 1];
        if (e == b) continue;
        const float qx = qry_pos[q * 3], qy = qry_pos[q * 3 + 1], qz = qry_pos[q * 3 + 2];
        const float g = d_pooled[q * H + lane] * (pooling == 1 ? 1.0f / (float)(e - b) : 1.0f);
        const int am = pooling == 0 ? argmax[q * H + lane] : -1;
        for (int p = b; p < e; ++p) {
            // max pooling: only edges that are some channel's argmax contribute
            if (pooling == 0 && !__any_sync(0xffffffffu, am == p)) continue;
            const int s = csr_src[p];
            const float dx = src_pos[(size_t)s * 3] - qx, dy = src_pos[(size_t)s * 3 + 1] - qy, dz = src_pos[(size_t)s * 3 + 2] - qz;
            const float z1 = fmaf(w12, dz, fmaf(w11, dy, fmaf(w10, dx, b1)));
            const float h1 = fmaxf(z1, 0.f);
            float z = b2;
#pragma unroll
            for (int j = 0; j < H; ++j) z = fmaf(w2r[j], __shfl_sync(0xffffffffu, h1, j), z);
            float gz2 = (z > 0.f && (pooling == 1 || am == p)) ? g : 0.f;
            db2 += gz2;
            float dh1 = 0.f;
#pragma unroll
            for (int j = 0; j < H; ++j) {
                dw2[j] = fmaf(gz2, __shfl_sync(0xffffffffu, h1, j), dw2[j]);          // dW2[lane][j] += gz2_lane * h1_j
                dh1 = fmaf(w2c[j], __shfl_sync(0xffffffffu, gz2, j), dh1);            // dh1_lane = sum_c gz2_c W2[c][lane]
            }
            const float gz1 = z1 > 0.f ? dh1 : 0.f;
            dw1x = fmaf(gz1, dx, dw1x); dw1y = fmaf(gz1, dy, dw1y); dw1z = fmaf(gz1, dz, dw1z);
            db1 += gz1;
        }
    }
    float* r = red[warp];
    r[lane * 3] = dw1x; r[lane * 3 + 1] = dw1y; r[lane * 3 + 2] = dw1z;
    r[O_B1 + lane] = db1;
#pragma unroll
    for (int j = 0; j < H; ++j) r[O_W2 + lane * H + j] = dw2[j];
    r[O_B2 + lane] = db2;
    __syncthreads();
    for (int i = threadIdx.x; i < NPAR; i += blockDim.x)
        partial[(size_t)blockIdx.x * NPAR + i] = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
}

int gno_bwd_reduce(const float* partial, int nparts, int n_params, float* d_params, cudaStream_t st);   // gno_bwd.cu

}  // namespace gaot

using namespace gaot;

extern "C" {

size_t gaot_pointnet_workspace_bytes(void) { return align_up((size_t)4 * kNumSMs * pn::NPAR * sizeof(float)) + 256; }

int gaot_pointnet_forward(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq, const int32_t* rowptr,
                          const int32_t* csr_src, const float* params, int pooling, float* pooled, int32_t* argmax, void* stream) {
    (void)n_src;
    GAOT_CHECK_ARG(nq >= 0 && params && pooled && (pooling == 0 || pooling == 1), "pointnet_forward: bad arguments");
    GAOT_CHECK_ARG(pooling == 1 || argmax != nullptr, "pointnet_forward: max pooling needs the argmax buffer");
    if (nq == 0) return GAOT_OK;
    const int64_t want = (nq + 3) / 4;
    const int grid = (int)(want < 8 * kNumSMs ? want : 8 * kNumSMs);
    pointnet_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src_pos, qry_pos, nq, rowptr, csr_src, params, pooling, pooled, argmax);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

// d_params = [dW1 (32 x 3) | db1 (32) | dW2 (32 x 32) | db2 (32)]
int gaot_pointnet_backward(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq, const int32_t* rowptr,
                           const int32_t* csr_src, const float* params, int pooling, const float* d_pooled, const int32_t* argmax,
                           void* ws, size_t ws_bytes, float* d_params, void* stream) {
    (void)n_src;
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(nq >= 0 && params && d_pooled && d_params && (pooling == 0 || pooling == 1), "pointnet_backward: bad arguments");
    GAOT_CHECK_ARG(pooling == 1 || argmax != nullptr, "pointnet_backward: max pooling needs the argmax buffer");
    if (nq == 0) { GAOT_CUDA(cudaMemsetAsync(d_params, 0, pn::NPAR * sizeof(float), st)); return GAOT_OK; }
    const int64_t want = (nq + 3) / 4;
    const int grid = (int)(want < 4 * kNumSMs ? want : 4 * kNumSMs);
    Arena ar(ws, ws_bytes);
    float* partial = ar.take<float>((size_t)grid * pn::NPAR);
    if (!ar.ok()) { set_error("pointnet_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    pointnet_bwd_kernel<<<grid, 128, 0, st>>>(src_pos, qry_pos, nq, rowptr, csr_src, params, pooling, d_pooled, argmax, partial);
    GAOT_LAUNCH_CHECK();
    return gno_bwd_reduce(partial, grid, pn::NPAR, d_params, st);
}

}  // extern "C"
