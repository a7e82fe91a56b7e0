// node_linear.cu -- one-layer node MLP with a tiny input width: y[n, C] = x[n, K] W^T + b, K <= 16, C <= 64 (fp32).
// The encoder's lifting layer (reference src/model/layers/magno.py:421-424, applied at :540-545): Linear / Conv1d(k=1) from
// the raw point features (pos + normals: K = 6) to the lifting channels (C = 16..32) over EVERY physical point.
// As a library GEMM it is a bad shape: cuBLAS runs the forward as a 128x128 TF32/SGEMM tile with K padded 6 -> 16/32 and
// the weight gradient dW[C, K] = dy^T x as an `nt` GEMM whose whole parallelism is a split over the 10^6..10^7 rows
// (0.455 ms at 1M points in the sharded 8M step, profiles/r02d_trace_shard8m_8_rank0.txt) -- both are pure streaming
// problems: 4 (K + C) bytes per point forward, the same again backward.
//   forward : lane = output column (two columns per lane when C > 32); the warp walks rows, x[row, :] are broadcast loads,
//             W[c, :] and b[c] live in registers, the 4 C-byte output rows are written as full 128-byte lines.
//   backward: same mapping; every lane accumulates dW[c, 0..K) and db[c] over the warp's rows in registers (fp32), the block
//             reduces its warps through shared memory and writes ONE partial per block; a second launch sums the partials in
//             fixed order (deterministic, no atomics).  d x (only when the input needs it) = dy W, one row per thread.
#include "common.cuh"
#include <algorithm>

namespace gaot {

namespace nl {
constexpr int MAXK = 16, THREADS = 256, WARPS = THREADS / 32, ROWS_IT = 4;
}

template <int CPL>   // columns per lane: 1 (C <= 32) or 2 (C <= 64)
__global__ void __launch_bounds__(nl::THREADS)
node_linear_fwd_kernel(const float* __restrict__ x, int64_t n, int K, int C, const float* __restrict__ W,
                       const float* __restrict__ b, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    float w[CPL][nl::MAXK], bias[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        bias[j] = (b != nullptr && c < C) ? b[c] : 0.f;
#pragma unroll
        for (int k = 0; k < nl::MAXK; ++k) w[j][k] = (c < C && k < K) ? W[(size_t)c * K + k] : 0.f;
    }
    const int64_t warp_global = (int64_t)blockIdx.x * nl::WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * nl::WARPS;
    for (int64_t r0 = warp_global * nl::ROWS_IT; r0 < n; r0 += nwarps * nl::ROWS_IT) {
        float acc[nl::ROWS_IT][CPL];
#pragma unroll
        for (int i = 0; i < nl::ROWS_IT; ++i) {
            const int64_t r = min(r0 + i, n - 1);
            const float* xr = x + r * K;
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[i][j] = bias[j];
#pragma unroll
            for (int k = 0; k < nl::MAXK; ++k) {
                if (k < K) {
                    const float xv = __ldg(xr + k);      // same address in every lane: one broadcast transaction
#pragma unroll
                    for (int j = 0; j < CPL; ++j) acc[i][j] = fmaf(xv, w[j][k], acc[i][j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < nl::ROWS_IT; ++i) {
            if (r0 + i < n) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const int c = lane + 32 * j;
                    if (c < C) y[(r0 + i) * C + c] = acc[i][j];
                }
            }
        }
    }
}

// partial[block][c][K + 1]: columns 0..K-1 = dW[c, :], column K = db[c]
template <int CPL>
__global__ void __launch_bounds__(nl::THREADS)
node_linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t n, int K, int C,
                       float* __restrict__ partial) {
    __shared__ float red[nl::WARPS][64][nl::MAXK + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[CPL][nl::MAXK + 1];
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
        for (int k = 0; k <= nl::MAXK; ++k) acc[j][k] = 0.f;
    const int64_t warp_global = (int64_t)blockIdx.x * nl::WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * nl::WARPS;
    for (int64_t r0 = warp_global * nl::ROWS_IT; r0 < n; r0 += nwarps * nl::ROWS_IT) {
        float g[nl::ROWS_IT][CPL];
#pragma unroll
        for (int i = 0; i < nl::ROWS_IT; ++i)
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int c = lane + 32 * j;
                g[i][j] = (r0 + i < n && c < C) ? __ldg(dy + (r0 + i) * C + c) : 0.f;
            }
#pragma unroll
        for (int i = 0; i < nl::ROWS_IT; ++i) {
            const float* xr = x + min(r0 + i, n - 1) * K;
#pragma unroll
            for (int k = 0; k < nl::MAXK; ++k) {
                if (k < K) {
                    const float xv = __ldg(xr + k);
#pragma unroll
                    for (int j = 0; j < CPL; ++j) acc[j][k] = fmaf(g[i][j], xv, acc[j][k]);
                }
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[j][nl::MAXK] += g[i][j];
        }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
        for (int k = 0; k <= nl::MAXK; ++k) red[warp][lane + 32 * j][k] = acc[j][k];
    __syncthreads();
    // fixed-order sum over the block's warps; thread t handles entries t, t + THREADS, ... of the [C][K + 1] partial
    const int per = C * (K + 1);
    for (int e = threadIdx.x; e < per; e += nl::THREADS) {
        const int c = e / (K + 1), k = e - c * (K + 1);
        const int kk = (k == K) ? nl::MAXK : k;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < nl::WARPS; ++w) s += red[w][c][kk];
        partial[(size_t)blockIdx.x * per + e] = s;
    }
}

__global__ void __launch_bounds__(256)
node_linear_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int per, int K, float* __restrict__ dW,
                              float* __restrict__ db) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per) return;
    float s = 0.f;
    for (int bk = 0; bk < nblocks; ++bk) s += partial[(size_t)bk * per + e];       // fixed order
    const int c = e / (K + 1), k = e - c * (K + 1);
    if (k == K) { if (db) db[c] = s; }
    else dW[(size_t)c * K + k] = s;
}

// d x[n, K] = dy[n, C] W[C, K]: one row per thread (rarely needed: the lifting input is data)
__global__ void __launch_bounds__(256)
node_linear_bwd_x_kernel(const float* __restrict__ dy, int64_t n, int K, int C, const float* __restrict__ W, float* __restrict__ dx) {
    extern __shared__ float ws[];                        // W [C][K]
    for (int e = threadIdx.x; e < C * K; e += blockDim.x) ws[e] = W[e];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float acc[nl::MAXK];
#pragma unroll
    for (int k = 0; k < nl::MAXK; ++k) acc[k] = 0.f;
    for (int c = 0; c < C; ++c) {
        const float g = __ldg(dy + r * C + c);
#pragma unroll
        for (int k = 0; k < nl::MAXK; ++k) if (k < K) acc[k] = fmaf(g, ws[c * K + k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < nl::MAXK; ++k) if (k < K) dx[r * K + k] = acc[k];
}

static int nl_grid(int64_t n) {
    const int64_t want = (n + nl::WARPS * nl::ROWS_IT - 1) / (nl::WARPS * nl::ROWS_IT);
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, 4 * kNumSMs));      // persistent: 4 CTAs of 256 threads per SM
}

}  // namespace gaot

using namespace gaot;

extern "C" {

int gaot_node_linear_supported(int32_t k_in, int32_t c_out) { return (k_in >= 1 && k_in <= nl::MAXK && c_out >= 1 && c_out <= 64) ? 1 : 0; }

size_t gaot_node_linear_workspace_bytes(int32_t k_in, int32_t c_out) {
    return align_up((size_t)4 * kNumSMs * c_out * (k_in + 1) * sizeof(float));
}

int gaot_node_linear_forward(const float* x, int64_t n, int32_t k_in, int32_t c_out, const float* w, const float* b, float* y,
                             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(gaot_node_linear_supported(k_in, c_out), "node_linear: k_in must be 1..16 and c_out 1..64");
    GAOT_CHECK_ARG(n >= 0 && (n == 0 || (x && w && y)), "node_linear_forward: null pointer");
    if (n == 0) return GAOT_OK;
    GAOT_TIME_KERNEL("node_linear_fwd", st, (double)n * (k_in + c_out) * 4.0);
    if (c_out <= 32) node_linear_fwd_kernel<1><<<nl_grid(n), nl::THREADS, 0, st>>>(x, n, k_in, c_out, w, b, y);
    else node_linear_fwd_kernel<2><<<nl_grid(n), nl::THREADS, 0, st>>>(x, n, k_in, c_out, w, b, y);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_node_linear_backward(const float* x, const float* d_y, int64_t n, int32_t k_in, int32_t c_out, const float* w,
                              void* ws, size_t ws_bytes, float* d_x, float* d_w, float* d_b, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(gaot_node_linear_supported(k_in, c_out), "node_linear: k_in must be 1..16 and c_out 1..64");
    GAOT_CHECK_ARG(d_w != nullptr && (n == 0 || (x && d_y)), "node_linear_backward: null pointer");
    GAOT_CHECK_ARG(ws_bytes >= gaot_node_linear_workspace_bytes(k_in, c_out) && ws != nullptr, "node_linear_backward: workspace too small");
    const int per = c_out * (k_in + 1);
    if (n == 0) {
        GAOT_CUDA(cudaMemsetAsync(d_w, 0, (size_t)c_out * k_in * 4, st));
        if (d_b) GAOT_CUDA(cudaMemsetAsync(d_b, 0, (size_t)c_out * 4, st));
        return GAOT_OK;
    }
    GAOT_TIME_KERNEL("node_linear_bwd", st, (double)n * (k_in + c_out) * 4.0);
    const int grid = nl_grid(n);
    float* partial = (float*)ws;
    if (c_out <= 32) node_linear_bwd_kernel<1><<<grid, nl::THREADS, 0, st>>>(x, d_y, n, k_in, c_out, partial);
    else node_linear_bwd_kernel<2><<<grid, nl::THREADS, 0, st>>>(x, d_y, n, k_in, c_out, partial);
    GAOT_LAUNCH_CHECK();
    node_linear_bwd_reduce_kernel<<<(per + 255) / 256, 256, 0, st>>>(partial, grid, per, k_in, d_w, d_b);
    GAOT_LAUNCH_CHECK();
    if (d_x) {
        GAOT_CHECK_ARG(w != nullptr, "node_linear_backward: d_x needs the weight");
        node_linear_bwd_x_kernel<<<(unsigned)((n + 255) / 256), 256, (size_t)c_out * k_in * 4, st>>>(d_y, n, k_in, c_out, w, d_x);
        GAOT_LAUNCH_CHECK();
    }
    return GAOT_OK;
}

}  // extern "C"
