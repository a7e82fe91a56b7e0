// tblock.cu -- row-wise kernels of the latent TransformerBlock (reference src/model/layers/attn.py):
// RMSNorm forward/backward (:167-178), SwiGLU gate silu(w1 x) * w3 x forward/backward (:163), bias-gradient
// column sums, fp32 -> bf16 operand casts.  All HBM-bound streaming kernels: one warp per token row,
// 128-bit loads, bf16 outputs written in the layout the tcgen05 dense kernel (dense.cu) reads as operands.
#include "common.cuh"
#include "tc05.cuh"
#include <algorithm>

namespace gaot {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- cast ----
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = tc::pack_bf16(a.x, a.y); o.y = tc::pack_bf16(a.z, a.w); o.z = tc::pack_bf16(b.x, b.y); o.w = tc::pack_bf16(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
}

// several casts in one launch (the weights of one transformer block): blockIdx.y selects the segment
struct CastBatch { const float* src[8]; bf16* dst[8]; int64_t n8[8]; };
__global__ void __launch_bounds__(256)
cast_bf16_batch_kernel(const CastBatch cb) {
    const int s = blockIdx.y;
    const int64_t n8 = cb.n8[s];
    const float4* src = reinterpret_cast<const float4*>(cb.src[s]);
    uint4* dst = reinterpret_cast<uint4*>(cb.dst[s]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
        uint4 o;
        o.x = tc::pack_bf16(a.x, a.y); o.y = tc::pack_bf16(a.z, a.w); o.z = tc::pack_bf16(b.x, b.y); o.w = tc::pack_bf16(b.z, b.w);
        dst[i] = o;
    }
}

// ---- RMSNorm: y = x * rsqrt(mean(x^2) + eps) * w   (fp32 statistics, reference attn.py:175-178) ----
// NV = H / 128 float4 per lane (H = 128 * NV)
template <int NV>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t M, float eps,
                   bf16* __restrict__ y_bf16, float* __restrict__ y_f32, float* __restrict__ rstd) {
    constexpr int H = NV * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * H);
    float4 v[NV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = __ldg(xr + i * 32 + lane);
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss * (1.0f / H) + eps);
    if (lane == 0 && rstd) rstd[row] = rs;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
        const float4 o = make_float4(v[i].x * rs * ww.x, v[i].y * rs * ww.y, v[i].z * rs * ww.z, v[i].w * rs * ww.w);
        if (y_f32) reinterpret_cast<float4*>(y_f32 + row * H)[i * 32 + lane] = o;
        if (y_bf16) {
            uint2 p;
            p.x = tc::pack_bf16(o.x, o.y); p.y = tc::pack_bf16(o.z, o.w);
            reinterpret_cast<uint2*>(y_bf16 + row * H)[i * 32 + lane] = p;
        }
    }
}

// backward: xh = x * rstd, g = dy * w, dx = rstd * (g - xh * mean(g * xh)) (+ dres), dw = sum_rows dy * xh.
// Each CTA (8 warps) walks rows with a grid stride and keeps its dw partial in registers -> dw_part[blockIdx][H];
// rmsnorm_dw_reduce_kernel sums the partials in fixed order (deterministic).
template <int NV>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ rstd,
                   const float* __restrict__ w, const float* __restrict__ dres, int64_t M,
                   float* __restrict__ dx, float* __restrict__ dw_part) {
    constexpr int H = NV * 128;
    __shared__ float4 red[8][NV * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 ww[NV], dwacc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        ww[i] = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
        dwacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < M; row += (int64_t)gridDim.x * 8) {
        const float rs = __ldg(rstd + row);
        const float4* xr = reinterpret_cast<const float4*>(x + row * H);
        const float4* gr = reinterpret_cast<const float4*>(dy + row * H);
        float4 xh[NV], g[NV];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 xv = __ldg(xr + i * 32 + lane), dv = __ldg(gr + i * 32 + lane);
            xh[i] = make_float4(xv.x * rs, xv.y * rs, xv.z * rs, xv.w * rs);
            dwacc[i].x += dv.x * xh[i].x; dwacc[i].y += dv.y * xh[i].y; dwacc[i].z += dv.z * xh[i].z; dwacc[i].w += dv.w * xh[i].w;
            g[i] = make_float4(dv.x * ww[i].x, dv.y * ww[i].y, dv.z * ww[i].z, dv.w * ww[i].w);
            dot += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        }
        dot = warp_sum(dot) * (1.0f / H);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float4 o = make_float4(rs * (g[i].x - xh[i].x * dot), rs * (g[i].y - xh[i].y * dot),
                                   rs * (g[i].z - xh[i].z * dot), rs * (g[i].w - xh[i].w * dot));
            if (dres) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(dres + row * H) + i * 32 + lane);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            reinterpret_cast<float4*>(dx + row * H)[i * 32 + lane] = o;
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i * 32 + lane] = dwacc[i];
    __syncthreads();
    for (int j = threadIdx.x; j < NV * 32; j += 256) {
        float4 s = red[0][j];
#pragma unroll
        for (int k = 1; k < 8; ++k) { const float4 t = red[k][j]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        reinterpret_cast<float4*>(dw_part + (size_t)blockIdx.x * H)[j] = s;
    }
}

// out[c] = sum_p part[p][c]   (fixed order: 8 interleaved part lanes, then lanes 0..7).  CTA = 32 columns x 8 part lanes.
__global__ void __launch_bounds__(256)
colsum_reduce_kernel(const float* __restrict__ part, int nparts, int64_t N, float* __restrict__ out) {
    __shared__ float red[8][32];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + cl;
    float s = 0.f;
    if (c < N)
        for (int p = pl; p < nparts; p += 8) s += part[(size_t)p * N + c];
    red[pl][cl] = s;
    __syncthreads();
    if (pl == 0 && c < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s += red[k][cl];
        out[c] = s;
    }
}

// column sums of x [M,N] (bias gradients): CTA (256 threads = 64 float4 columns x 4 row lanes) over a row slab
__global__ void __launch_bounds__(256)
colsum_part_kernel(const float* __restrict__ x, int64_t M, int64_t N, int64_t rows_per_cta, float* __restrict__ part) {
    __shared__ float4 red[4][64];
    const int c4 = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int64_t col = ((int64_t)blockIdx.x * 64 + c4) * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < N)
        for (int64_t r = r0 + rl; r < r1; r += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * N + col));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    red[rl][c4] = s;
    __syncthreads();
    if (rl == 0 && col < N) {
#pragma unroll
        for (int k = 1; k < 4; ++k) { const float4 t = red[k][c4]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        *reinterpret_cast<float4*>(part + (size_t)blockIdx.y * N + col) = s;
    }
}

// ---- SwiGLU gate on the fused [gate | up] projection output GU [M, 2F] (bf16): a = silu(g) * u ----
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = tc::pack_bf16(f[0], f[1]); o.y = tc::pack_bf16(f[2], f[3]); o.z = tc::pack_bf16(f[4], f[5]); o.w = tc::pack_bf16(f[6], f[7]);
    return o;
}

__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const bf16* __restrict__ gu, int64_t M, int F, bf16* __restrict__ a) {
    const int f8 = F >> 3;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * f8) return;
    const int64_t row = idx / f8;
    const int c = (int)(idx % f8) * 8;
    float g[8], u[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gu + row * 2 * F + c)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(gu + row * 2 * F + F + c)), u);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = g[i] / (1.0f + __expf(-g[i])) * u[i];
    *reinterpret_cast<uint4*>(a + row * F + c) = pack8(o);
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const bf16* __restrict__ da, const bf16* __restrict__ gu, int64_t M, int F, bf16* __restrict__ dgu) {
    const int f8 = F >> 3;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * f8) return;
    const int64_t row = idx / f8;
    const int c = (int)(idx % f8) * 8;
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gu + row * 2 * F + c)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(gu + row * 2 * F + F + c)), u);
    unpack8(__ldg(reinterpret_cast<const uint4*>(da + row * F + c)), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float sg = 1.0f / (1.0f + __expf(-g[i]));
        const float silu = g[i] * sg;
        du[i] = d[i] * silu;
        dg[i] = d[i] * u[i] * sg * (1.0f + g[i] * (1.0f - sg));
    }
    *reinterpret_cast<uint4*>(dgu + row * 2 * F + c) = pack8(dg);
    *reinterpret_cast<uint4*>(dgu + row * 2 * F + F + c) = pack8(du);
}

static inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }
constexpr int kNormBwdCtas = 2 * kNumSMs;

}  // namespace gaot

using namespace gaot;

extern "C" {

int gaot_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
    GAOT_CHECK_ARG(n % 8 == 0, "cast_bf16: n must be a multiple of 8");
    if (n == 0) return GAOT_OK;
    cast_bf16_kernel<<<nblk(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n / 8);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_cast_bf16_batch(const float* const* src, void* const* dst, const int64_t* n, int32_t count, void* stream) {
    GAOT_CHECK_ARG(count >= 0 && count <= 8, "cast_bf16_batch: at most 8 segments per call");
    if (count == 0) return GAOT_OK;
    CastBatch cb;
    int64_t nmax = 0;
    for (int i = 0; i < count; ++i) {
        GAOT_CHECK_ARG(n[i] % 8 == 0, "cast_bf16_batch: segment lengths must be multiples of 8");
        cb.src[i] = src[i]; cb.dst[i] = (bf16*)dst[i]; cb.n8[i] = n[i] / 8;
        nmax = cb.n8[i] > nmax ? cb.n8[i] : nmax;
    }
    if (nmax == 0) return GAOT_OK;
    const unsigned gx = (unsigned)std::min<int64_t>((nmax + 255) / 256, 64);
    cast_bf16_batch_kernel<<<dim3(gx, (unsigned)count), 256, 0, (cudaStream_t)stream>>>(cb);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_rmsnorm_forward(const float* x, const float* w, int64_t M, int32_t H, float eps,
                         void* y_bf16, float* y_f32, float* rstd, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (H % 128 != 0 || H < 128 || H > 1024) { set_error("rmsnorm: hidden size %d unsupported (multiple of 128, <= 1024)", H); return GAOT_ERR_UNSUPPORTED; }
    if (M == 0) return GAOT_OK;
    GAOT_TIME_KERNEL("rmsnorm_fwd", st, (double)M * H * (4.0 + (y_bf16 ? 2.0 : 0.0) + (y_f32 ? 4.0 : 0.0)));
    const unsigned grid = nblk(M, 8);
#define GAOT_RN(NV) case NV: rmsnorm_fwd_kernel<NV><<<grid, 256, 0, st>>>(x, w, M, eps, (bf16*)y_bf16, y_f32, rstd); break;
    switch (H / 128) { GAOT_RN(1) GAOT_RN(2) GAOT_RN(3) GAOT_RN(4) GAOT_RN(5) GAOT_RN(6) GAOT_RN(7) GAOT_RN(8) }
#undef GAOT_RN
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

size_t gaot_rmsnorm_backward_workspace_bytes(int32_t H) { return align_up((size_t)kNormBwdCtas * H * sizeof(float)); }

int gaot_rmsnorm_backward(const float* dy, const float* x, const float* rstd, const float* w, const float* dres,
                          int64_t M, int32_t H, float* dx, float* dw, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (H % 128 != 0 || H < 128 || H > 1024) { set_error("rmsnorm: hidden size %d unsupported (multiple of 128, <= 1024)", H); return GAOT_ERR_UNSUPPORTED; }
    if (ws_bytes < gaot_rmsnorm_backward_workspace_bytes(H)) { set_error("rmsnorm_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    float* part = (float*)ws;
    const unsigned grid = (unsigned)std::min<int64_t>(kNormBwdCtas, std::max<int64_t>(1, (M + 7) / 8));
    {
        GAOT_TIME_KERNEL("rmsnorm_bwd", st, (double)M * H * (12.0 + (dres ? 4.0 : 0.0)));
#define GAOT_RN(NV) case NV: rmsnorm_bwd_kernel<NV><<<grid, 256, 0, st>>>(dy, x, rstd, w, dres, M, dx, part); break;
        switch (H / 128) { GAOT_RN(1) GAOT_RN(2) GAOT_RN(3) GAOT_RN(4) GAOT_RN(5) GAOT_RN(6) GAOT_RN(7) GAOT_RN(8) }
#undef GAOT_RN
        GAOT_LAUNCH_CHECK();
    }
    colsum_reduce_kernel<<<nblk(H, 32), 256, 0, st>>>(part, (int)grid, H, dw);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

size_t gaot_colsum_workspace_bytes(int64_t M, int64_t N) { return align_up((size_t)std::min<int64_t>(256, (M + 63) / 64) * N * sizeof(float) + 256); }

int gaot_colsum(const float* x, int64_t M, int64_t N, float* out, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(N % 4 == 0 && N > 0, "colsum: N must be a positive multiple of 4");
    if (ws_bytes < gaot_colsum_workspace_bytes(M, N)) { set_error("colsum: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(256, (M + 63) / 64));
    const int64_t rows_per = (M + slabs - 1) / slabs;
    dim3 grid(nblk(N, 256), (unsigned)slabs);
    colsum_part_kernel<<<grid, 256, 0, st>>>(x, M, N, rows_per, (float*)ws);
    GAOT_LAUNCH_CHECK();
    colsum_reduce_kernel<<<nblk(N, 32), 256, 0, st>>>((const float*)ws, (int)slabs, N, out);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_swiglu_forward(const void* gu, int64_t M, int32_t F, void* a, void* stream) {
    GAOT_CHECK_ARG(F % 8 == 0 && F > 0, "swiglu: F must be a positive multiple of 8");
    if (M == 0) return GAOT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_TIME_KERNEL("swiglu_fwd", st, (double)M * F * 6.0);
    swiglu_fwd_kernel<<<nblk(M * (F / 8), 256), 256, 0, st>>>((const bf16*)gu, M, F, (bf16*)a);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_swiglu_backward(const void* da, const void* gu, int64_t M, int32_t F, void* dgu, void* stream) {
    GAOT_CHECK_ARG(F % 8 == 0 && F > 0, "swiglu: F must be a positive multiple of 8");
    if (M == 0) return GAOT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_TIME_KERNEL("swiglu_bwd", st, (double)M * F * 10.0);
    swiglu_bwd_kernel<<<nblk(M * (F / 8), 256), 256, 0, st>>>((const bf16*)da, (const bf16*)gu, M, F, (bf16*)dgu);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // extern "C"
