// node_mlp.cu -- fused two-layer node MLP  y = W2 gelu(W1 x + b1) + b2  on [N_points, 32] rows: the decoder's
// projection (reference src/model/layers/magno.py:640-644, :796-797: Linear(C -> 256) - GELU - Linear(256 -> C_out) on
// every physical point).  As torch ops this is 2 GEMMs + an elementwise GELU forward and 4 GEMMs + GELU' + 2 bias
// reductions backward, all of them streaming the [N, 256] hidden tensor through HBM (1 KB per point each way, 8 GB at
// 8 M points).  Here the hidden activations never leave the SM:
//   forward : two 128-point streams per CTA (one row per thread);  [x | 1] f16  x  [W1 | b1]  -> TMEM (256 columns)
//             -> packed-half GELU (2 gelu(z); the 1/2 sits in W2) -> f16 tile -> x [W2/2 | b2] (N = 16) -> y
//   backward: recompute gelu and gelu' from one tanh; dW2^T (+ db2) and dW1 (+ db1) accumulate in TMEM across all tiles
//             of the CTA; dA = dY W2 and dx = dZ1 W1 re-read the weight tiles as MN-major operands (tc05.cuh chunk-major
//             tiles are both); bf16 operands for everything that carries a gradient.
// Shape envelope: C_in = 32, hidden = 256, C_out <= 8 (the projection head); anything else stays on the torch path.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc05.cuh"

namespace gaot {

namespace nm {
constexpr int CIN = 32, HID = 256, KP1 = CIN + 16, KP2 = HID + 16, NOUT = 16, TE = 128;
constexpr int FG = 2, FT = 128 * FG;                 // forward: streams per CTA, threads
constexpr int W1_B = HID * KP1 * 2;                  // 24576: [256 x 48]
constexpr int W2_B = NOUT * KP2 * 2;                 // 8704:  [16 x 272]
constexpr int A0_B = TE * KP1 * 2;                   // 12288: [128 x 48]
constexpr int ACT_B = TE * KP2 * 2;                  // 69632: [128 x 272]
constexpr int F_W1 = 0, F_W2 = W1_B, F_GRP = 33792, F_GSTRIDE = A0_B + ACT_B;      // per stream: [A0 | ACT]
static_assert(W1_B + W2_B <= F_GRP && F_GRP % 1024 == 0, "weight tiles must fit in front of the per-stream regions");
constexpr int F_SMEM = F_GRP + FG * F_GSTRIDE;       // 197632
// backward
constexpr int BT = 256;
constexpr int B_W1 = 0, B_W2 = W1_B, B_A0 = 33792, B_DY = B_A0 + A0_B, B_G = B_DY + TE * NOUT * 2, B_GP = B_G + ACT_B;
constexpr int B_SMEM = B_GP + TE * HID * 2;          // 33792 + 12288 + 4096 + 69632 + 65536 = 185344
constexpr int TM_DW2 = 256, TM_DW1 = 256 + 48;       // TMEM columns: work 0..255, dW2^T (3 x 16), dW1 (2 x 48)
}

__device__ __forceinline__ uint32_t nm_h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 nm_u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ __half2 nm_tanh(__half2 u) {
    uint32_t r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(nm_h2u(u)));
    return nm_u2h(r);
}
// 2 gelu(x) (tanh form): x + x tanh(x (c0 + c1 x^2))
__device__ __forceinline__ __half2 nm_gelu2x(__half2 x) {
    const __half2 c1 = __float2half2_rn(0.0356774081f), c0 = __float2half2_rn(0.7978845608f);
    const __half2 u = __hmul2(x, __hfma2(__hmul2(x, x), c1, c0));
    return __hfma2(x, nm_tanh(u), x);
}
__device__ __forceinline__ void nm_gelu_and_grad(__half2 x, __half2& g, __half2& dg) {
    const __half2 c1 = __float2half2_rn(0.0356774081f), c0 = __float2half2_rn(0.7978845608f), hf = __float2half2_rn(0.5f);
    const __half2 c3 = __float2half2_rn(0.1070322243f), one = __float2half2_rn(1.0f), cap = __float2half2_rn(16384.f);
    const __half2 x2 = __hmin2(__hmul2(x, x), cap);
    const __half2 t = nm_tanh(__hmul2(x, __hfma2(x2, c1, c0)));
    const __half2 hx = __hmul2(x, hf);
    g = __hfma2(hx, t, hx);
    const __half2 sech2 = __hfma2(__hneg2(t), t, one);
    dg = __hfma2(__hmul2(hx, sech2), __hfma2(x2, c3, c0), __hfma2(t, hf, hf));
}
__host__ __device__ constexpr uint32_t nm_idesc_f16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void nm_group_bar(int g) { asm volatile("bar.sync %0, 128;\n" ::"r"(g + 1) : "memory"); }

// weight tiles: W1 -> [256 x 48] = [W1 | b1 | 0], W2 -> [16 x 272] = [s W2 | b2 | 0] (rows >= c_out zero); T = __half or bf16
template <typename T>
__device__ __forceinline__ void nm_stage_weights(uint8_t* sm_w1, uint8_t* sm_w2, const float* __restrict__ w1, const float* __restrict__ b1,
                                                 const float* __restrict__ w2, const float* __restrict__ b2, int c_out, float s2,
                                                 int tid, int nthreads) {
    for (int idx = tid; idx < nm::HID * nm::KP1; idx += nthreads) {
        const int n = idx / nm::KP1, k = idx - n * nm::KP1;
        const float v = k < nm::CIN ? w1[n * nm::CIN + k] : (k == nm::CIN ? b1[n] : 0.f);
        *reinterpret_cast<T*>(sm_w1 + tc::cm_off(nm::HID, n, k)) = T(v);
    }
    for (int idx = tid; idx < nm::NOUT * nm::KP2; idx += nthreads) {
        const int n = idx / nm::KP2, k = idx - n * nm::KP2;
        float v = 0.f;
        if (n < c_out) v = k < nm::HID ? s2 * w2[n * nm::HID + k] : (k == nm::HID && b2 ? b2[n] : 0.f);
        *reinterpret_cast<T*>(sm_w2 + tc::cm_off(nm::NOUT, n, k)) = T(v);
    }
}

// =====================================================================================================
// forward
// =====================================================================================================
__global__ void __launch_bounds__(nm::FT, 1)
node_mlp2_fwd_kernel(const float* __restrict__ x, int64_t n, int c_out, const float* __restrict__ w1, const float* __restrict__ b1,
                     const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ y, int ntiles) {
    using namespace nm;
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar[FG];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int g = tid >> 7, row = tid & 127, wg = warp & 3;
    uint8_t* A0 = sm + F_GRP + g * F_GSTRIDE;
    uint8_t* ACT = A0 + A0_B;

    nm_stage_weights<__half>(sm + F_W1, sm + F_W2, w1, b1, w2, b2, c_out, 0.5f, tid, FT);
    // constant ones / zero chunks: A0 chunk 4 (K 32..39) and 5, ACT chunks 32 and 33
    *reinterpret_cast<uint4*>(A0 + 4 * (128 * 16) + row * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(A0 + 5 * (128 * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ACT + 32 * (128 * 16) + row * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ACT + 33 * (128 * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { for (int i = 0; i < FG; ++i) tc::mbar_init(&mbar[i], 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_g = tmem_base_s + g * 256;
    const uint32_t tlane = tmem_g + ((uint32_t)(wg * 32) << 16);
    uint32_t ph = 0;
    const tc::Desc dA0 = tc::kmajor(tc::smem_u32(A0), 128), dAct = tc::kmajor(tc::smem_u32(ACT), 128);
    const tc::Desc dW1 = tc::kmajor(tc::smem_u32(sm + F_W1), HID), dW2 = tc::kmajor(tc::smem_u32(sm + F_W2), NOUT);
    constexpr uint32_t KS128 = tc::kstep_kmajor(128);

    const int tstep = gridDim.x * FG;
    int tile = blockIdx.x * FG + g;
    float4 xr[8];                                       // this thread's input row (prefetched one tile ahead)
    auto load_row = [&](int t) {
        const int64_t p = (int64_t)t * TE + row;
        if (t < ntiles && p < n) {
            const float4* src = reinterpret_cast<const float4*>(x + p * CIN);
#pragma unroll
            for (int j = 0; j < 8; ++j) xr[j] = __ldg(src + j);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) xr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_row(tile);
    for (; tile < ntiles; tile += tstep) {
        // ---- [x | 1] -> f16 A operand ----
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            const float4 a = xr[2 * c8], b = xr[2 * c8 + 1];
            *reinterpret_cast<uint4*>(A0 + c8 * (128 * 16) + row * 16) =
                make_uint4(nm_h2u(__floats2half2_rn(a.x, a.y)), nm_h2u(__floats2half2_rn(a.z, a.w)),
                           nm_h2u(__floats2half2_rn(b.x, b.y)), nm_h2u(__floats2half2_rn(b.z, b.w)));
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        nm_group_bar(g);
        if (wg == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                constexpr uint32_t idesc = nm_idesc_f16(128, HID, 0, 0);
#pragma unroll
                for (int s = 0; s < KP1 / 16; ++s)
                    tc::mma_bf16(tmem_g, dA0.adv(s * KS128).u64(), dW1.adv(s * tc::kstep_kmajor(HID)).u64(), idesc, s > 0);
                tc::mma_commit(&mbar[g]);
            }
            __syncwarp();
        }
        load_row(tile + tstep);                          // next tile's row while the tensor core and the GELU work
        tc::mbar_wait(&mbar[g], ph); ph ^= 1;
        tc::fence_after_sync();
        // ---- hidden layer: 256 columns per thread, 64 at a time ----
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            uint32_t v0[32], v1[32];
            tc::tmem_ld32_nowait(tlane + q * 64, v0);
            tc::tmem_ld32_nowait(tlane + q * 64 + 32, v1);
            tc::tmem_wait_ld();
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = (c8 & 3) * 8 + 2 * j;
                    const float lo_ = __uint_as_float(c8 < 4 ? v0[c] : v1[c]);
                    const float hi_ = __uint_as_float(c8 < 4 ? v0[c + 1] : v1[c + 1]);
                    o[j] = nm_h2u(nm_gelu2x(__floats2half2_rn(lo_, hi_)));
                }
                *reinterpret_cast<uint4*>(ACT + (q * 8 + c8) * (128 * 16) + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        nm_group_bar(g);
        if (wg == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                constexpr uint32_t idesc = nm_idesc_f16(128, NOUT, 0, 0);
#pragma unroll
                for (int s = 0; s < KP2 / 16; ++s)
                    tc::mma_bf16(tmem_g, dAct.adv(s * KS128).u64(), dW2.adv(s * tc::kstep_kmajor(NOUT)).u64(), idesc, s > 0);
                tc::mma_commit(&mbar[g]);
            }
            __syncwarp();
        }
        tc::mbar_wait(&mbar[g], ph); ph ^= 1;
        tc::fence_after_sync();
        {
            float v[8];
            tc::tmem_ld8(tlane, v);
            const int64_t p = (int64_t)tile * TE + row;
            if (p < n) {
                if (c_out == 4) *reinterpret_cast<float4*>(y + p * 4) = make_float4(v[0], v[1], v[2], v[3]);
                else for (int c = 0; c < c_out; ++c) y[p * c_out + c] = v[c];
            }
        }
        tc::fence_before_sync();
        nm_group_bar(g);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

// =====================================================================================================
// backward
// =====================================================================================================
__global__ void __launch_bounds__(nm::BT, 1)
node_mlp2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ d_y, int64_t n, int c_out, const float* __restrict__ w1,
                     const float* __restrict__ b1, const float* __restrict__ w2, float* __restrict__ d_x, float* __restrict__ partial,
                     int n_params, int ntiles) {
    using namespace nm;
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    uint8_t* A0 = sm + B_A0;
    uint8_t* DY = sm + B_DY;
    uint8_t* G = sm + B_G;
    uint8_t* GP = sm + B_GP;

    nm_stage_weights<__nv_bfloat16>(sm + B_W1, sm + B_W2, w1, b1, w2, nullptr, c_out, 1.0f, tid, BT);
    if (tid < TE) {
        *reinterpret_cast<uint4*>(A0 + 5 * (128 * 16) + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(G + 32 * (128 * 16) + tid * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(G + 33 * (128 * 16) + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(DY + 1 * (128 * 16) + tid * 16) = make_uint4(0u, 0u, 0u, 0u);      // columns 8..15 of dY
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph = 0;
    constexpr uint32_t KS128 = tc::kstep_kmajor(128);
    const uint32_t sA0 = tc::smem_u32(A0), sDY = tc::smem_u32(DY), sG = tc::smem_u32(G), sGP = tc::smem_u32(GP);
    const uint32_t sW1 = tc::smem_u32(sm + B_W1), sW2 = tc::smem_u32(sm + B_W2);
    bool first_tile = true;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p = (int64_t)tile * TE + row;
        const bool valid = p < n;
        // ---- [x | 1] (bf16) and dY (bf16, columns >= c_out zero): half 0 stages x, half 1 stages dY ----
        if (half == 0) {
            const float4* src = reinterpret_cast<const float4*>(x + p * CIN);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                const float4 a = valid ? __ldg(src + 2 * c8) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 b = valid ? __ldg(src + 2 * c8 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<uint4*>(A0 + c8 * (128 * 16) + row * 16) =
                    make_uint4(tc::pack_bf16(a.x, a.y), tc::pack_bf16(a.z, a.w), tc::pack_bf16(b.x, b.y), tc::pack_bf16(b.z, b.w));
            }
            // the ones column: bias input of the recompute AND the column that returns db1 (0 for padding rows)
            *reinterpret_cast<uint4*>(A0 + 4 * (128 * 16) + row * 16) = make_uint4(valid ? 0x00003F80u : 0u, 0u, 0u, 0u);
        } else {
            float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) for (int c = 0; c < c_out; ++c) d[c] = d_y[p * c_out + c];
            *reinterpret_cast<uint4*>(DY + row * 16) =
                make_uint4(tc::pack_bf16(d[0], d[1]), tc::pack_bf16(d[2], d[3]), tc::pack_bf16(d[4], d[5]), tc::pack_bf16(d[6], d[7]));
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // ---- recompute Z = [x | 1] [W1 | b1]^T ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const tc::Desc dA = tc::kmajor(sA0, 128), dW = tc::kmajor(sW1, HID);
                constexpr uint32_t idesc = tc::make_idesc_bf16(128, HID, 0, 0);
#pragma unroll
                for (int s = 0; s < KP1 / 16; ++s)
                    tc::mma_bf16(tmem, dA.adv(s * KS128).u64(), dW.adv(s * tc::kstep_kmajor(HID)).u64(), idesc, s > 0);
                tc::mma_commit(&mbar);
            }
            __syncwarp();
        }
        tc::mbar_wait(&mbar, ph); ph ^= 1;
        tc::fence_after_sync();
        // gelu -> bf16 activation tile G, gelu' -> f16 tile GP; this thread: columns [128 half, 128 half + 128)
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float v[32];
            tc::tmem_ld32(tlane + half * 128 + q * 32, v);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint32_t og[4], od[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __half2 gg, dg;
                    nm_gelu_and_grad(__floats2half2_rn(v[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1]), gg, dg);
                    const float2 gf = __half22float2(gg);
                    og[j] = tc::pack_bf16(gf.x, gf.y);
                    od[j] = nm_h2u(dg);
                }
                const int ch = half * 16 + q * 4 + c8;
                *reinterpret_cast<uint4*>(G + ch * (128 * 16) + row * 16) = make_uint4(og[0], og[1], og[2], og[3]);
                *reinterpret_cast<uint4*>(GP + ch * (128 * 16) + row * 16) = make_uint4(od[0], od[1], od[2], od[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // ---- dW2^T (+ db2) += [G | 1]^T dY   and   dA = dY W2 ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const tc::Desc dB = tc::mnmajor(sDY, 128);                                   // [128 points x 16]
                constexpr uint32_t idw = tc::make_idesc_bf16(128, NOUT, 1, 1);
#pragma unroll
                for (int m = 0; m < 3; ++m) {                                                // hidden rows 0..127, 128..255, ones chunk
                    const tc::Desc dAm = tc::mnmajor(sG + m * 16 * (128 * 16), 128);
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        tc::mma_bf16(tmem + TM_DW2 + m * 16, dAm.adv(s * tc::KSTEP_MN).u64(), dB.adv(s * tc::KSTEP_MN).u64(), idw,
                                     !(first_tile && s == 0));
                }
                const tc::Desc dAk = tc::kmajor(sDY, 128), dWm = tc::mnmajor(sW2, NOUT);      // B[h, c] = W2[c, h]
                constexpr uint32_t ida = tc::make_idesc_bf16(128, HID, 0, 1);
                tc::mma_bf16(tmem, dAk.u64(), dWm.u64(), ida, false);
                tc::mma_commit(&mbar);
            }
            __syncwarp();
        }
        tc::mbar_wait(&mbar, ph); ph ^= 1;
        tc::fence_after_sync();
        // dZ1 = dA * gelu'(z) -> bf16, in place over GP
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float v[32];
            tc::tmem_ld32(tlane + half * 128 + q * 32, v);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint8_t* pp = GP + (half * 16 + q * 4 + c8) * (128 * 16) + row * 16;
                const uint4 dd = *reinterpret_cast<const uint4*>(pp);
                const uint32_t dw[4] = {dd.x, dd.y, dd.z, dd.w};
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 gpf = __half22float2(nm_u2h(dw[j]));
                    o[j] = tc::pack_bf16(v[c8 * 8 + 2 * j] * gpf.x, v[c8 * 8 + 2 * j + 1] * gpf.y);
                }
                *reinterpret_cast<uint4*>(pp) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // ---- dW1 (+ db1) += dZ1^T [x | 1]   and   dx = dZ1 W1 ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const tc::Desc dB = tc::mnmajor(sA0, 128);                                   // [128 points x 48]
                constexpr uint32_t idw = tc::make_idesc_bf16(128, KP1, 1, 1);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const tc::Desc dAm = tc::mnmajor(sGP + m * 16 * (128 * 16), 128);
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        tc::mma_bf16(tmem + TM_DW1 + m * KP1, dAm.adv(s * tc::KSTEP_MN).u64(), dB.adv(s * tc::KSTEP_MN).u64(), idw,
                                     !(first_tile && s == 0));
                }
                const tc::Desc dAk = tc::kmajor(sGP, 128), dWm = tc::mnmajor(sW1, HID);       // B[c_in, h] = W1[h, c_in]
                constexpr uint32_t idx_ = tc::make_idesc_bf16(128, CIN, 0, 1);
#pragma unroll
                for (int s = 0; s < HID / 16; ++s)
                    tc::mma_bf16(tmem, dAk.adv(s * KS128).u64(), dWm.adv(s * tc::KSTEP_MN).u64(), idx_, s > 0);
                tc::mma_commit(&mbar);
            }
            __syncwarp();
        }
        tc::mbar_wait(&mbar, ph); ph ^= 1;
        tc::fence_after_sync();
        if (d_x) {
            float v[16];
            tc::tmem_ld16(tlane + half * 16, v);
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(d_x + p * CIN + half * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        first_tile = false;
        tc::fence_before_sync();
        __syncthreads();
    }

    // ---- flush the per-CTA accumulators: partial = [dW1 (256 x 32) | db1 (256) | dW2 (c_out x 256) | db2 (c_out)] ----
    tc::fence_after_sync();
    float* mine = partial + (size_t)blockIdx.x * n_params;
    if (first_tile) {
        for (int i = tid; i < n_params; i += BT) mine[i] = 0.f;
    } else if (warp < 4) {
        const int r = warp * 32 + lane;                         // accumulator row within a 128-row block
        const int o_b1 = HID * CIN, o_w2 = o_b1 + HID, o_b2 = o_w2 + c_out * HID;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            float t0[32], t1[16];
            tc::tmem_ld32(tlane + TM_DW1 + m * KP1, t0);
            tc::tmem_ld16(tlane + TM_DW1 + m * KP1 + 32, t1);
            const int h = m * 128 + r;
            for (int k = 0; k < 32; ++k) mine[h * CIN + k] = t0[k];
            mine[o_b1 + h] = t1[0];
            float t2[16];
            tc::tmem_ld16(tlane + TM_DW2 + m * 16, t2);
            for (int c = 0; c < c_out; ++c) mine[o_w2 + c * HID + h] = t2[c];
        }
        float t3[16];
        tc::tmem_ld16(tlane + TM_DW2 + 2 * 16, t3);
        if (r == 0) for (int c = 0; c < c_out; ++c) mine[o_b2 + c] = t3[c];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int gno_bwd_reduce(const float* partial, int nparts, int n_params, float* d_params, cudaStream_t st);   // gno_bwd.cu

}  // namespace gaot

using namespace gaot;

extern "C" {

int gaot_node_mlp2_supported(int32_t c_in, int32_t hidden, int32_t c_out) {
    return c_in == nm::CIN && hidden == nm::HID && c_out >= 1 && c_out <= 8;
}

size_t gaot_node_mlp2_workspace_bytes(int32_t c_in, int32_t hidden, int32_t c_out) {
    const size_t np = (size_t)hidden * c_in + hidden + (size_t)c_out * hidden + c_out;
    return align_up((size_t)kNumSMs * np * sizeof(float)) + align_up(np * sizeof(float)) + 512;
}

int gaot_node_mlp2_forward(const float* x, int64_t n, int32_t c_in, int32_t hidden, int32_t c_out, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* y, void* stream) {
    GAOT_CHECK_ARG(x && w1 && b1 && w2 && b2 && y && n >= 0, "node_mlp2_forward: bad arguments");
    if (!gaot_node_mlp2_supported(c_in, hidden, c_out)) { set_error("node_mlp2: shape %d -> %d -> %d outside the fused kernel (32 -> 256 -> <= 8)", c_in, hidden, c_out); return GAOT_ERR_UNSUPPORTED; }
    if (n == 0) return GAOT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int ntiles = (int)((n + nm::TE - 1) / nm::TE);
    const int want = (ntiles + nm::FG - 1) / nm::FG;
    const int grid = want < kNumSMs ? want : kNumSMs;
    GAOT_CUDA(cudaFuncSetAttribute(node_mlp2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nm::F_SMEM));
    {
        GAOT_TIME_KERNEL("node_mlp_fwd", st, (double)n * 4.0 * (c_in + c_out));
        node_mlp2_fwd_kernel<<<grid, nm::FT, nm::F_SMEM, st>>>(x, n, c_out, w1, b1, w2, b2, y, ntiles);
    }
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

// d_params = [dW1 (hidden x c_in) | db1 (hidden) | dW2 (c_out x hidden) | db2 (c_out)], d_x may be null
int gaot_node_mlp2_backward(const float* x, const float* d_y, int64_t n, int32_t c_in, int32_t hidden, int32_t c_out, const float* w1,
                            const float* b1, const float* w2, void* ws, size_t ws_bytes, float* d_x, float* d_params, void* stream) {
    GAOT_CHECK_ARG(x && d_y && w1 && b1 && w2 && d_params && n >= 0, "node_mlp2_backward: bad arguments");
    if (!gaot_node_mlp2_supported(c_in, hidden, c_out)) { set_error("node_mlp2: shape outside the fused kernel"); return GAOT_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    const int np = hidden * c_in + hidden + c_out * hidden + c_out;
    if (n == 0) { GAOT_CUDA(cudaMemsetAsync(d_params, 0, (size_t)np * sizeof(float), st)); return GAOT_OK; }
    const int ntiles = (int)((n + nm::TE - 1) / nm::TE);
    const int grid = ntiles < kNumSMs ? ntiles : kNumSMs;
    Arena ar(ws, ws_bytes);
    float* partial = ar.take<float>((size_t)grid * np);
    if (!ar.ok()) { set_error("node_mlp2_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    GAOT_CUDA(cudaFuncSetAttribute(node_mlp2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nm::B_SMEM));
    {
        GAOT_TIME_KERNEL("node_mlp_bwd", st, (double)n * 4.0 * (2 * c_in + c_out));
        node_mlp2_bwd_kernel<<<grid, nm::BT, nm::B_SMEM, st>>>(x, d_y, n, c_out, w1, b1, w2, d_x, partial, np, ntiles);
    }
    GAOT_LAUNCH_CHECK();
    return gno_bwd_reduce(partial, grid, np, d_params, st);
}

}  // extern "C"
