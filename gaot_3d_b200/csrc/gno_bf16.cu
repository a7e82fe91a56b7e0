// gno_bf16.cu -- fused GNO forward with the per-edge kernel MLP on tcgen05 tensor cores
// (BF16 operands, FP32 accumulation in TMEM).  Same contract as gno_fwd.cu (FP32 path).
//
// Tile = 128 consecutive CSR edges = the M dimension of one tcgen05.mma; 256 threads (2 per edge).
//   * coordinates are split into bf16 hi + lo parts (K = 12 -> 16) so that the first layer sees
//     16-bit-mantissa positions: plain bf16 coordinates (8 bits) would smear the geometry at the
//     scale of the GNO radius;
//   * weights live in shared memory as bf16 chunk-major K-major B operands for the whole kernel;
//   * layer l: D[128, N_l] = A_l[128, K_l] W_l^T (K_l/16 MMAs) -> TMEM -> each thread pulls its
//     32 columns (tcgen05.ld), adds the bias, applies GELU, rounds to bf16 and writes the next
//     A operand in place (chunk-major, conflict-free 16-byte stores);
//   * source feature rows f_y[src] are staged by the TMA bulk-copy engine (cp.async.bulk, one
//     128-byte row per request, mbarrier completion) while the MLP runs;
//   * the last layer output is multiplied by f_y[src] in FP32 and reduced with the same
//     deterministic CSR-ordered segmented mean as the FP32 kernel.
// GELU: in this mode the activations are rounded to bf16 (rel. resolution 4e-3) anyway, so the
// erf form is evaluated through its tanh representation with tanh.approx (|dev| < 1e-3).
#include "gno_common.cuh"
#include "tc05.cuh"

namespace gaot {

constexpr int TTE = 128;
constexpr int TTHREADS = 256;

struct TcLayout {
    int w_off[GNO_MAX_LAYERS];     // byte offsets of the bf16 weight tiles
    int b_off[GNO_MAX_LAYERS];     // float offsets (from bias base) of the biases
    int kpad[GNO_MAX_LAYERS];      // padded K of each layer (layer 0: 16)
    int bias_base, a0, act, fsm, ints, total_bytes;
};

static TcLayout tc_layout(const GnoArgs& a) {
    TcLayout L;
    int off = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        L.kpad[l] = (l == 0) ? 16 : a.dims[l];
        L.w_off[l] = off; off += a.dims[l + 1] * L.kpad[l] * 2;
    }
    off = (off + 1023) / 1024 * 1024;
    L.bias_base = off;
    int bo = 0;
    for (int l = 0; l < a.n_layers; ++l) { L.b_off[l] = bo; bo += a.dims[l + 1]; }
    off += (bo * 4 + 127) / 128 * 128;
    off = (off + 1023) / 1024 * 1024;
    L.a0 = off; off += TTE * 16 * 2;                       // layer-0 operand [128 x 16] bf16
    L.act = off; off += 36 * 1024;                         // two [128 x 64] bf16 operand tiles; re-used as the fp32 value tile [128 x (Cout+4)]
    L.fsm = off; off += (a.f_y ? TTE * a.c_f * 4 : 0);
    L.ints = off; off += (3 * TTE + 16) * 4;
    L.total_bytes = off;
    return L;
}

__device__ __forceinline__ float gelu_tanh_fast(float x) {
    // 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) ))
    const float x2 = x * x;
    const float u = x * fmaf(x2, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(tc::smem_u32(mbar)), "r"(bytes) : "memory");
}

template <int NL>
__global__ void __launch_bounds__(TTHREADS, 2)
gno_fwd_tc_kernel(const GnoArgs a, const TcLayout L, float* __restrict__ out, float* __restrict__ head_partial) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar_mma, mbar_f;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    float* bias = reinterpret_cast<float*>(sm + L.bias_base);
    uint8_t* A0 = sm + L.a0;
    uint8_t* ACT = sm + L.act;
    float* fsm = reinterpret_cast<float*>(sm + L.fsm);
    int* s_src = reinterpret_cast<int*>(sm + L.ints);
    int* s_qry = s_src + TTE;
    int* seg_first = s_qry + TTE;
    int* s_misc = seg_first + TTE + 1;

    // ---- weights -> bf16 chunk-major K-major B tiles, biases -> fp32 ----
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1], KP = L.kpad[l];
        const float* W = a.params + a.w_off[l];
        uint8_t* Ws = sm + L.w_off[l];
        for (int idx = tid; idx < N * KP; idx += TTHREADS) {
            const int n = idx / KP, k = idx - n * KP;
            float v;
            if (l == 0) v = k < 12 ? W[n * K + (k % 6)] : 0.f;        // [hi(6) | lo(6) | 0 0 0 0] all see the same weights
            else v = W[n * K + k];
            *reinterpret_cast<__nv_bfloat16*>(Ws + tc::cm_off(N, n, k)) = __float2bfloat16(v);
        }
        for (int j = tid; j < N; j += TTHREADS) bias[L.b_off[l] + j] = a.params[a.b_off[l] + j];
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { tc::mbar_init(&mbar_mma, 1); tc::mbar_init(&mbar_f, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_mma = 0, ph_f = 0;

    const int Cout = a.dims[NL];
    const int CP = Cout + 4;
    const bool use_f_mul = (a.transform == 0);
    const tc::Desc dA0 = tc::kmajor(tc::smem_u32(A0), 128);
    const tc::Desc dAct0 = tc::kmajor(tc::smem_u32(ACT), 128), dAct1 = tc::kmajor(tc::smem_u32(ACT + TTE * 64 * 2), 128);
    constexpr uint32_t KS128 = tc::kstep_kmajor(128);

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int e0 = tile * TTE;
        const int ne = min(TTE, a.E - e0);
        if (tid < TTE) {
            const bool valid = tid < ne;
            s_src[tid] = valid ? a.csr_src[e0 + tid] : 0;
            s_qry[tid] = valid ? a.csr_qry[e0 + tid] : -1;
        }
        __syncthreads();
        // ---- TMA bulk copies of the feature rows (half 1 threads), layer-0 operand (half 0 threads) ----
        if (half == 1) {
            if (a.f_y) {
                if (row == 0) mbar_expect_tx(&mbar_f, (uint32_t)ne * a.c_f * 4);
                __syncwarp();
                if (row < ne) bulk_copy_g2s(fsm + row * a.c_f, a.f_y + (size_t)s_src[row] * a.c_f, a.c_f * 4, &mbar_f);
            }
            const bool valid = row < ne;
            const bool head = valid && (row == 0 || s_qry[row - 1] != s_qry[row]);
            const unsigned bm = __ballot_sync(0xffffffffu, head);
            if (lane == 0) s_misc[warp & 3] = (int)bm;
        } else {
            const bool valid = row < ne;
            float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) {
                const float* py = a.y_pos + (size_t)s_src[row] * 3;
                const float* px = a.x_pos + (size_t)s_qry[row] * 3;
                c6[0] = py[0]; c6[1] = py[1]; c6[2] = py[2]; c6[3] = px[0]; c6[4] = px[1]; c6[5] = px[2];
            }
            float hi[6], lo[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                hi[j] = __bfloat162float(__float2bfloat16(c6[j]));
                lo[j] = c6[j] - hi[j];
            }
            uint4 c0, c1;
            c0.x = tc::pack_bf16(hi[0], hi[1]); c0.y = tc::pack_bf16(hi[2], hi[3]);
            c0.z = tc::pack_bf16(hi[4], hi[5]); c0.w = tc::pack_bf16(lo[0], lo[1]);
            c1.x = tc::pack_bf16(lo[2], lo[3]); c1.y = tc::pack_bf16(lo[4], lo[5]); c1.z = 0u; c1.w = 0u;
            *reinterpret_cast<uint4*>(A0 + row * 16) = c0;
            *reinterpret_cast<uint4*>(A0 + 128 * 16 + row * 16) = c1;
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid >= TTE) {                                  // segment table (threads of half 1)
            const int w4 = warp & 3;
            int before = 0;
            for (int w = 0; w < w4; ++w) before += __popc((unsigned)s_misc[w]);
            const unsigned mine = (unsigned)s_misc[w4];
            if ((mine >> lane) & 1u) seg_first[before + __popc(mine & ((1u << lane) - 1u))] = row;
            if (row == 0) {
                const int nseg = __popc((unsigned)s_misc[0]) + __popc((unsigned)s_misc[1]) +
                                 __popc((unsigned)s_misc[2]) + __popc((unsigned)s_misc[3]);
                s_misc[4] = nseg;
                seg_first[nseg] = ne;
            }
        }

        // ---- MLP: one tcgen05 GEMM per layer, activations stay on chip ----
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const int N = a.dims[l + 1], KP = L.kpad[l];
            const uint32_t tm_d = tmem + (l & 1) * 64;
            if (warp == 0) {
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    const tc::Desc dA = (l == 0) ? dA0 : (((l - 1) & 1) ? dAct1 : dAct0);
                    const tc::Desc dW = tc::kmajor(tc::smem_u32(sm + L.w_off[l]), N);
                    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
                    const uint32_t ksw = tc::kstep_kmajor(N);
                    for (int s = 0; s < KP / 16; ++s)
                        tc::mma_bf16(tm_d, dA.adv(s * KS128).u64(), dW.adv(s * ksw).u64(), idesc, s > 0);
                    tc::mma_commit(&mbar_mma);
                }
                __syncwarp();
            }
            tc::mbar_wait(&mbar_mma, ph_mma); ph_mma ^= 1;
            tc::fence_after_sync();
            const uint32_t tl = tlane + (l & 1) * 64;
            if (l < NL - 1) {
                // this thread: edge `row`, columns [32*half, 32*half + 32) of the 64-wide hidden layer
                float v[32];
                tc::tmem_ld32(tl + half * 32, v);
                const float* bs = bias + L.b_off[l] + half * 32;
                uint8_t* dst = ACT + (l & 1) * (TTE * 64 * 2);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float g[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) g[c] = gelu_tanh_fast(v[c8 * 8 + c] + bs[c8 * 8 + c]);
                    uint4 o;
                    o.x = tc::pack_bf16(g[0], g[1]); o.y = tc::pack_bf16(g[2], g[3]);
                    o.z = tc::pack_bf16(g[4], g[5]); o.w = tc::pack_bf16(g[6], g[7]);
                    *reinterpret_cast<uint4*>(dst + (half * 4 + c8) * (128 * 16) + row * 16) = o;
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
            } else {
                // last layer: bias, (* f_y[src]) in fp32, edge-major value tile for the reduction
                if (a.f_y) { tc::mbar_wait(&mbar_f, ph_f); ph_f ^= 1; }
                const int nc = Cout / 2;                      // columns per thread: 16 (Cout 32) or 32 (Cout 64)
                float v[32];
                if (nc == 16) { float t[16]; tc::tmem_ld16(tl + half * 16, t);
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] = t[c]; }
                else tc::tmem_ld32(tl + half * 32, v);
                tc::fence_before_sync();
                __syncthreads();                              // every MMA operand read is done: ACT can hold the value tile
                float* val = reinterpret_cast<float*>(ACT);
                const float* bs = bias + L.b_off[l] + half * nc;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    if (c < nc) {
                        float4 o = make_float4(v[c] + bs[c], v[c + 1] + bs[c + 1], v[c + 2] + bs[c + 2], v[c + 3] + bs[c + 3]);
                        if (use_f_mul) {
                            const float4 f = *reinterpret_cast<const float4*>(fsm + row * a.c_f + half * nc + c);
                            o.x *= f.x; o.y *= f.y; o.z *= f.z; o.w *= f.w;
                        }
                        *reinterpret_cast<float4*>(val + row * CP + half * nc + c) = o;
                    }
                }
                __syncthreads();
                const int nseg = s_misc[4];
                for (int s = warp; s < nseg; s += TTHREADS / 32) {
                    const int first = seg_first[s], lastE = seg_first[s + 1];
                    const int q = s_qry[first];
                    const int rb = a.rowptr[q], re = a.rowptr[q + 1];
                    const bool starts_here = rb >= e0;
                    const bool ends_here = re <= e0 + ne;
                    for (int c = lane; c < Cout; c += 32) {
                        float sum = 0.f;
                        for (int e = first; e < lastE; ++e) sum += val[e * CP + c];
                        if (starts_here) {
                            if (ends_here && a.reduce == 0) sum = sum / (float)(re - rb);
                            out[(size_t)q * Cout + c] = sum;
                        } else {
                            head_partial[(size_t)tile * Cout + c] = sum;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

int gno_fwd_fixup(const GnoArgs& a, float* out, const float* head_partial, cudaStream_t st);   // gno_fwd.cu

// MLP shapes the tensor-core forward is built for; anything else runs the (more precise) FP32 CUDA-core kernel
bool gno_forward_bf16_supported(const GnoArgs& a) {
    if (a.transform != 0 && a.transform != 3) return false;
    if (a.n_layers < 2 || a.n_layers > 5) return false;
    for (int l = 1; l < a.n_layers; ++l) if (a.dims[l] != 64) return false;
    const int Cout = a.dims[a.n_layers];
    if (Cout != 32 && Cout != 64) return false;
    if (a.f_y && ((a.c_f * 4) % 16)) return false;
    return true;
}

int gno_forward_bf16(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st) {
    const int Cout = a.dims[a.n_layers];
    GAOT_CUDA(cudaMemsetAsync(out, 0, (size_t)a.nq * Cout * sizeof(float), st));
    if (a.E == 0) return GAOT_OK;
    Arena ar(ws, ws_bytes);
    float* head_partial = ar.take<float>((size_t)a.ntiles * Cout);
    if (!ar.ok()) { set_error("gno_forward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const TcLayout L = tc_layout(a);
    const size_t smem = (size_t)L.total_bytes;
    if (smem > 227 * 1024) { set_error("gno bf16: shared memory %zu B too large", smem); return GAOT_ERR_UNSUPPORTED; }
    int grid = a.ntiles < 2 * kNumSMs ? a.ntiles : 2 * kNumSMs;
    {
        GAOT_TIME_KERNEL("gno_fwd", st, (double)a.E * (16.0 + 12.0 + 4.0 * a.c_f) + (double)a.nq * (12.0 + 4.0 * Cout));
#define GAOT_TC_CASE(NL)                                                                                              \
    case NL:                                                                                                          \
        GAOT_CUDA(cudaFuncSetAttribute(gno_fwd_tc_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gno_fwd_tc_kernel<NL><<<grid, TTHREADS, smem, st>>>(a, L, out, head_partial);                                  \
        break;
        switch (a.n_layers) { GAOT_TC_CASE(2) GAOT_TC_CASE(3) GAOT_TC_CASE(4) GAOT_TC_CASE(5) default: break; }
#undef GAOT_TC_CASE
    }
    GAOT_LAUNCH_CHECK();
    return gno_fwd_fixup(a, out, head_partial, st);
}


// =====================================================================================================
// Backward on tensor cores.  Per 128-edge tile: forward recompute (z_l kept as bf16 tiles for gelu'),
// output-side gradients in fp32, then per layer TWO GEMM families queued together:
//   dW_l[n, k] += sum_e dZ_l[e, n] * [A_l | 1][e, k]   (M = channels padded to 128, K = 128 edges; the ones
//                                                       column of the activation tile returns db_l for free;
//                                                       accumulates in TMEM across ALL tiles of the CTA)
//   dA_l[e, k]  = sum_n dZ_l[e, n] * W_l[n, k]          (M = 128 edges; W_l tile re-read as an MN-major operand)
// followed by dZ_{l-1} = dA_l * gelu'(z_l) in registers -> bf16 -> next operand tile.
// TMEM: 2 x 64 working columns + one [channels x (K_l + 16)] fp32 accumulator per layer (<= 512 columns).
// =====================================================================================================
struct TcBwdLayout {
    int w_off[GNO_MAX_LAYERS], b_off[GNO_MAX_LAYERS], kpad[GNO_MAX_LAYERS], h_off[GNO_MAX_LAYERS], z_off[GNO_MAX_LAYERS];
    int tm_dw[GNO_MAX_LAYERS];
    int bias_base, a0, dza, dzb, fsm, gsm, ints, total_bytes;
};

static TcBwdLayout tc_bwd_layout(const GnoArgs& a) {
    TcBwdLayout L;
    int off = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        L.kpad[l] = (l == 0) ? 16 : a.dims[l];
        L.w_off[l] = off; off += a.dims[l + 1] * L.kpad[l] * 2;
    }
    off = (off + 1023) / 1024 * 1024;
    L.bias_base = off;
    int bo = 0;
    for (int l = 0; l < a.n_layers; ++l) { L.b_off[l] = bo; bo += a.dims[l + 1]; }
    off += (bo * 4 + 127) / 128 * 128;
    off = (off + 1023) / 1024 * 1024;
    L.a0 = off; off += TTE * 16 * 2;
    for (int l = 1; l < a.n_layers; ++l) { L.h_off[l] = off; off += TTE * 80 * 2; }      // [128 x (64 + ones chunk + zero chunk)]
    for (int l = 1; l < a.n_layers; ++l) { L.z_off[l] = off; off += TTE * 64 * 2; }
    L.h_off[0] = L.a0; L.z_off[0] = 0;
    L.dza = off; off += TTE * 64 * 2;
    L.dzb = off; off += TTE * 64 * 2;
    L.fsm = off; off += TTE * 64 * 4 / 2;                     // >= 16 KB: also the tail that an M=128 read of dzb may touch
    L.gsm = off; off += TTE * 64 * 4 / 2;
    L.ints = off; off += (3 * TTE + 16) * 4;
    L.total_bytes = off;
    int col = 128;
    for (int l = 0; l < a.n_layers; ++l) { L.tm_dw[l] = col; col += (l == 0) ? 32 : 80; }
    return L;
}

__device__ __forceinline__ float gelu_tanh_grad_fast(float x) {
    const float x2 = x * x;
    const float u = x * fmaf(x2, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float du = fmaf(x2, 0.1070322243f, 0.7978845608f);          // d u / d x
    const float sech2 = fmaf(-t, t, 1.0f);
    return fmaf(0.5f * x * sech2, du, 0.5f + 0.5f * t);
}

template <int NL>
__global__ void __launch_bounds__(TTHREADS, 1)
gno_bwd_tc_kernel(const GnoArgs a, const TcBwdLayout L, const float* __restrict__ d_out, float* __restrict__ d_f,
                  float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar_mma, mbar_f;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    float* bias = reinterpret_cast<float*>(sm + L.bias_base);
    uint8_t* A0 = sm + L.a0;
    float* fsm = reinterpret_cast<float*>(sm + L.fsm);
    float* gsm = reinterpret_cast<float*>(sm + L.gsm);
    int* s_src = reinterpret_cast<int*>(sm + L.ints);
    int* s_qry = s_src + TTE;
    float* s_inv = reinterpret_cast<float*>(s_qry + TTE);

#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1], KP = L.kpad[l];
        const float* W = a.params + a.w_off[l];
        uint8_t* Ws = sm + L.w_off[l];
        for (int idx = tid; idx < N * KP; idx += TTHREADS) {
            const int n = idx / KP, k = idx - n * KP;
            float v;
            if (l == 0) v = k < 12 ? W[n * K + (k % 6)] : 0.f;
            else v = W[n * K + k];
            *reinterpret_cast<__nv_bfloat16*>(Ws + tc::cm_off(N, n, k)) = __float2bfloat16(v);
        }
        for (int j = tid; j < N; j += TTHREADS) bias[L.b_off[l] + j] = a.params[a.b_off[l] + j];
        if (l >= 1 && tid < TTE) {                 // constant ones / zero chunks of the activation tiles
            *reinterpret_cast<uint4*>(sm + L.h_off[l] + 8 * (128 * 16) + tid * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sm + L.h_off[l] + 9 * (128 * 16) + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&mbar_mma, 1); tc::mbar_init(&mbar_f, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_mma = 0, ph_f = 0;

    const int Cout = a.dims[NL];
    const int nc = Cout / 2;                               // output columns per thread (16 or 32)
    const bool use_f_mul = (a.transform == 0);
    constexpr uint32_t KS128 = tc::kstep_kmajor(128);
    const uint32_t sDZ[2] = {tc::smem_u32(sm + L.dza), tc::smem_u32(sm + L.dzb)};
    bool first_tile = true;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int e0 = tile * TTE;
        const int ne = min(TTE, a.E - e0);
        if (tid < TTE) {
            const bool valid = tid < ne;
            const int q = valid ? a.csr_qry[e0 + tid] : 0;
            s_src[tid] = valid ? a.csr_src[e0 + tid] : 0;
            s_qry[tid] = q;
            s_inv[tid] = !valid ? 0.f : (a.reduce == 0 ? 1.0f / (float)(a.rowptr[q + 1] - a.rowptr[q]) : 1.0f);
        }
        __syncthreads();
        if (half == 1) {
            const uint32_t fbytes = a.f_y ? a.c_f * 4 : 0, gbytes = Cout * 4;
            if (row == 0) mbar_expect_tx(&mbar_f, (uint32_t)ne * (fbytes + gbytes));
            __syncwarp();
            if (row < ne) {
                if (a.f_y) bulk_copy_g2s(fsm + row * a.c_f, a.f_y + (size_t)s_src[row] * a.c_f, fbytes, &mbar_f);
                bulk_copy_g2s(gsm + row * Cout, d_out + (size_t)s_qry[row] * Cout, gbytes, &mbar_f);
            }
        } else {
            const bool valid = row < ne;
            float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) {
                const float* py = a.y_pos + (size_t)s_src[row] * 3;
                const float* px = a.x_pos + (size_t)s_qry[row] * 3;
                c6[0] = py[0]; c6[1] = py[1]; c6[2] = py[2]; c6[3] = px[0]; c6[4] = px[1]; c6[5] = px[2];
            }
            float hi[6], lo[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) { hi[j] = __bfloat162float(__float2bfloat16(c6[j])); lo[j] = c6[j] - hi[j]; }
            uint4 c0, c1;
            c0.x = tc::pack_bf16(hi[0], hi[1]); c0.y = tc::pack_bf16(hi[2], hi[3]);
            c0.z = tc::pack_bf16(hi[4], hi[5]); c0.w = tc::pack_bf16(lo[0], lo[1]);
            c1.x = tc::pack_bf16(lo[2], lo[3]); c1.y = tc::pack_bf16(lo[4], lo[5]);
            c1.z = valid ? 0x00003F80u : 0u;                   // K index 12 = 1.0: db_0 comes out of the dW_0 GEMM
            c1.w = 0u;
            *reinterpret_cast<uint4*>(A0 + row * 16) = c0;
            *reinterpret_cast<uint4*>(A0 + 128 * 16 + row * 16) = c1;
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // =============== forward recompute ===============
        float kv[32];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const int N = a.dims[l + 1], KP = L.kpad[l];
            if (warp == 0) {
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    const tc::Desc dA = tc::kmajor(tc::smem_u32(sm + L.h_off[l]), 128);
                    const tc::Desc dW = tc::kmajor(tc::smem_u32(sm + L.w_off[l]), N);
                    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
                    const uint32_t ksw = tc::kstep_kmajor(N);
                    for (int s = 0; s < KP / 16; ++s)
                        tc::mma_bf16(tmem + (l & 1) * 64, dA.adv(s * KS128).u64(), dW.adv(s * ksw).u64(), idesc, s > 0);
                    tc::mma_commit(&mbar_mma);
                }
                __syncwarp();
            }
            tc::mbar_wait(&mbar_mma, ph_mma); ph_mma ^= 1;
            tc::fence_after_sync();
            const uint32_t tl = tlane + (l & 1) * 64;
            if (l < NL - 1) {
                float v[32];
                tc::tmem_ld32(tl + half * 32, v);
                const float* bs = bias + L.b_off[l] + half * 32;
                uint8_t* zt = sm + L.z_off[l + 1];
                uint8_t* ht = sm + L.h_off[l + 1];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float z[8], g[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) { z[c] = v[c8 * 8 + c] + bs[c8 * 8 + c]; g[c] = gelu_tanh_fast(z[c]); }
                    uint4 o;
                    o.x = tc::pack_bf16(z[0], z[1]); o.y = tc::pack_bf16(z[2], z[3]);
                    o.z = tc::pack_bf16(z[4], z[5]); o.w = tc::pack_bf16(z[6], z[7]);
                    *reinterpret_cast<uint4*>(zt + (half * 4 + c8) * (128 * 16) + row * 16) = o;
                    o.x = tc::pack_bf16(g[0], g[1]); o.y = tc::pack_bf16(g[2], g[3]);
                    o.z = tc::pack_bf16(g[4], g[5]); o.w = tc::pack_bf16(g[6], g[7]);
                    *reinterpret_cast<uint4*>(ht + (half * 4 + c8) * (128 * 16) + row * 16) = o;
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
            } else {
                if (nc == 16) { float t[16]; tc::tmem_ld16(tl + half * 16, t);
#pragma unroll
                    for (int c = 0; c < 16; ++c) kv[c] = t[c]; }
                else tc::tmem_ld32(tl + half * 32, kv);
                const float* bs = bias + L.b_off[l] + half * nc;
#pragma unroll
                for (int c = 0; c < 32; ++c) if (c < nc) kv[c] += bs[c];
            }
        }

        // =============== output-side gradients (fp32) ===============
        tc::mbar_wait(&mbar_f, ph_f); ph_f ^= 1;
        {
            const bool valid = row < ne;
            const float inv = s_inv[row];
            uint8_t* dz = sm + L.dza;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                if (c8 * 8 < nc) {
                    float dk[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int cc = half * nc + c8 * 8 + c;
                        float g = valid ? gsm[row * Cout + cc] * inv : 0.f;
                        if (use_f_mul) {
                            const float f = valid ? fsm[row * a.c_f + cc] : 0.f;
                            kv[c8 * 8 + c] *= g;                       // d f_y contribution g * k
                            g *= f;
                        }
                        dk[c] = g;
                    }
                    if (use_f_mul && d_f && valid) {
                        float* dst = d_f + (size_t)s_src[row] * a.c_f + half * nc + c8 * 8;
                        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(kv[c8 * 8], kv[c8 * 8 + 1], kv[c8 * 8 + 2], kv[c8 * 8 + 3]));
                        atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(kv[c8 * 8 + 4], kv[c8 * 8 + 5], kv[c8 * 8 + 6], kv[c8 * 8 + 7]));
                    }
                    uint4 o;
                    o.x = tc::pack_bf16(dk[0], dk[1]); o.y = tc::pack_bf16(dk[2], dk[3]);
                    o.z = tc::pack_bf16(dk[4], dk[5]); o.w = tc::pack_bf16(dk[6], dk[7]);
                    *reinterpret_cast<uint4*>(dz + ((half * nc) / 8 + c8) * (128 * 16) + row * 16) = o;
                }
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // =============== backward through the layers ===============
        int cur = 0;
#pragma unroll
        for (int l = NL - 1; l >= 0; --l) {
            const int N = a.dims[l + 1];
            if (warp == 0) {
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    // dW_l (+ db_l): A = dZ^T (MN-major over the edge-major dZ tile), B = [A_l | 1] (MN-major)
                    const tc::Desc dAm = tc::mnmajor(sDZ[cur], 128);
                    const tc::Desc dBm = tc::mnmajor(tc::smem_u32(sm + L.h_off[l]), 128);
                    const uint32_t idw = tc::make_idesc_bf16(128, l == 0 ? 16 : 80, 1, 1);
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        tc::mma_bf16(tmem + L.tm_dw[l], dAm.adv(s * tc::KSTEP_MN).u64(), dBm.adv(s * tc::KSTEP_MN).u64(), idw,
                                     !(first_tile && s == 0));
                    if (l >= 1) {
                        // dA_l = dZ_l W_l: A = dZ (K-major over channels), B = W_l tile re-read MN-major (N = K_l)
                        const tc::Desc dAk = tc::kmajor(sDZ[cur], 128);
                        const tc::Desc dWm = tc::mnmajor(tc::smem_u32(sm + L.w_off[l]), N);
                        const uint32_t idh = tc::make_idesc_bf16(128, a.dims[l], 0, 1);
                        for (int s = 0; s < N / 16; ++s)
                            tc::mma_bf16(tmem + (l & 1) * 64, dAk.adv(s * KS128).u64(), dWm.adv(s * tc::KSTEP_MN).u64(), idh, s > 0);
                    }
                    tc::mma_commit(&mbar_mma);
                }
                __syncwarp();
            }
            tc::mbar_wait(&mbar_mma, ph_mma); ph_mma ^= 1;
            tc::fence_after_sync();
            if (l >= 1) {
                float v[32];
                tc::tmem_ld32(tlane + (l & 1) * 64 + half * 32, v);
                const uint8_t* zt = sm + L.z_off[l];
                uint8_t* dzn = sm + (cur ? L.dza : L.dzb);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const uint4 zz = *reinterpret_cast<const uint4*>(zt + (half * 4 + c8) * (128 * 16) + row * 16);
                    const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
                    float d[8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float z0 = __uint_as_float(zw[c] << 16), z1 = __uint_as_float(zw[c] & 0xffff0000u);
                        d[2 * c] = v[c8 * 8 + 2 * c] * gelu_tanh_grad_fast(z0);
                        d[2 * c + 1] = v[c8 * 8 + 2 * c + 1] * gelu_tanh_grad_fast(z1);
                    }
                    uint4 o;
                    o.x = tc::pack_bf16(d[0], d[1]); o.y = tc::pack_bf16(d[2], d[3]);
                    o.z = tc::pack_bf16(d[4], d[5]); o.w = tc::pack_bf16(d[6], d[7]);
                    *reinterpret_cast<uint4*>(dzn + (half * 4 + c8) * (128 * 16) + row * 16) = o;
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
                cur ^= 1;
            }
        }
        first_tile = false;
        tc::fence_before_sync();
        __syncthreads();
    }

    // =============== flush the per-CTA weight-gradient accumulators ===============
    tc::fence_after_sync();
    float* mine = partial + (size_t)blockIdx.x * a.n_params;
    if (warp < 4) {
        const int n = warp * 32 + lane;                      // accumulator row = output channel
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const int K = a.dims[l], N = a.dims[l + 1];
            if (l == 0) {
                float t[16];
                tc::tmem_ld16(tlane + L.tm_dw[0], t);
                if (n < N && !first_tile) {
                    for (int k = 0; k < 6; ++k) mine[a.w_off[0] + n * K + k] = t[k] + t[6 + k];     // hi and lo parts share the weight
                    mine[a.b_off[0] + n] = t[12];
                }
            } else {
                float t0[32], t1[32], t2[16];
                tc::tmem_ld32(tlane + L.tm_dw[l], t0);
                tc::tmem_ld32(tlane + L.tm_dw[l] + 32, t1);
                tc::tmem_ld16(tlane + L.tm_dw[l] + 64, t2);
                if (n < N && !first_tile) {
                    for (int k = 0; k < 32; ++k) { mine[a.w_off[l] + n * K + k] = t0[k]; mine[a.w_off[l] + n * K + 32 + k] = t1[k]; }
                    mine[a.b_off[l] + n] = t2[0];
                }
            }
        }
    }
    if (first_tile) {                                        // CTA without tiles (cannot happen with grid <= ntiles)
        for (int i = tid; i < a.n_params; i += TTHREADS) mine[i] = 0.f;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int gno_bwd_reduce(const float* partial, int nparts, int n_params, float* d_params, cudaStream_t st);   // gno_bwd.cu

bool gno_backward_bf16_supported(const GnoArgs& a) {
    if (a.transform != 0 && a.transform != 3) return false;
    if (a.n_layers < 2 || a.n_layers > 4) return false;
    for (int l = 1; l < a.n_layers; ++l) if (a.dims[l] != 64) return false;
    const int Cout = a.dims[a.n_layers];
    if (Cout != 32) return false;                       // staging tiles are sized for 32-wide features
    if (a.f_y && a.c_f != 32) return false;
    return true;
}

int gno_backward_bf16(const GnoArgs& a, const float* d_out, void* ws, size_t ws_bytes, float* d_params, float* d_f,
                      cudaStream_t st) {
    if (d_f) GAOT_CUDA(cudaMemsetAsync(d_f, 0, (size_t)a.n_src * a.c_f * sizeof(float), st));
    if (a.E == 0) {
        GAOT_CUDA(cudaMemsetAsync(d_params, 0, (size_t)a.n_params * sizeof(float), st));
        return GAOT_OK;
    }
    Arena ar(ws, ws_bytes);
    const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
    float* partial = ar.take<float>((size_t)grid * a.n_params);
    if (!ar.ok()) { set_error("gno_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const TcBwdLayout L = tc_bwd_layout(a);
    const size_t smem = (size_t)L.total_bytes;
    if (smem > 227 * 1024) { set_error("gno bf16 backward: shared memory %zu B too large", smem); return GAOT_ERR_UNSUPPORTED; }
    const int Cout = a.dims[a.n_layers];
    {
        GAOT_TIME_KERNEL("gno_bwd", st, (double)a.E * (16.0 + 12.0 + 4.0 * a.c_f) + (double)a.nq * (12.0 + 8.0 * Cout) + (double)a.n_src * 4.0 * a.c_f);
#define GAOT_TCB_CASE(NL)                                                                                              \
    case NL:                                                                                                           \
        GAOT_CUDA(cudaFuncSetAttribute(gno_bwd_tc_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gno_bwd_tc_kernel<NL><<<grid, TTHREADS, smem, st>>>(a, L, d_out, d_f, partial);                                 \
        break;
        switch (a.n_layers) { GAOT_TCB_CASE(2) GAOT_TCB_CASE(3) GAOT_TCB_CASE(4) default: break; }
#undef GAOT_TCB_CASE
    }
    GAOT_LAUNCH_CHECK();
    return gno_bwd_reduce(partial, grid, a.n_params, d_params, st);
}

}  // namespace gaot
