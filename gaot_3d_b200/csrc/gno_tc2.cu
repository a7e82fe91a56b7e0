// gno_tc2.cu -- second-generation tensor-core GNO kernels (same contract as gno_bf16.cu, which stays as
// the path for shapes outside this envelope).
//
// What ncu said about the first generation (profiles/r01e_gno_*): tensor pipe 4 %, issue slots 18-39 %,
// 8-16 resident warps: one 128-edge tile is a strictly serial chain  MMA -> commit -> wake -> tcgen05.ld ->
// GELU -> st.shared -> fence.proxy.async -> barrier  per layer, and nothing else runs on the SM meanwhile.
// The GELU itself costs ~9 issue slots per element in fp32.
//
// FORWARD (gno_fwd_tc2_kernel): one CTA per SM, FOUR independent tile streams per CTA (128 threads = one
// edge row per thread each; named barriers, own mbarriers, own 64 TMEM columns), sharing one copy of the
// weights.  Operands are FP16 (11-bit mantissa: coordinates as hi+lo halves, activations in [-,+]65504),
// accumulation FP32 in TMEM; the GELU runs on packed half2 (HFMA2 + MUFU.TANH.F16): ~5 issue slots per
// element, and its output is already the next layer's A operand.  Feature rows land (cp.async.bulk) in a
// padded staging tile that is multiplied IN PLACE and reduced per CSR segment; rowptr is prefetched per row
// so that the segmented mean has no dependent global loads.
//
// BACKWARD (gno_bwd_tc2_kernel): same dataflow as gno_bwd_tc_kernel (recompute, dW/db accumulated in TMEM
// across tiles, dA by MN-major weight re-read) with bf16 MMA operands (gradients need the exponent range), but
//   * gelu'(z) is produced together with gelu(z) during the recompute (one tanh instead of two, packed half2
//     arithmetic) and kept as an f16 tile; the backward layers only multiply;
//   * d_out / f_y rows are read straight from global memory, dZ is single-buffered: the decoder-sized MLP
//     (3 layers) fits 2 CTAs per SM (256 TMEM columns, 108 KB shared memory each).
#include <cuda_fp16.h>
#include "gno_common.cuh"
#include "tc05.cuh"

namespace gaot {

constexpr int T2E = 128;           // edges per tile = M of one tcgen05.mma
constexpr int F2G = 4;             // tile streams per CTA (forward)
constexpr int F2T = 128 * F2G;
constexpr int FROW = 36;           // floats per staged feature row (32 + 4 padding: conflict-free float4 rows AND columns)

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;\n" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ __half2 tanh_h2(__half2 u) {
    uint32_t r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(h2_as_u32(u)));
    return u32_as_h2(r);
}
// gelu and its derivative from one tanh.  x*x is clamped so that |x| > 255 gives (x or 0, 1 or 0), not inf*0
__device__ __forceinline__ void gelu_and_grad_h2(__half2 x, __half2& g, __half2& dg) {
    const __half2 c1 = __float2half2_rn(0.0356774081f), c0 = __float2half2_rn(0.7978845608f), hf = __float2half2_rn(0.5f);
    const __half2 c3 = __float2half2_rn(0.1070322243f), one = __float2half2_rn(1.0f), cap = __float2half2_rn(16384.f);
    const __half2 x2 = __hmin2(__hmul2(x, x), cap);
    const __half2 u = __hmul2(x, __hfma2(x2, c1, c0));
    const __half2 t = tanh_h2(u);
    const __half2 hx = __hmul2(x, hf);
    g = __hfma2(hx, t, hx);
    const __half2 sech2 = __hfma2(__hneg2(t), t, one);
    const __half2 du = __hfma2(x2, c3, c0);
    dg = __hfma2(__hmul2(hx, sech2), du, __hfma2(t, hf, hf));
}
__device__ __forceinline__ uint32_t h2_to_bf16x2(__half2 h) {
    const float2 f = __half22float2(h);
    return tc::pack_bf16(f.x, f.y);
}
// instruction descriptor kind::f16 with F16 A/B operands (format field 0), FP32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// =====================================================================================================
// forward
// =====================================================================================================
struct Tc2Layout {
    int w_off[GNO_MAX_LAYERS];     // byte offsets of the f16 weight tiles [N x kpad]: [W (x 1/2 for l >= 1) | bias | 0]
    int kpad[GNO_MAX_LAYERS];      // 16 for layer 0 (12 coordinate halves + ones column), K + 16 otherwise (ones chunk + zero chunk)
    int grp_base, grp_stride;      // per-stream regions
    int total_bytes;
};
// per-stream region: [A0 4K | ACT 20K = 64 activation columns + ones chunk + zero chunk | F 18K | per-row ints x2 | segment table]
constexpr int G_A0 = 0, G_ACT = 4096, G_F = G_ACT + 20480, G_INTS = G_F + T2E * FROW * 4;
constexpr int G_ROWINTS = 3 * T2E;                       // s_qry, s_rb, s_re of one buffer
constexpr int G_STRIDE = 47104;
static_assert(G_INTS + (2 * G_ROWINTS + T2E + 4 + 8) * 4 <= G_STRIDE, "per-stream region too small");

static Tc2Layout tc2_layout(const GnoArgs& a) {
    Tc2Layout L;
    memset(&L, 0, sizeof(L));
    int off = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        L.kpad[l] = (l == 0) ? 16 : a.dims[l] + 16;
        L.w_off[l] = off; off += a.dims[l + 1] * L.kpad[l] * 2;
    }
    off = (off + 1023) / 1024 * 1024;
    L.grp_base = off; L.grp_stride = G_STRIDE;
    L.total_bytes = off + F2G * G_STRIDE;
    return L;
}

// 2 * gelu(x) on a packed pair: x + x * tanh(...).  The factor 1/2 lives in the next layer's weights (exact: power of two).
__device__ __forceinline__ __half2 gelu2x_h2(__half2 x) {
    const __half2 c1 = __float2half2_rn(0.0356774081f), c0 = __float2half2_rn(0.7978845608f);
    const __half2 x2 = __hmul2(x, x);
    const __half2 u = __hmul2(x, __hfma2(x2, c1, c0));
    return __hfma2(x, tanh_h2(u), x);
}

template <int NL>
__global__ void __launch_bounds__(F2T, 1)
gno_fwd_tc2_kernel(const GnoArgs a, const Tc2Layout L, float* __restrict__ out, float* __restrict__ head_partial) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar_mma[F2G], mbar_f[F2G];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid >> 7, row = tid & 127, wg = warp & 3;

    // ---- weights -> f16 chunk-major K-major B tiles.  Layers >= 1 carry the 1/2 of the GELU; the bias of every
    //      layer is one more K column that meets a constant 1 in the A operand (no bias add in the epilogues) ----
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1], KP = L.kpad[l];
        const float* W = a.params + a.w_off[l];
        uint8_t* Ws = sm + L.w_off[l];
        for (int idx = tid; idx < N * KP; idx += F2T) {
            const int n = idx / KP, k = idx - n * KP;
            float v;
            if (l == 0) v = k < 12 ? W[n * K + (k % 6)] : (k == 12 ? a.params[a.b_off[0] + n] : 0.f);   // [hi(6) | lo(6) | bias | 0 0 0]
            else v = k < K ? 0.5f * W[n * K + k] : (k == K ? a.params[a.b_off[l] + n] : 0.f);
            *reinterpret_cast<__half*>(Ws + tc::cm_off(N, n, k)) = __float2half_rn(v);
        }
    }
    if (row < T2E) {                               // constant ones / zero chunks behind the 64 activation columns
        uint8_t* act = sm + L.grp_base + g * L.grp_stride + G_ACT;
        *reinterpret_cast<uint4*>(act + 8 * (128 * 16) + row * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(act + 9 * (128 * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64 * F2G);
    if (tid == 0) {
        for (int i = 0; i < F2G; ++i) { tc::mbar_init(&mbar_mma[i], 1); tc::mbar_init(&mbar_f[i], 128); }
        tc::mbar_fence_init();
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_g = tmem_base_s + g * 64;                            // this stream's accumulator columns
    const uint32_t tlane = tmem_g + ((uint32_t)(wg * 32) << 16);
    uint32_t ph_mma = 0, ph_f = 0;

    uint8_t* G0 = sm + L.grp_base + g * L.grp_stride;
    uint8_t* A0 = G0 + G_A0;
    uint8_t* ACT = G0 + G_ACT;
    float* F = reinterpret_cast<float*>(G0 + G_F);
    int* rowints = reinterpret_cast<int*>(G0 + G_INTS);          // [2][s_qry | s_rb | s_re]
    int* seg_first = rowints + 2 * G_ROWINTS;                    // [T2E + 1]
    int* s_misc = seg_first + T2E + 4;                           // [8]

    const bool has_f = a.f_y != nullptr;
    const tc::Desc dA0 = tc::kmajor(tc::smem_u32(A0), 128);
    const tc::Desc dAct = tc::kmajor(tc::smem_u32(ACT), 128);
    constexpr uint32_t KS128 = tc::kstep_kmajor(128);
    constexpr int Cout = 32;

    auto issue_layer = [&](int l) {
        if (wg == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const int N = a.dims[l + 1], KP = L.kpad[l];
                const tc::Desc dA = (l == 0) ? dA0 : dAct;
                const tc::Desc dW = tc::kmajor(tc::smem_u32(sm + L.w_off[l]), N);
                const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
                const uint32_t ksw = tc::kstep_kmajor(N);
                for (int s = 0; s < KP / 16; ++s)
                    tc::mma_bf16(tmem_g, dA.adv(s * KS128).u64(), dW.adv(s * ksw).u64(), idesc, s > 0);
                tc::mma_commit(&mbar_mma[g]);
            }
            __syncwarp();
        }
    };
    // layer-0 operand row [y hi | x hi | y lo | x lo | 1 | 0] in f16 + the per-row ints of tile buffer `buf`
    auto write_row = [&](int buf, const float (&c6)[6], int qry, int rb, int re) {
        __half hi[6], lo[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            hi[j] = __float2half_rn(c6[j]);
            lo[j] = __float2half_rn(c6[j] - __half2float(hi[j]));
        }
        uint4 c0, c1;
        c0.x = h2_as_u32(__halves2half2(hi[0], hi[1])); c0.y = h2_as_u32(__halves2half2(hi[2], hi[3]));
        c0.z = h2_as_u32(__halves2half2(hi[4], hi[5])); c0.w = h2_as_u32(__halves2half2(lo[0], lo[1]));
        c1.x = h2_as_u32(__halves2half2(lo[2], lo[3])); c1.y = h2_as_u32(__halves2half2(lo[4], lo[5])); c1.z = 0x00003C00u; c1.w = 0u;
        *reinterpret_cast<uint4*>(A0 + row * 16) = c0;
        *reinterpret_cast<uint4*>(A0 + 128 * 16 + row * 16) = c1;
        int* ri = rowints + buf * G_ROWINTS;
        ri[row] = qry; ri[T2E + row] = rb; ri[2 * T2E + row] = re;
    };

    const int tstep = gridDim.x * F2G;
    int tile = blockIdx.x * F2G + g;
    // ---- software pipeline: indices of tile k+1 are loaded during layer 0 of tile k, its coordinates during
    //      layer 1, its layer-0 operand is written before the last layer -> no dependent global latency per tile ----
    int src = 0, qry = -1, qprev = -2;
    if (tile < a.ntiles) {
        const int e0 = tile * T2E;
        const bool valid = row < min(T2E, a.E - e0);
        if (valid) { src = a.csr_src[e0 + row]; qry = a.csr_qry[e0 + row]; }
        if (valid && row > 0) qprev = a.csr_qry[e0 + row - 1];
        float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int rb = 0, re = 0;
        if (valid) {
            const float* py = a.y_pos + (size_t)src * 3;
            const float* px = a.x_pos + (size_t)qry * 3;
            c6[0] = py[0]; c6[1] = py[1]; c6[2] = py[2]; c6[3] = px[0]; c6[4] = px[1]; c6[5] = px[2];
            rb = a.rowptr[qry]; re = a.rowptr[qry + 1];
        }
        write_row(0, c6, qry, rb, re);
    }

    for (int it = 0; tile < a.ntiles; tile += tstep, ++it) {
        const int buf = it & 1;
        const int e0 = tile * T2E;
        const int ne = min(T2E, a.E - e0);
        const bool valid = row < ne;
        const int* s_qry = rowints + buf * G_ROWINTS;
        const int* s_rb = s_qry + T2E;
        const int* s_re = s_rb + T2E;
        // ---- feature rows -> padded staging tile: 8 lanes per 128-byte row, 4 rows per warp instruction ----
        if (has_f) {
            const uint32_t fbase = tc::smem_u32(F) + (uint32_t)(wg * 32) * (FROW * 4) + (lane & 7) * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j * 4 + (lane >> 3);
                const int sr = __shfl_sync(0xffffffffu, src, r);
                tc::cp_async16(fbase + (uint32_t)r * (FROW * 4), a.f_y + (size_t)sr * Cout + (lane & 7) * 4, wg * 32 + r < ne);
            }
            tc::cp_async_mbar_arrive_noinc(&mbar_f[g]);
        }
        const bool head = valid && (row == 0 || qprev != qry);
        const unsigned bm = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_misc[wg] = (int)bm;
        tc::fence_async_smem();
        tc::fence_before_sync();
        group_bar(g);
        issue_layer(0);
        {   // segment table (visible to the reduction after the per-layer barriers)
            int before = 0;
            for (int w = 0; w < wg; ++w) before += __popc((unsigned)s_misc[w]);
            if (head) seg_first[before + __popc(bm & ((1u << lane) - 1u))] = row;
            if (row == 0) {
                const int nseg = __popc((unsigned)s_misc[0]) + __popc((unsigned)s_misc[1]) +
                                 __popc((unsigned)s_misc[2]) + __popc((unsigned)s_misc[3]);
                s_misc[4] = nseg;
                seg_first[nseg] = ne;
            }
        }
        // pipeline step A: indices of the next tile
        const int ntile = tile + tstep;
        const bool nvalid = ntile < a.ntiles && row < min(T2E, a.E - ntile * T2E);
        int nsrc = 0, nqry = -1, nprev = -2;
        if (nvalid) { nsrc = a.csr_src[ntile * T2E + row]; nqry = a.csr_qry[ntile * T2E + row]; }
        if (nvalid && row > 0) nprev = a.csr_qry[ntile * T2E + row - 1];
        float nc6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int nrb = 0, nre = 0;

#pragma unroll
        for (int l = 0; l < NL; ++l) {
            // pipeline step C (the layer-0 MMA of this tile has long read A0)
            if (l == NL - 1 && ntile < a.ntiles) write_row(buf ^ 1, nc6, nqry, nrb, nre);
            tc::mbar_wait(&mbar_mma[g], ph_mma); ph_mma ^= 1;
            tc::fence_after_sync();
            if (l < NL - 1) {
                // this thread: edge `row`, all 64 hidden columns
                uint32_t v0[32], v1[32];
                tc::tmem_ld32_nowait(tlane, v0);
                tc::tmem_ld32_nowait(tlane + 32, v1);
                tc::tmem_wait_ld();
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = (c8 & 3) * 8 + 2 * j;
                        const float lo_ = __uint_as_float(c8 < 4 ? v0[c] : v1[c]);
                        const float hi_ = __uint_as_float(c8 < 4 ? v0[c + 1] : v1[c + 1]);
                        o[j] = h2_as_u32(gelu2x_h2(__floats2half2_rn(lo_, hi_)));
                    }
                    *reinterpret_cast<uint4*>(ACT + c8 * (128 * 16) + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                group_bar(g);
                issue_layer(l + 1);
                if (l == 0 && nvalid) {                                                           // pipeline step B
                    const float* py = a.y_pos + (size_t)nsrc * 3;
                    const float* px = a.x_pos + (size_t)nqry * 3;
                    nc6[0] = py[0]; nc6[1] = py[1]; nc6[2] = py[2]; nc6[3] = px[0]; nc6[4] = px[1]; nc6[5] = px[2];
                    nrb = a.rowptr[nqry]; nre = a.rowptr[nqry + 1];
                }
            } else {
                // last layer: (* f_y[src]) in fp32, written in place over the staged feature row
                float v[32];
                tc::tmem_ld32(tlane, v);
                if (has_f) { tc::mbar_wait(&mbar_f[g], ph_f); ph_f ^= 1; }
                const float ew = (a.edge_w && valid) ? a.edge_w[e0 + row] : 1.0f;
                float* fr = F + row * FROW;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                    if (has_f) {
                        const float4 f = *reinterpret_cast<const float4*>(fr + c);
                        o.x *= f.x; o.y *= f.y; o.z *= f.z; o.w *= f.w;
                    }
                    o.x *= ew; o.y *= ew; o.z *= ew; o.w *= ew;       // attention weight of the edge (1 when there is none)
                    *reinterpret_cast<float4*>(fr + c) = o;
                }
                tc::fence_before_sync();
                group_bar(g);
                // segmented sums: 8 threads (one float4 of channels each) per segment, 16 segments at a time
                const int nseg = s_misc[4];
                const int quad = row & 7;
                const float4* F4 = reinterpret_cast<const float4*>(F);
                for (int s = row >> 3; s < nseg; s += 16) {
                    const int first = seg_first[s], lastE = seg_first[s + 1];
                    const int q = s_qry[first];
                    const int rb = s_rb[first], re = s_re[first];
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int e = first; e < lastE; ++e) {
                        const float4 t = F4[e * (FROW / 4) + quad];
                        sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
                    }
                    if (rb >= e0) {                                          // segment starts in this tile
                        if (re <= e0 + ne && a.reduce == 0) {
                            const float inv = 1.0f / (float)(re - rb);
                            sum.x *= inv; sum.y *= inv; sum.z *= inv; sum.w *= inv;
                        }
                        *reinterpret_cast<float4*>(out + (size_t)q * Cout + quad * 4) = sum;
                    } else {
                        *reinterpret_cast<float4*>(head_partial + (size_t)tile * Cout + quad * 4) = sum;
                    }
                }
            }
        }
        src = nsrc; qry = nqry; qprev = nprev;
        tc::fence_async_smem();        // generic accesses of F / A0 before the next tile's copies and MMA reads
        group_bar(g);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 64 * F2G);
}

int gno_fwd_fixup(const GnoArgs& a, float* out, const float* head_partial, cudaStream_t st);   // gno_fwd.cu

bool gno_forward_tc2_supported(const GnoArgs& a) {
    if (a.transform != 0 && a.transform != 3) return false;
    if (a.n_layers < 2 || a.n_layers > 5) return false;
    for (int l = 1; l < a.n_layers; ++l) if (a.dims[l] != 64) return false;
    if (a.dims[a.n_layers] != 32) return false;
    if (a.f_y && a.c_f != 32) return false;
    return true;
}

int gno_forward_tc2(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st) {
    constexpr int Cout = 32;
    GAOT_CUDA(cudaMemsetAsync(out, 0, (size_t)a.nq * Cout * sizeof(float), st));
    if (a.E == 0) return GAOT_OK;
    Arena ar(ws, ws_bytes);
    float* head_partial = ar.take<float>((size_t)a.ntiles * Cout);
    if (!ar.ok()) { set_error("gno_forward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const Tc2Layout L = tc2_layout(a);
    const size_t smem = (size_t)L.total_bytes;
    if (smem > 227 * 1024) { set_error("gno tc2: shared memory %zu B too large", smem); return GAOT_ERR_UNSUPPORTED; }
    const int want = (a.ntiles + F2G - 1) / F2G;
    const int grid = want < kNumSMs ? want : kNumSMs;
    {
        GAOT_TIME_KERNEL("gno_fwd", st, (double)a.E * (16.0 + 12.0 + 4.0 * a.c_f) + (double)a.nq * (12.0 + 4.0 * Cout));
#define GAOT_TC2_CASE(NL)                                                                                              \
    case NL:                                                                                                           \
        GAOT_CUDA(cudaFuncSetAttribute(gno_fwd_tc2_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gno_fwd_tc2_kernel<NL><<<grid, F2T, smem, st>>>(a, L, out, head_partial);                                       \
        break;
        switch (a.n_layers) { GAOT_TC2_CASE(2) GAOT_TC2_CASE(3) GAOT_TC2_CASE(4) GAOT_TC2_CASE(5) default: break; }
#undef GAOT_TC2_CASE
    }
    GAOT_LAUNCH_CHECK();
    return gno_fwd_fixup(a, out, head_partial, st);
}

// =====================================================================================================
// backward
// =====================================================================================================
constexpr int B2T = 256;

struct Tc2BwdLayout {
    int w_off[GNO_MAX_LAYERS], kpad[GNO_MAX_LAYERS], h_off[GNO_MAX_LAYERS], gp_off[GNO_MAX_LAYERS];
    int tm_dw[GNO_MAX_LAYERS];
    int a0, dz, total_bytes, tmem_cols;
};

static Tc2BwdLayout tc2_bwd_layout(const GnoArgs& a) {
    Tc2BwdLayout L;
    memset(&L, 0, sizeof(L));
    int off = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        L.kpad[l] = (l == 0) ? 16 : a.dims[l] + 16;           // [W | bias | 0]: the bias meets the ones column / chunk of the A operand
        L.w_off[l] = off; off += a.dims[l + 1] * L.kpad[l] * 2;
    }
    off = (off + 1023) / 1024 * 1024;
    L.a0 = off; off += T2E * 16 * 2;
    // dZ tile (single buffer).  An M = 128 read of it (dW GEMM, channels padded to 128) touches the 16 KB behind
    // it: the activation tiles follow, so that read stays inside the allocation.
    L.dz = off; off += T2E * 64 * 2;
    for (int l = 1; l < a.n_layers; ++l) { L.h_off[l] = off; off += T2E * 80 * 2; }      // bf16 [128 x (64 + ones chunk + zero chunk)]
    for (int l = 1; l < a.n_layers; ++l) { L.gp_off[l] = off; off += T2E * 64 * 2; }     // f16 gelu'(z_{l-1})
    L.h_off[0] = L.a0;
    L.total_bytes = off;
    int col = 64;
    for (int l = 0; l < a.n_layers; ++l) { L.tm_dw[l] = col; col += (l == 0) ? 32 : 80; }
    L.tmem_cols = col <= 256 ? 256 : 512;
    return L;
}

template <int NL>
__global__ void __launch_bounds__(B2T, (NL <= 3 ? 2 : 1))
gno_bwd_tc2_kernel(const GnoArgs a, const Tc2BwdLayout L, const float* __restrict__ d_out, float* __restrict__ d_f,
                   float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    uint8_t* A0 = sm + L.a0;
    uint8_t* DZ = sm + L.dz;

#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1], KP = L.kpad[l];
        const float* W = a.params + a.w_off[l];
        uint8_t* Ws = sm + L.w_off[l];
        for (int idx = tid; idx < N * KP; idx += B2T) {
            const int n = idx / KP, k = idx - n * KP;
            float v;
            if (l == 0) v = k < 12 ? W[n * K + (k % 6)] : (k == 12 ? a.params[a.b_off[0] + n] : 0.f);
            else v = k < K ? W[n * K + k] : (k == K ? a.params[a.b_off[l] + n] : 0.f);
            *reinterpret_cast<__nv_bfloat16*>(Ws + tc::cm_off(N, n, k)) = __float2bfloat16(v);
        }
        if (l >= 1 && tid < T2E) {                 // constant ones / zero chunks of the activation tiles
            *reinterpret_cast<uint4*>(sm + L.h_off[l] + 8 * (128 * 16) + tid * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sm + L.h_off[l] + 9 * (128 * 16) + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)L.tmem_cols);
    if (tid == 0) { tc::mbar_init(&mbar_mma, 1); tc::mbar_fence_init(); }

    constexpr int Cout = 32, nc = 16;                       // output columns per thread
    const bool use_f_mul = (a.transform == 0);
    constexpr uint32_t KS128 = tc::kstep_kmajor(128);
    const uint32_t sDZ = tc::smem_u32(DZ);
    bool first_tile = true;

    // layer-0 operand row [y hi | x hi | y lo | x lo | 1 | 0] in bf16 (the 1 is both the bias input of the recompute and
    // the column that returns db_0 from the dW_0 GEMM; 0 for padding rows)
    auto write_a0 = [&](const float (&c6)[6], bool valid) {
        float hi[6], lo[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) { hi[j] = __bfloat162float(__float2bfloat16(c6[j])); lo[j] = c6[j] - hi[j]; }
        uint4 c0, c1;
        c0.x = tc::pack_bf16(hi[0], hi[1]); c0.y = tc::pack_bf16(hi[2], hi[3]);
        c0.z = tc::pack_bf16(hi[4], hi[5]); c0.w = tc::pack_bf16(lo[0], lo[1]);
        c1.x = tc::pack_bf16(lo[2], lo[3]); c1.y = tc::pack_bf16(lo[4], lo[5]);
        c1.z = valid ? 0x00003F80u : 0u;
        c1.w = 0u;
        *reinterpret_cast<uint4*>(A0 + row * 16) = c0;
        *reinterpret_cast<uint4*>(A0 + 128 * 16 + row * 16) = c1;
    };
    // loads only (no arithmetic on the results: a consumer would stall the in-order warp on the load latency)
    auto load_row = [&](int src, int qry, int e, float (&c6)[6], int& rb, int& re) {
        if (half == 0) {
            const float* py = a.y_pos + (size_t)src * 3;
            const float* px = a.x_pos + (size_t)qry * 3;
            c6[0] = py[0]; c6[1] = py[1]; c6[2] = py[2]; c6[3] = px[0]; c6[4] = px[1]; c6[5] = px[2];
        }
        if (a.edge_w) { rb = __float_as_int(a.edge_w[e]); re = 0; }
        else { rb = a.rowptr[qry]; re = a.rowptr[qry + 1]; }
    };
    // the factor every edge gradient carries: 1/count (mean), 1 (sum), or the edge's attention weight (`rb` then holds its bits)
    auto inv_count = [&](int rb, int re, bool valid) {
        if (!valid) return 0.f;
        if (a.edge_w) return __int_as_float(rb);
        return a.reduce == 0 ? 1.0f / (float)(re - rb) : 1.0f;
    };

    // ---- first tile of this CTA: synchronous staging; later tiles are staged by the software pipeline below ----
    int tile = blockIdx.x;
    int src = 0, qry = 0;
    float inv = 0.f;
    if (tile < a.ntiles) {
        const bool valid = row < min(T2E, a.E - tile * T2E);
        float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int rb = 0, re = 1;
        if (valid) { src = a.csr_src[tile * T2E + row]; qry = a.csr_qry[tile * T2E + row]; load_row(src, qry, tile * T2E + row, c6, rb, re); }
        inv = inv_count(rb, re, valid);
        if (half == 0) write_a0(c6, valid);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_mma = 0;

    for (; tile < a.ntiles; tile += gridDim.x) {
        const int e0 = tile * T2E;
        const int ne = min(T2E, a.E - e0);
        const bool valid = row < ne;
        // pipeline: indices of the next tile now, its coordinates / counts after the first layer, its operand row
        // after the last MMA of this tile
        const int ntile = tile + gridDim.x;
        const bool nvalid = ntile < a.ntiles && row < min(T2E, a.E - ntile * T2E);
        int nsrc = 0, nqry = 0;
        float nc6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int nrb = 0, nre = 1;

        // source feature rows -> dZ buffer (free until the output stage): 8 lanes per 128-byte row, 16-byte pieces
        // XOR-swizzled by the row so that the row-per-thread reads below are conflict-free
        if (use_f_mul) {
            const int wr = (warp & 3) * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = (half * 4 + j) * 4 + (lane >> 3);
                const int sr = __shfl_sync(0xffffffffu, src, r);
                tc::cp_async16(sDZ + (uint32_t)(wr + r) * 128u + (uint32_t)(((lane & 7) ^ ((wr + r) & 7)) * 16),
                               a.f_y + (size_t)sr * Cout + (lane & 7) * 4, wr + r < ne);
            }
            tc::cp_async_commit();
        }

        // =============== forward recompute: gelu AND gelu' ===============
        float kv[nc], gr[nc], fr[nc];
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
                        tc::mma_bf16(tmem, dA.adv(s * KS128).u64(), dW.adv(s * ksw).u64(), idesc, s > 0);
                    tc::mma_commit(&mbar_mma);
                }
                __syncwarp();
            }
            if (l == 0 && nvalid) { nsrc = a.csr_src[ntile * T2E + row]; nqry = a.csr_qry[ntile * T2E + row]; }
            if (l == 1 && nvalid) load_row(nsrc, nqry, ntile * T2E + row, nc6, nrb, nre);
            if (l == (NL >= 3 ? NL - 2 : NL - 1)) {
                // d_out rows of this tile (query-major CSR: neighbouring edges share them), issued one layer early
                const float4* gp4 = reinterpret_cast<const float4*>(d_out + (size_t)qry * Cout + half * nc);
#pragma unroll
                for (int j = 0; j < nc / 4; ++j) {
                    const float4 t = valid ? gp4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    gr[4 * j] = t.x; gr[4 * j + 1] = t.y; gr[4 * j + 2] = t.z; gr[4 * j + 3] = t.w;
                }
            }
            if (l == NL - 1 && use_f_mul) {
                // f_y rows were gathered into the (still unused) dZ buffer at the top of the tile; the barrier after the
                // previous layer made them visible.  Everyone must have its copy before dZ is written below.
                const int sw = row & 7;
#pragma unroll
                for (int j = 0; j < nc / 4; ++j) {
                    const float4 t = *reinterpret_cast<const float4*>(DZ + row * 128 + (((half * 4 + j) ^ sw) * 16));
                    fr[4 * j] = t.x; fr[4 * j + 1] = t.y; fr[4 * j + 2] = t.z; fr[4 * j + 3] = t.w;
                }
                __syncthreads();
            }
            tc::mbar_wait(&mbar_mma, ph_mma); ph_mma ^= 1;
            tc::fence_after_sync();
            if (l < NL - 1) {
                float v[32];
                tc::tmem_ld32(tlane + half * 32, v);
                uint8_t* gpt = sm + L.gp_off[l + 1];
                uint8_t* ht = sm + L.h_off[l + 1];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint32_t og[4], od[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c8 * 8 + 2 * j;
                        __half2 gg, dg;
                        gelu_and_grad_h2(__floats2half2_rn(v[c], v[c + 1]), gg, dg);
                        og[j] = h2_to_bf16x2(gg);
                        od[j] = h2_as_u32(dg);
                    }
                    *reinterpret_cast<uint4*>(ht + (half * 4 + c8) * (128 * 16) + row * 16) = make_uint4(og[0], og[1], og[2], og[3]);
                    *reinterpret_cast<uint4*>(gpt + (half * 4 + c8) * (128 * 16) + row * 16) = make_uint4(od[0], od[1], od[2], od[3]);
                }
                if (l == NL - 2) tc::cp_async_wait_all();          // this thread's pieces of the f_y gather have landed
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
            } else {
                tc::tmem_ld16(tlane + half * nc, kv);
            }
        }

        // =============== output-side gradients (fp32) ===============
        {
            if (a.d_edge_w && valid) {                  // d loss / d w_e = <d_out[q], k_e (* f_e)>: this thread's 16 of the 32 terms
                float dot = 0.f;
#pragma unroll
                for (int c = 0; c < nc; ++c) dot = fmaf(gr[c] * kv[c], use_f_mul ? fr[c] : 1.0f, dot);
                atomicAdd(a.d_edge_w + e0 + row, dot);
            }
#pragma unroll
            for (int c = 0; c < nc; ++c) gr[c] *= inv;
            if (use_f_mul) {
                if (d_f && valid) {
                    float* dst = d_f + (size_t)src * Cout + half * nc;
#pragma unroll
                    for (int j = 0; j < nc / 4; ++j)
                        atomicAdd(reinterpret_cast<float4*>(dst + 4 * j),
                                  make_float4(kv[4 * j] * gr[4 * j], kv[4 * j + 1] * gr[4 * j + 1],
                                              kv[4 * j + 2] * gr[4 * j + 2], kv[4 * j + 3] * gr[4 * j + 3]));
                }
#pragma unroll
                for (int c = 0; c < nc; ++c) gr[c] *= fr[c];
            }
#pragma unroll
            for (int c8 = 0; c8 < nc / 8; ++c8) {
                uint4 o;
                o.x = tc::pack_bf16(gr[c8 * 8], gr[c8 * 8 + 1]); o.y = tc::pack_bf16(gr[c8 * 8 + 2], gr[c8 * 8 + 3]);
                o.z = tc::pack_bf16(gr[c8 * 8 + 4], gr[c8 * 8 + 5]); o.w = tc::pack_bf16(gr[c8 * 8 + 6], gr[c8 * 8 + 7]);
                *reinterpret_cast<uint4*>(DZ + ((half * nc) / 8 + c8) * (128 * 16) + row * 16) = o;
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // =============== backward through the layers ===============
#pragma unroll
        for (int l = NL - 1; l >= 0; --l) {
            const int N = a.dims[l + 1];
            if (warp == 0) {
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    // dW_l (+ db_l): A = dZ^T (MN-major over the edge-major dZ tile), B = [A_l | 1] (MN-major)
                    const tc::Desc dAm = tc::mnmajor(sDZ, 128);
                    const tc::Desc dBm = tc::mnmajor(tc::smem_u32(sm + L.h_off[l]), 128);
                    const uint32_t idw = tc::make_idesc_bf16(128, l == 0 ? 16 : 80, 1, 1);
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        tc::mma_bf16(tmem + L.tm_dw[l], dAm.adv(s * tc::KSTEP_MN).u64(), dBm.adv(s * tc::KSTEP_MN).u64(), idw,
                                     !(first_tile && s == 0));
                    if (l >= 1) {
                        // dA_l = dZ_l W_l: A = dZ (K-major over channels), B = W_l tile re-read MN-major (N = K_l)
                        const tc::Desc dAk = tc::kmajor(sDZ, 128);
                        const tc::Desc dWm = tc::mnmajor(tc::smem_u32(sm + L.w_off[l]), N);
                        const uint32_t idh = tc::make_idesc_bf16(128, a.dims[l], 0, 1);
                        for (int s = 0; s < N / 16; ++s)
                            tc::mma_bf16(tmem, dAk.adv(s * KS128).u64(), dWm.adv(s * tc::KSTEP_MN).u64(), idh, s > 0);
                    }
                    tc::mma_commit(&mbar_mma);
                }
                __syncwarp();
            }
            tc::mbar_wait(&mbar_mma, ph_mma); ph_mma ^= 1;
            tc::fence_after_sync();
            if (l >= 1) {
                float v[32];
                tc::tmem_ld32(tlane + half * 32, v);
                const uint8_t* gpt = sm + L.gp_off[l];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const uint4 dd = *reinterpret_cast<const uint4*>(gpt + (half * 4 + c8) * (128 * 16) + row * 16);
                    const uint32_t dw[4] = {dd.x, dd.y, dd.z, dd.w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 gpf = __half22float2(u32_as_h2(dw[j]));
                        o[j] = tc::pack_bf16(v[c8 * 8 + 2 * j] * gpf.x, v[c8 * 8 + 2 * j + 1] * gpf.y);
                    }
                    *reinterpret_cast<uint4*>(DZ + (half * 4 + c8) * (128 * 16) + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
            }
        }
        first_tile = false;
        // every MMA of this tile has completed (A0 was last read by dW_0): stage the next tile's layer-0 operand
        if (ntile < a.ntiles && half == 0) write_a0(nc6, nvalid);
        src = nsrc; qry = nqry; inv = inv_count(nrb, nre, nvalid);
        tc::fence_async_smem();
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
        for (int i = tid; i < a.n_params; i += B2T) mine[i] = 0.f;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)L.tmem_cols);
}

int gno_bwd_reduce(const float* partial, int nparts, int n_params, float* d_params, cudaStream_t st);   // gno_bwd.cu

bool gno_backward_tc2_supported(const GnoArgs& a) {
    if (a.transform != 0 && a.transform != 3) return false;
    if (a.n_layers < 2 || a.n_layers > 4) return false;
    for (int l = 1; l < a.n_layers; ++l) if (a.dims[l] != 64) return false;
    if (a.dims[a.n_layers] != 32) return false;
    if (a.f_y && a.c_f != 32) return false;
    return true;
}

int gno_backward_tc2(const GnoArgs& a, const float* d_out, void* ws, size_t ws_bytes, float* d_params, float* d_f,
                     cudaStream_t st) {
    if (d_f) GAOT_CUDA(cudaMemsetAsync(d_f, 0, (size_t)a.n_src * a.c_f * sizeof(float), st));
    if (a.d_edge_w && a.E > 0) GAOT_CUDA(cudaMemsetAsync(a.d_edge_w, 0, (size_t)a.E * sizeof(float), st));
    if (a.E == 0) {
        GAOT_CUDA(cudaMemsetAsync(d_params, 0, (size_t)a.n_params * sizeof(float), st));
        return GAOT_OK;
    }
    Arena ar(ws, ws_bytes);
    const int per_sm = a.n_layers <= 3 ? 2 : 1;
    const int grid = a.ntiles < per_sm * kNumSMs ? a.ntiles : per_sm * kNumSMs;
    float* partial = ar.take<float>((size_t)grid * a.n_params);
    if (!ar.ok()) { set_error("gno_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const Tc2BwdLayout L = tc2_bwd_layout(a);
    const size_t smem = (size_t)L.total_bytes;
    if (smem > 227 * 1024) { set_error("gno tc2 backward: shared memory %zu B too large", smem); return GAOT_ERR_UNSUPPORTED; }
    constexpr int Cout = 32;
    {
        GAOT_TIME_KERNEL("gno_bwd", st, (double)a.E * (16.0 + 12.0 + 4.0 * a.c_f) + (double)a.nq * (12.0 + 8.0 * Cout) + (double)a.n_src * 4.0 * a.c_f);
#define GAOT_TC2B_CASE(NL)                                                                                              \
    case NL:                                                                                                            \
        GAOT_CUDA(cudaFuncSetAttribute(gno_bwd_tc2_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gno_bwd_tc2_kernel<NL><<<grid, B2T, smem, st>>>(a, L, d_out, d_f, partial);                                      \
        break;
        switch (a.n_layers) { GAOT_TC2B_CASE(2) GAOT_TC2B_CASE(3) GAOT_TC2B_CASE(4) default: break; }
#undef GAOT_TC2B_CASE
    }
    GAOT_LAUNCH_CHECK();
    return gno_bwd_reduce(partial, grid, a.n_params, d_params, st);
}

}  // namespace gaot
