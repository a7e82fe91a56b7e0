// dense.cu -- token-level dense layers of the latent transformer on tcgen05 tensor cores.
// Replaces the fp32 cuBLAS SGEMMs behind nn.Linear in reference src/model/layers/attn.py
// (q/k/v/o_proj :104-106,:129; FFN w1/w2/w3 :163; skip_proj :223) and gaot_3d.py:205 (patch_linear):
// with allow_tf32 = False those run on the FP32 SIMT pipe (~30 ms of a 70 ms step at S = 16384).
//
// One kernel, C[M,N] = sum_k A(m,k) * B(n,k), BF16 operands / FP32 accumulation in TMEM.  Operands
// stay in their natural row-major global layout; the loader (LDG -> registers -> STS; fp32 converted on the fly) writes
// "chunk-major" shared tiles (tc05.cuh) that serve as K-major or MN-major tensor-core operands, so the
// three products of a linear layer need no transposed copies:
//     forward   y  = x  W^T     A = x  [M,K]  K-major     B = W [N,K]  K-major
//     d input   dx = dy W       A = dy [M,N]  K-major     B = W [N,K]  MN-major (contraction over N)
//     d weight  dW = dy^T x     A = dy [M,N]  MN-major    B = x [M,K]  MN-major (contraction over M, split-K)
// CTA = 128x128 output tile, 256 threads: all threads stage operands (coalesced float4 loads, 8-byte
// conflict-free shared stores into slabs padded by 16 B), one elected thread issues tcgen05.mma
// (4 x K16 per 64-deep block) into a 3-stage ring released by tcgen05.commit -> mbarrier; 2 CTAs/SM.  (The bf16
// forward / d-input products normally run on the persistent gemm2 kernel further down.)
// Epilogue: tcgen05.ld -> warp-private XOR-swizzled transpose -> coalesced 128 B row segments with
// fused bias / residual / accumulate, fp32 or bf16 output.
#include "common.cuh"
#include "tc05.cuh"
#include <algorithm>
#include <cstdlib>
#include <cuda.h>

namespace gaot {

using bf16 = __nv_bfloat16;

namespace dn {
constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3, THREADS = 256;
constexpr uint32_t SLAB_K = 128 * 16 + 16;      // K-major operand: 8 slabs (one per 8-wide K chunk) of 128 rows
constexpr uint32_t SLAB_MN = 64 * 16 + 16;      // MN-major operand: 16 slabs (one per 8-wide M/N chunk) of 64 K rows
constexpr uint32_t OP_BYTES = 16 * SLAB_MN;     // 16640 >= 8 * SLAB_K = 16512
constexpr uint32_t STAGE_BYTES = 2 * OP_BYTES;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES;
}

struct GemmArgs {
    const void* A; const void* A2; const void* B;
    int64_t lda, lda2, ldb;
    int64_t M, N, K;            // output rows, output columns, contraction length
    int64_t k_split;            // contraction index where A switches to A2 (K-major A only; multiple of BK), or K
    const float* bias;          // [N] or null
    const float* residual;      // [M, ldr] or null
    int64_t ldr;
    void* C; int64_t ldc;
    int c_bf16;                 // output dtype
    int accumulate;             // C += result (fp32 output only)
    int kb_per_split;           // split-K: K blocks per blockIdx.z; partial results go to `partial`
    float* partial;             // [splits, M, N] fp32 or null
};

// ---- operand staging: global (row-major, fp32 or bf16) -> registers -> bf16 chunk-major shared tile ----
// K-major tile: 128 (M/N) rows x 64 contraction columns.  MN-major tile: 64 contraction rows x 128 (M/N) columns.
template <typename T, bool MN> struct Stager;

template <bool MN> struct Stager<float, MN> {
    float4 r[8];
    // row-major source: element (row, col) at src[row * ld + col]
    __device__ __forceinline__ void load(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t col0,
                                         int64_t row_lim, int64_t col_lim, int tid) {
        constexpr int F4_PER_ROW = MN ? 32 : 16;
        constexpr int ROWS_PER_PASS = dn::THREADS / F4_PER_ROW;
        const int f4 = tid % F4_PER_ROW, rr = tid / F4_PER_ROW;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int64_t row = row0 + p * ROWS_PER_PASS + rr, col = col0 + f4 * 4;
            if (row < row_lim && col < col_lim)      // col_lim is a multiple of 4 (checked by the host)
                r[p] = __ldg(reinterpret_cast<const float4*>(src + row * ld + col));
            else
                r[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __device__ __forceinline__ void store(uint8_t* tile, int tid) const {
        constexpr int F4_PER_ROW = MN ? 32 : 16;
        constexpr int ROWS_PER_PASS = dn::THREADS / F4_PER_ROW;
        constexpr uint32_t SLAB = MN ? dn::SLAB_MN : dn::SLAB_K;
        const int f4 = tid % F4_PER_ROW, rr = tid / F4_PER_ROW;
        uint8_t* base = tile + (f4 >> 1) * SLAB + (f4 & 1) * 8;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            uint2 v;
            v.x = tc::pack_bf16(r[p].x, r[p].y);
            v.y = tc::pack_bf16(r[p].z, r[p].w);
            *reinterpret_cast<uint2*>(base + (p * ROWS_PER_PASS + rr) * 16) = v;
        }
    }
};

// bf16 sources need no conversion, but they are register-staged too (LDG.128 -> STS.128): cp.async (LDGSTS) with
// 16-byte pieces scattered over the padded slabs sustains only one warp instruction per ~150 cycles (measured on B200),
// which made every bf16 product loader-bound.
template <bool MN> struct Stager<bf16, MN> {
    uint4 r[4];
    __device__ __forceinline__ void load(const bf16* __restrict__ src, int64_t ld, int64_t row0, int64_t col0,
                                         int64_t row_lim, int64_t col_lim, int tid) {
        constexpr int CH_PER_ROW = MN ? 16 : 8;
        constexpr int ROWS_PER_PASS = dn::THREADS / CH_PER_ROW;
        const int ch = tid % CH_PER_ROW, rr = tid / CH_PER_ROW;
        const int64_t col = col0 + ch * 8;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t row = row0 + p * ROWS_PER_PASS + rr;
            r[p] = (row < row_lim && col < col_lim) ? __ldg(reinterpret_cast<const uint4*>(src + row * ld + col)) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
    __device__ __forceinline__ void store(uint8_t* tile, int tid) const {
        constexpr int CH_PER_ROW = MN ? 16 : 8;
        constexpr int ROWS_PER_PASS = dn::THREADS / CH_PER_ROW;
        constexpr uint32_t SLAB = MN ? dn::SLAB_MN : dn::SLAB_K;
        const int ch = tid % CH_PER_ROW, rr = tid / CH_PER_ROW;
        uint8_t* base = tile + ch * SLAB;
#pragma unroll
        for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(base + (p * ROWS_PER_PASS + rr) * 16) = r[p];
    }
};

template <typename T> struct is_bf16 { static constexpr bool value = false; };
template <> struct is_bf16<bf16> { static constexpr bool value = true; };

template <typename TA, typename TB, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(dn::THREADS, 2)
gemm_tc_kernel(const GemmArgs g) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t mbar_free[dn::STAGES];
    __shared__ uint64_t mbar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * dn::BN, m0 = (int64_t)blockIdx.y * dn::BM;
    const int nkb_total = (int)((g.K + dn::BK - 1) / dn::BK);
    const int kb0 = blockIdx.z * g.kb_per_split;
    const int nkb = min(g.kb_per_split, nkb_total - kb0);

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < dn::STAGES; ++s) tc::mbar_init(&mbar_free[s], 1);
        tc::mbar_init(&mbar_done, 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    constexpr uint32_t idesc = tc::make_idesc_bf16(dn::BM, dn::BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t A_LBO = A_MN ? 128u : dn::SLAB_K, A_SBO = A_MN ? dn::SLAB_MN : 128u, A_KSTEP = A_MN ? 256u : 2u * dn::SLAB_K;
    constexpr uint32_t B_LBO = B_MN ? 128u : dn::SLAB_K, B_SBO = B_MN ? dn::SLAB_MN : 128u, B_KSTEP = B_MN ? 256u : 2u * dn::SLAB_K;
    const uint32_t sm_base = tc::smem_u32(sm);

    // register-staged operands (fp32 converted on the fly, bf16 copied): one K block ahead
    Stager<TA, A_MN> sa;
    Stager<TB, B_MN> sb;
    auto fetch_regs = [&](int kb) {
        const int64_t k0 = (int64_t)(kb0 + kb) * dn::BK;
        if (A_MN) sa.load((const TA*)g.A, g.lda, k0, m0, g.K, g.M, tid);
        else if (k0 < g.k_split) sa.load((const TA*)g.A, g.lda, m0, k0, g.M, g.k_split, tid);
        else sa.load((const TA*)g.A2, g.lda2, m0, k0 - g.k_split, g.M, g.K - g.k_split, tid);
        if (B_MN) sb.load((const TB*)g.B, g.ldb, k0, n0, g.K, g.N, tid);
        else      sb.load((const TB*)g.B, g.ldb, n0, k0, g.N, g.K, tid);
    };
    if (nkb > 0) fetch_regs(0);
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % dn::STAGES;
        if (kb >= dn::STAGES) tc::mbar_wait(&mbar_free[s], (uint32_t)((kb / dn::STAGES - 1) & 1));   // MMA(kb - STAGES) has read the stage
        uint8_t* stA = sm + s * dn::STAGE_BYTES;
        sa.store(stA, tid);
        sb.store(stA + dn::OP_BYTES, tid);
        if (kb + 1 < nkb) fetch_regs(kb + 1);
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const tc::Desc dA = tc::make_desc2(sm_base + s * dn::STAGE_BYTES, A_LBO, A_SBO);
                const tc::Desc dB = tc::make_desc2(sm_base + s * dn::STAGE_BYTES + dn::OP_BYTES, B_LBO, B_SBO);
#pragma unroll
                for (int ks = 0; ks < dn::BK / 16; ++ks)
                    tc::mma_bf16(tmem, dA.adv(ks * A_KSTEP).u64(), dB.adv(ks * B_KSTEP).u64(), idesc, (kb | ks) != 0);
                tc::mma_commit(kb + 1 == nkb ? &mbar_done : &mbar_free[s]);
            }
            __syncwarp();
        }
    }

    // ---- epilogue: warp w owns TMEM lanes 32*(w%4).. (+32 rows) and columns 64*(w/4).. (+64) ----
    if (nkb > 0) {
        tc::mbar_wait(&mbar_done, 0);
        tc::fence_after_sync();
    }
    const int lg = warp & 3, ch = warp >> 2;
    float* tr = reinterpret_cast<float*>(sm) + warp * 1024;       // 32 x 32 fp32, float4 slots XOR-swizzled by row
    const bool partial = g.partial != nullptr;
    float* outp = partial ? g.partial + (size_t)blockIdx.z * g.M * g.N : nullptr;
#pragma unroll 1
    for (int slab = 0; slab < 2; ++slab) {
        const int c0 = ch * 64 + slab * 32;
        float v[32];
        if (nkb > 0) {
            tc::tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + c0, v);
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(tr + lane * 32 + 4 * (j ^ (lane & 7))) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int j = lane & 7;
        const int64_t col = n0 + c0 + 4 * j;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int R = it * 4 + (lane >> 3);
            const int64_t row = m0 + lg * 32 + R;
            float4 x = *reinterpret_cast<const float4*>(tr + R * 32 + 4 * (j ^ (R & 7)));
            if (row < g.M && col < g.N) {
                if (partial) {
                    *reinterpret_cast<float4*>(outp + row * g.N + col) = x;
                } else {
                    if (g.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col));
                        x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
                    }
                    if (g.residual) {
                        const float4 r = __ldg(reinterpret_cast<const float4*>(g.residual + row * g.ldr + col));
                        x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
                    }
                    if (g.c_bf16) {
                        uint2 o;
                        o.x = tc::pack_bf16(x.x, x.y); o.y = tc::pack_bf16(x.z, x.w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(g.C) + row * g.ldc + col) = o;
                    } else {
                        float* cp = reinterpret_cast<float*>(g.C) + row * g.ldc + col;
                        if (g.accumulate) {
                            const float4 c = *reinterpret_cast<const float4*>(cp);
                            x.x += c.x; x.y += c.y; x.z += c.z; x.w += c.w;
                        }
                        *reinterpret_cast<float4*>(cp) = x;
                    }
                }
            }
        }
        __syncwarp();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// gemm2: persistent, warp-specialised variant for the two products whose A operand is a K-major bf16 activation
// (forward y = x W^T and d-input dx = dy W).  The one-tile-per-CTA kernel above spends most of its time in its
// prologue, its per-K-block CTA barrier and an epilogue that nothing overlaps; here a CTA walks a strided list of
// 128x128 output tiles and the roles never meet at a CTA barrier:
//   loader warps (6, two per ring stage): LDG.128 -> registers -> STS.128 into the stage's padded chunk-major slabs,
//       generic->async proxy fence, arrive(full[stage]);  refill after empty[stage]
//   MMA warp (1 lane): wait full -> 4 x tcgen05.mma (K16) -> commit(empty[stage]); the last K block of a tile also
//       commits acc_full[buf]; the accumulator is DOUBLE BUFFERED in TMEM (2 x 128 columns), so the next tile's
//       products start while the epilogue drains this one
//   epilogue warps (4): wait acc_full[buf] -> tcgen05.ld (32 columns at a time) -> bias / residual -> fp32 or bf16
//       row segments straight to global -> arrive(acc_empty[buf])
// 2 CTAs/SM (3 x 32.5 KB ring + 256 TMEM columns each), 12 warps per CTA.
namespace dn2 {
constexpr int STAGES = 3, NLOAD = 2 * STAGES, THREADS = 384;
constexpr int W_MMA = NLOAD, W_EPI = 8;             // warps 0..5 loaders (stage w % 3, half w / 3), 6 MMA, 7 idle, 8..11 epilogue
constexpr uint32_t SMEM_BYTES = STAGES * dn::STAGE_BYTES;
}

// half of a stage's operand tile: global (row-major bf16) -> registers (8 x LDG.128 in flight) -> chunk-major shared slab.
// cp.async (LDGSTS) is NOT used here: with 16-byte pieces scattered over the slabs it sustains one warp instruction per
// ~150 cycles (measured), an order of magnitude below plain LDG + STS.
template <bool MN>
__device__ __forceinline__ void stage_half(const bf16* __restrict__ src, int64_t ld, int64_t row0, int64_t col0,
                                           int64_t row_lim, int64_t col_lim, uint8_t* tile, int half, int lane) {
    constexpr int CH = MN ? 16 : 8;                 // 16-byte chunks per tile row
    constexpr int ROWS_HALF = MN ? 32 : 64;
    constexpr uint32_t SLAB = MN ? dn::SLAB_MN : dn::SLAB_K;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = (b * 8 + j) * 32 + lane;
            const int r = half * ROWS_HALF + i / CH, ch = i % CH;
            const int64_t row = row0 + r, col = col0 + ch * 8;
            v[j] = (row < row_lim && col < col_lim) ? __ldg(reinterpret_cast<const uint4*>(src + row * ld + col)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = (b * 8 + j) * 32 + lane;
            const int r = half * ROWS_HALF + i / CH, ch = i % CH;
            *reinterpret_cast<uint4*>(tile + ch * SLAB + r * 16) = v[j];
        }
    }
}

template <bool B_MN>
__global__ void __launch_bounds__(dn2::THREADS, 2)
gemm2_kernel(const GemmArgs g, int tiles_n, int total_tiles) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t full[dn2::STAGES], empty[dn2::STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (int)((g.K + dn::BK - 1) / dn::BK);
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < dn2::STAGES; ++s) { tc::mbar_init(&full[s], 2); tc::mbar_init(&empty[s], 1); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_empty[b], 4); }
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sm_base = tc::smem_u32(sm);
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < dn2::NLOAD) {
        // ------------------------------------------------------------------ loader: ring stage warp % 3, tile half warp / 3
        const int stg = warp % dn2::STAGES, half = warp / dn2::STAGES;
        uint8_t* st = sm + stg * dn::STAGE_BYTES;
        const int total_kb = my_tiles * nkb;          // K blocks are numbered across the CTA's tiles: block G lives in stage G % STAGES
        for (int G = stg, it = 0; G < total_kb; G += dn2::STAGES, ++it) {
            const int t = G / nkb, kb = G - t * nkb;
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int64_t m0 = (int64_t)(tile / tiles_n) * dn::BM, n0 = (int64_t)(tile % tiles_n) * dn::BN;
            if (it > 0) tc::mbar_wait(&empty[stg], (uint32_t)((it - 1) & 1));           // the MMAs that read the stage are done
            const int64_t k0 = (int64_t)kb * dn::BK;
            stage_half<false>((const bf16*)g.A, g.lda, m0, k0, g.M, g.K, st, half, lane);
            if (B_MN) stage_half<true>((const bf16*)g.B, g.ldb, k0, n0, g.K, g.N, st + dn::OP_BYTES, half, lane);
            else      stage_half<false>((const bf16*)g.B, g.ldb, n0, k0, g.N, g.K, st + dn::OP_BYTES, half, lane);
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[stg]);
        }
    } else if (warp == dn2::W_MMA) {
        // ------------------------------------------------------------------ tensor-core issue (one lane)
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(dn::BM, dn::BN, 0, B_MN ? 1 : 0);
            constexpr uint32_t A_KSTEP = 2u * dn::SLAB_K;
            constexpr uint32_t B_LBO = B_MN ? 128u : dn::SLAB_K, B_SBO = B_MN ? dn::SLAB_MN : 128u, B_KSTEP = B_MN ? 256u : 2u * dn::SLAB_K;
            int G = 0;                                // running K-block number (see the loader)
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t & 1;
                if (t >= 2) { tc::mbar_wait(&acc_empty[buf], (uint32_t)(((t - 2) >> 1) & 1)); tc::fence_after_sync(); }
                for (int kb = 0; kb < nkb; ++kb, ++G) {
                    const int s = G % dn2::STAGES;
                    tc::mbar_wait(&full[s], (uint32_t)((G / dn2::STAGES) & 1));
                    tc::fence_after_sync();
                    const tc::Desc dA = tc::make_desc2(sm_base + s * dn::STAGE_BYTES, dn::SLAB_K, 128u);
                    const tc::Desc dB = tc::make_desc2(sm_base + s * dn::STAGE_BYTES + dn::OP_BYTES, B_LBO, B_SBO);
#pragma unroll
                    for (int ks = 0; ks < dn::BK / 16; ++ks)
                        tc::mma_bf16(tmem + (uint32_t)buf * 128u, dA.adv(ks * A_KSTEP).u64(), dB.adv(ks * B_KSTEP).u64(), idesc, (kb | ks) != 0);
                    tc::mma_commit(&empty[s]);
                }
                tc::mma_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp >= dn2::W_EPI) {
        // ------------------------------------------------------------------ epilogue: lane quarter = warp & 3
        const int lg = warp & 3;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x, buf = t & 1;
            const int64_t m0 = (int64_t)(tile / tiles_n) * dn::BM, n0 = (int64_t)(tile % tiles_n) * dn::BN;
            const int64_t row = m0 + lg * 32 + lane;
            tc::mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));
            tc::fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < dn::BN; c0 += 32) {
                float v[32];
                tc::tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)buf * 128u + c0, v);
                const int64_t col = n0 + c0;
                if (row < g.M && col < g.N) {          // N is a multiple of 32 on this path (checked by the host)
                    if (g.bias) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + c));
                            v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
                        }
                    }
                    if (g.residual) {
                        const float* rp = g.residual + row * g.ldr + col;
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 r = __ldg(reinterpret_cast<const float4*>(rp + c));
                            v[c] += r.x; v[c + 1] += r.y; v[c + 2] += r.z; v[c + 3] += r.w;
                        }
                    }
                    if (g.c_bf16) {
                        bf16* cp = reinterpret_cast<bf16*>(g.C) + row * g.ldc + col;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            uint4 o;
                            o.x = tc::pack_bf16(v[c], v[c + 1]); o.y = tc::pack_bf16(v[c + 2], v[c + 3]);
                            o.z = tc::pack_bf16(v[c + 4], v[c + 5]); o.w = tc::pack_bf16(v[c + 6], v[c + 7]);
                            *reinterpret_cast<uint4*>(cp + c) = o;
                        }
                    } else {
                        float* cp = reinterpret_cast<float*>(g.C) + row * g.ldc + col;
#pragma unroll
                        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(cp + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

template <bool B_MN>
static int launch_gemm2(const GemmArgs& g, cudaStream_t st) {
    static bool attr_done_dev[64] = {};      // the attribute is per device: one flag per device ordinal
    int dev_ = 0; cudaGetDevice(&dev_);
    bool& attr_done = attr_done_dev[dev_ & 63];
    auto kern = gemm2_kernel<B_MN>;
    if (!attr_done) {
        GAOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dn2::SMEM_BYTES));
        attr_done = true;
    }
    const int tiles_n = (int)((g.N + dn::BN - 1) / dn::BN), tiles_m = (int)((g.M + dn::BM - 1) / dn::BM);
    const int total = tiles_n * tiles_m;
    const int grid = std::min(total, 2 * kNumSMs);
    kern<<<grid, dn2::THREADS, dn2::SMEM_BYTES, st>>>(g, tiles_n, total);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

// =====================================================================================================
// gemm3: the same warp-specialised persistent GEMM with the operand tiles moved by the TMA engine.
// What ncu said about gemm2 (profiles/r01f_gemm2_*): 47 % of the stall samples are the loader warps waiting
// on LDG (each loader keeps 8 x 16 B in flight: ~30 KB per SM against the ~120 KB that L2 latency x bandwidth
// asks for), and most launches have <= 1 tile per CTA, so nothing hides that latency.  Here ONE lane issues
// cp.async.bulk.tensor for a whole [128 x 64] bf16 box per operand and K block (32 KB per stage in flight per
// request pair, no registers, no generic->async proxy fence); the boxes land in the 128-byte-swizzled
// canonical layouts that tcgen05.mma reads directly:
//   K-major operand  (rows = M/N, 64 contraction elements = 128 B per row): SBO = 1024 B (8-row groups),
//                    K16 step = +32 B inside the swizzle atom
//   MN-major operand (rows = contraction, 64 M/N elements = 128 B per row; two boxes for 128 columns):
//                    LBO = 8192 B (next 64-wide M/N block), SBO = 1024 B (next 8 contraction rows),
//                    K16 step = +2048 B
// (cute::UMMA canonical layouts, mma_traits_sm100.hpp make_umma_desc; layout_type SWIZZLE_128B = 2.)
// 3-stage ring, 2 CTAs/SM, accumulator double-buffered in TMEM, epilogue as in gemm2.
// =====================================================================================================
namespace dn3 {
constexpr int STAGES = 3, THREADS = 192;          // warp 0: TMA producer lane, warp 1: MMA lane, warps 2..5: epilogue
constexpr uint32_t OP_BYTES = 128 * 64 * 2;       // one operand tile (16 KB)
constexpr uint32_t STAGE_BYTES = 2 * OP_BYTES;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;     // + slack for the 1024-byte alignment of the ring
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled tensor_map_encoder() {
    static PFN_encodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (PFN_encodeTiled)p;
    }();
    return fn;
}

// row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = 64 contiguous elements x box_rows rows
static bool make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    PFN_encodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tm, int c_inner, int c_outer, uint64_t* mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_dst), "l"(tm), "r"(c_inner), "r"(c_outer), "r"(tc::smem_u32(mbar)) : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2 in bits 61..63), version 1
__device__ __forceinline__ tc::Desc make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return tc::Desc{((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16),
                    ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29)};
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(dn3::THREADS, 2)
gemm3_kernel(const GemmArgs g, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int tiles_n, int tiles_mn,
             int total_tiles) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    __shared__ uint64_t full[dn3::STAGES], empty[dn3::STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work item = (output tile, K split): item / tiles_mn is the split, whose K blocks are [z * kb_per_split, ...)
    const int nkb_all = (int)((g.K + dn::BK - 1) / dn::BK);
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < dn3::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_empty[b], 4); }
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sm_base = (tc::smem_u32(sm_raw) + 1023u) & ~1023u;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one lane)
        if (tc::elect_one()) {
            int G = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int item = (int)blockIdx.x + t * (int)gridDim.x;
                const int z = item / tiles_mn, tile = item - z * tiles_mn;
                const int m0 = (tile / tiles_n) * dn::BM, n0 = (tile % tiles_n) * dn::BN;
                const int kb0 = z * g.kb_per_split, kb1 = min(nkb_all, kb0 + g.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++G) {
                    const int s = G % dn3::STAGES;
                    if (G >= dn3::STAGES) tc::mbar_wait(&empty[s], (uint32_t)(((G / dn3::STAGES) - 1) & 1));
                    const uint32_t sa = sm_base + s * dn3::STAGE_BYTES, sb = sa + dn3::OP_BYTES;
                    tc::mbar_arrive_expect_tx(&full[s], dn3::STAGE_BYTES);
                    const int k0 = kb * dn::BK;
                    if (A_MN) {
                        tma_load_2d(sa, &tmA, m0, k0, &full[s]);                     // [64 contraction rows x 64 columns] x 2
                        tma_load_2d(sa + dn3::OP_BYTES / 2, &tmA, m0 + 64, k0, &full[s]);
                    } else {
                        tma_load_2d(sa, &tmA, k0, m0, &full[s]);
                    }
                    if (B_MN) {
                        tma_load_2d(sb, &tmB, n0, k0, &full[s]);                     // [64 contraction rows x 64 columns] x 2
                        tma_load_2d(sb + dn3::OP_BYTES / 2, &tmB, n0 + 64, k0, &full[s]);
                    } else {
                        tma_load_2d(sb, &tmB, k0, n0, &full[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ tensor-core issue (one lane)
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(dn::BM, dn::BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            constexpr uint32_t A_LBO = A_MN ? dn3::OP_BYTES / 2 : 16u, A_KSTEP = A_MN ? 2048u : 32u;
            constexpr uint32_t B_LBO = B_MN ? dn3::OP_BYTES / 2 : 16u, B_KSTEP = B_MN ? 2048u : 32u;
            int G = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t & 1;
                const int z = ((int)blockIdx.x + t * (int)gridDim.x) / tiles_mn;
                const int kb0 = z * g.kb_per_split, nkb = min(nkb_all, kb0 + g.kb_per_split) - kb0;
                if (t >= 2) { tc::mbar_wait(&acc_empty[buf], (uint32_t)(((t - 2) >> 1) & 1)); tc::fence_after_sync(); }
                for (int kb = 0; kb < nkb; ++kb, ++G) {
                    const int s = G % dn3::STAGES;
                    tc::mbar_wait(&full[s], (uint32_t)((G / dn3::STAGES) & 1));
                    tc::fence_after_sync();
                    const tc::Desc dA = make_desc_sw128(sm_base + s * dn3::STAGE_BYTES, A_LBO, 1024u);
                    const tc::Desc dB = make_desc_sw128(sm_base + s * dn3::STAGE_BYTES + dn3::OP_BYTES, B_LBO, 1024u);
#pragma unroll
                    for (int ks = 0; ks < dn::BK / 16; ++ks)
                        tc::mma_bf16(tmem + (uint32_t)buf * 128u, dA.adv(ks * A_KSTEP).u64(), dB.adv(ks * B_KSTEP).u64(), idesc, (kb | ks) != 0);
                    tc::mma_commit(&empty[s]);
                }
                tc::mma_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM lane quarter = warp & 3
        const int lg = warp & 3;
        for (int t = 0; t < my_tiles; ++t) {
            const int item = (int)blockIdx.x + t * (int)gridDim.x, buf = t & 1;
            const int z = item / tiles_mn, tile = item - z * tiles_mn;
            const int64_t m0 = (int64_t)(tile / tiles_n) * dn::BM, n0 = (int64_t)(tile % tiles_n) * dn::BN;
            const int64_t row = m0 + lg * 32 + lane;
            // the residual rows of this tile are requested BEFORE the wait on the accumulator (two 32-column groups in
            // flight, refilled while the previous group is written): their latency hides behind the main loop
            float4 rr[2][8];
            const float* rrow = g.residual ? g.residual + row * g.ldr + n0 : nullptr;
            if (rrow && row < g.M) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int c = 0; c < 8; ++c) rr[q][c] = __ldg(reinterpret_cast<const float4*>(rrow + q * 32) + c);
            }
            tc::mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));
            tc::fence_after_sync();
            if (g.partial) {
                // split-K: fp32 partial tile of split z (fixed-order reduction afterwards)
                float* pp = g.partial + (size_t)z * g.M * g.N + row * g.N + n0;
#pragma unroll 1
                for (int c0 = 0; c0 < dn::BN; c0 += 32) {
                    float v[32];
                    tc::tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)buf * 128u + c0, v);
                    if (row < g.M && n0 + c0 < g.N) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(pp + c0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
                continue;
            }
#pragma unroll
            for (int c0 = 0; c0 < dn::BN; c0 += 32) {
                float v[32];
                tc::tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)buf * 128u + c0, v);
                const int64_t col = n0 + c0;
                if (row < g.M && col < g.N) {          // N is a multiple of 128 on this path (checked by the host)
                    if (g.bias) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + c));
                            v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
                        }
                    }
                    if (rrow) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 r = rr[(c0 / 32) & 1][c];
                            v[4 * c] += r.x; v[4 * c + 1] += r.y; v[4 * c + 2] += r.z; v[4 * c + 3] += r.w;
                        }
                        if (c0 + 64 < dn::BN) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) rr[(c0 / 32) & 1][c] = __ldg(reinterpret_cast<const float4*>(rrow + c0 + 64) + c);
                        }
                    }
                    if (g.c_bf16) {
                        bf16* cp = reinterpret_cast<bf16*>(g.C) + row * g.ldc + col;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            uint4 o;
                            o.x = tc::pack_bf16(v[c], v[c + 1]); o.y = tc::pack_bf16(v[c + 2], v[c + 3]);
                            o.z = tc::pack_bf16(v[c + 4], v[c + 5]); o.w = tc::pack_bf16(v[c + 6], v[c + 7]);
                            *reinterpret_cast<uint4*>(cp + c) = o;
                        }
                    } else {
                        float* cp = reinterpret_cast<float*>(g.C) + row * g.ldc + col;
#pragma unroll
                        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(cp + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// returns GAOT_ERR_UNSUPPORTED when the operands do not meet the TMA constraints (caller falls back to the register-staged kernels)
template <bool A_MN, bool B_MN>
static int launch_gemm3(const GemmArgs& g, int splits, cudaStream_t st) {
    static bool attr_done_dev[64] = {};      // the attribute is per device: one flag per device ordinal
    int dev_ = 0; cudaGetDevice(&dev_);
    bool& attr_done = attr_done_dev[dev_ & 63];
    auto kern = gemm3_kernel<A_MN, B_MN>;
    if (((uintptr_t)g.A | (uintptr_t)g.B) & 15) return GAOT_ERR_UNSUPPORTED;
    CUtensorMap tmA, tmB;
    // K-major operand -> matrix [M or N, K] (box 128 rows x 64 contraction elements);
    // MN-major operand -> matrix [K, M or N] (box 64 contraction rows x 64 columns, two boxes per tile)
    if (!(A_MN ? make_tmap(&tmA, g.A, g.K, g.M, g.lda, 64) : make_tmap(&tmA, g.A, g.M, g.K, g.lda, 128))) return GAOT_ERR_UNSUPPORTED;
    if (!(B_MN ? make_tmap(&tmB, g.B, g.K, g.N, g.ldb, 64) : make_tmap(&tmB, g.B, g.N, g.K, g.ldb, 128))) return GAOT_ERR_UNSUPPORTED;
    if (!attr_done) {
        GAOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dn3::SMEM_BYTES));
        attr_done = true;
    }
    const int tiles_n = (int)((g.N + dn::BN - 1) / dn::BN), tiles_m = (int)((g.M + dn::BM - 1) / dn::BM);
    const int tiles_mn = tiles_n * tiles_m, total = tiles_mn * splits;
    const int grid = std::min(total, 2 * kNumSMs);
    kern<<<grid, dn3::THREADS, dn3::SMEM_BYTES, st>>>(g, tmA, tmB, tiles_n, tiles_mn, total);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

// fixed-order sum of the split-K partials (+ bias / accumulate): deterministic weight gradients.
// Block = 32 float4 columns x 8 split groups: group g sums splits g, g + 8, ... (four loads in flight), the eight group sums
// are added in fixed order through shared memory.  (Round 1 had one thread walk all <= 64 partials of a float4 serially:
// 11.5 us per call, 0.59 ms per step over 51 calls.)
__global__ void __launch_bounds__(256)
gemm_splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t M, int64_t N, const float* __restrict__ bias,
                          float* __restrict__ C, int64_t ldc, int accumulate) {
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t i4 = (int64_t)blockIdx.x * 32 + lane;
    const int64_t total4 = M * N / 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < total4) {
        const float4* p = reinterpret_cast<const float4*>(part) + i4;
        const size_t zs = (size_t)(M * N / 4);
        int z = g;
        for (; z + 24 < splits; z += 32) {
            const float4 a = __ldg(p + (size_t)z * zs), b = __ldg(p + (size_t)(z + 8) * zs);
            const float4 c = __ldg(p + (size_t)(z + 16) * zs), d = __ldg(p + (size_t)(z + 24) * zs);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
        }
        for (; z < splits; z += 8) {
            const float4 a = __ldg(p + (size_t)z * zs);
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
    }
    red[g][lane] = acc;
    __syncthreads();
    if (g != 0 || i4 >= total4) return;
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 r = red[k][lane]; acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w; }
    const int64_t row = (i4 * 4) / N, col = (i4 * 4) % N;
    if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col));
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    float* cp = C + row * ldc + col;
    if (accumulate) {
        const float4 c = *reinterpret_cast<const float4*>(cp);
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
    }
    *reinterpret_cast<float4*>(cp) = acc;
}

static int pick_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ((M + dn::BM - 1) / dn::BM) * ((N + dn::BN - 1) / dn::BN);
    const int64_t nkb = (K + dn::BK - 1) / dn::BK;
    if (tiles >= kNumSMs || nkb < 8) return 1;
    int64_t s = (2 * kNumSMs + tiles - 1) / tiles;
    s = std::min<int64_t>(s, nkb / 4);
    return (int)std::max<int64_t>(s, 1);
}

template <typename TA, typename TB, bool A_MN, bool B_MN>
static int launch_gemm(const GemmArgs& g, int splits, cudaStream_t st) {
    static bool attr_done_dev[64] = {};      // the attribute is per device: one flag per device ordinal
    int dev_ = 0; cudaGetDevice(&dev_);
    bool& attr_done = attr_done_dev[dev_ & 63];
    auto kern = gemm_tc_kernel<TA, TB, A_MN, B_MN>;
    if (!attr_done) {
        GAOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dn::SMEM_BYTES));
        attr_done = true;
    }
    dim3 grid((unsigned)((g.N + dn::BN - 1) / dn::BN), (unsigned)((g.M + dn::BM - 1) / dn::BM), (unsigned)splits);
    kern<<<grid, dn::THREADS, dn::SMEM_BYTES, st>>>(g);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

// dtype codes of the C ABI: 0 = float32, 1 = bfloat16
static int run_gemm(GemmArgs g, int a_dtype, int b_dtype, bool a_mn, bool b_mn, void* ws, size_t ws_bytes, cudaStream_t st,
                    const char* timer_name) {
    GAOT_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem");
    GAOT_CHECK_ARG(g.N % 4 == 0 && g.ldc % 4 == 0, "gemm: N and ldc must be multiples of 4");
    const int64_t a_al = a_dtype ? 8 : 4, b_al = b_dtype ? 8 : 4;
    if (g.lda % a_al || g.ldb % b_al || (g.A2 && g.lda2 % a_al) || (a_mn ? g.M : g.K) % a_al || (b_mn ? g.N : g.K) % b_al ||
        (g.A2 && (g.k_split % dn::BK || a_mn))) {
        set_error("gemm: operand rows must be 16-byte aligned (contiguous extents multiples of %d/%d elements)", (int)a_al, (int)b_al);
        return GAOT_ERR_UNSUPPORTED;
    }
    if (!g.A2) g.k_split = g.K;
    const int nkb = (int)((g.K + dn::BK - 1) / dn::BK);
    int splits = (g.c_bf16 || g.residual) ? 1 : pick_splits(g.M, g.N, g.K);
    g.kb_per_split = (nkb + splits - 1) / splits;
    splits = (nkb + g.kb_per_split - 1) / g.kb_per_split;
    g.partial = nullptr;
    if (splits > 1) {
        const size_t need = (size_t)splits * g.M * g.N * sizeof(float);
        if (ws_bytes < need || !ws) { set_error("gemm: workspace too small (%zu < %zu)", ws_bytes, need); return GAOT_ERR_WORKSPACE; }
        g.partial = (float*)ws;
    }
    GAOT_TIME_KERNEL(timer_name, st, 2.0 * (double)g.M * (double)g.N * (double)g.K);
    int rc;
    {   // persistent warp-specialised kernel for bf16 K-major-A products (GAOT_GEMM_OLD=1 selects the one-tile kernel)
        static const bool use_old = getenv("GAOT_GEMM_OLD") && atoi(getenv("GAOT_GEMM_OLD")) != 0;
        // GAOT_GEMM_GEN=2 pins the register-staged kernels (A/B timing); default: the TMA-fed gemm3 when the operands allow it
        static const bool no_tma = getenv("GAOT_GEMM_GEN") && atoi(getenv("GAOT_GEMM_GEN")) == 2;
        const bool tma_ok = !no_tma && !use_old && a_dtype && b_dtype && !g.A2 && !g.accumulate && g.K % dn::BK == 0 &&
                            g.N % dn::BN == 0 && g.M >= dn::BM && (!a_mn || g.M % dn::BM == 0);
        if (tma_ok && (!a_mn || b_mn)) {
            rc = GAOT_ERR_UNSUPPORTED;
            if (!a_mn && !b_mn && splits == 1) rc = launch_gemm3<false, false>(g, 1, st);
            else if (!a_mn && b_mn && splits == 1) rc = launch_gemm3<false, true>(g, 1, st);
            else if (a_mn && b_mn) rc = launch_gemm3<true, true>(g, splits, st);
            if (rc == GAOT_OK && splits > 1) {
                const int64_t total4 = g.M * g.N / 4;
                gemm_splitk_reduce_kernel<<<(unsigned)((total4 + 31) / 32), 256, 0, st>>>(g.partial, splits, g.M, g.N, g.bias, (float*)g.C, g.ldc, g.accumulate);
                GAOT_LAUNCH_CHECK();
            }
            if (rc != GAOT_ERR_UNSUPPORTED) return rc;
        }
        if (!use_old && a_dtype && b_dtype && !a_mn && !g.A2 && !g.accumulate && splits == 1 && g.N % 32 == 0)
            return b_mn ? launch_gemm2<true>(g, st) : launch_gemm2<false>(g, st);
    }
    const int sel = (a_dtype ? 8 : 0) | (b_dtype ? 4 : 0) | (a_mn ? 2 : 0) | (b_mn ? 1 : 0);
    switch (sel) {
        case 0:  rc = launch_gemm<float, float, false, false>(g, splits, st); break;
        case 1:  rc = launch_gemm<float, float, false, true>(g, splits, st); break;
        case 3:  rc = launch_gemm<float, float, true, true>(g, splits, st); break;
        case 4:  rc = launch_gemm<float, bf16, false, false>(g, splits, st); break;
        case 5:  rc = launch_gemm<float, bf16, false, true>(g, splits, st); break;
        case 12: rc = launch_gemm<bf16, bf16, false, false>(g, splits, st); break;
        case 13: rc = launch_gemm<bf16, bf16, false, true>(g, splits, st); break;
        case 15: rc = launch_gemm<bf16, bf16, true, true>(g, splits, st); break;
        case 11: rc = launch_gemm<bf16, float, true, true>(g, splits, st); break;
        case 7:  rc = launch_gemm<float, bf16, true, true>(g, splits, st); break;
        default: set_error("gemm: operand combination %d not instantiated", sel); return GAOT_ERR_UNSUPPORTED;
    }
    if (rc != GAOT_OK) return rc;
    if (splits > 1) {
        const int64_t total4 = g.M * g.N / 4;
        gemm_splitk_reduce_kernel<<<(unsigned)((total4 + 31) / 32), 256, 0, st>>>(g.partial, splits, g.M, g.N, g.bias, (float*)g.C, g.ldc, g.accumulate);
        GAOT_LAUNCH_CHECK();
    }
    return GAOT_OK;
}

}  // namespace gaot

using namespace gaot;

extern "C" {

size_t gaot_linear_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    // upper bound over the three products of one layer (split-K partials of the largest split count)
    auto need = [](int64_t m, int64_t n, int64_t k) { return (size_t)pick_splits(m, n, k) * m * n * sizeof(float); };
    size_t b = std::max(need(M, N, K), std::max(need(M, K, N), need(N, K, M)));
    return align_up(b + 256);
}

int gaot_linear_forward(const void* x, int x_dtype, int64_t ldx, const void* x2, int64_t ldx2, int64_t k_split,
                        const void* w, int w_dtype, int64_t M, int64_t N, int64_t K,
                        const float* bias, const float* residual, int64_t ldr,
                        void* y, int y_dtype, int64_t ldy, void* ws, size_t ws_bytes, void* stream) {
    GAOT_CHECK_ARG(x && w && y, "linear_forward: null pointer");
    GemmArgs g{};
    g.A = x; g.lda = ldx; g.A2 = x2; g.lda2 = ldx2; g.k_split = x2 ? k_split : K;
    g.B = w; g.ldb = K; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.residual = residual; g.ldr = ldr; g.C = y; g.ldc = ldy; g.c_bf16 = y_dtype; g.accumulate = 0;
    return run_gemm(g, x_dtype, w_dtype, false, false, ws, ws_bytes, (cudaStream_t)stream, "linear_fwd");
}

int gaot_linear_backward_input(const void* dy, int dy_dtype, int64_t lddy, const void* w, int w_dtype,
                               int64_t M, int64_t N, int64_t K, const float* residual, int64_t ldr,
                               void* dx, int dx_dtype, int64_t lddx, int accumulate,
                               void* ws, size_t ws_bytes, void* stream) {
    GAOT_CHECK_ARG(dy && w && dx, "linear_backward_input: null pointer");
    GemmArgs g{};
    g.A = dy; g.lda = lddy; g.B = w; g.ldb = K;
    g.M = M; g.N = K; g.K = N;                                    // dx[M,K] = dy[M,N] * W[N,K], contraction over N
    g.residual = residual; g.ldr = ldr; g.C = dx; g.ldc = lddx; g.c_bf16 = dx_dtype; g.accumulate = accumulate;
    return run_gemm(g, dy_dtype, w_dtype, false, true, ws, ws_bytes, (cudaStream_t)stream, "linear_bwd_x");
}

int gaot_linear_backward_weight(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                                int64_t M, int64_t N, int64_t K, float* dw, int64_t lddw, int accumulate,
                                void* ws, size_t ws_bytes, void* stream) {
    GAOT_CHECK_ARG(dy && x && dw, "linear_backward_weight: null pointer");
    GemmArgs g{};
    g.A = dy; g.lda = lddy; g.B = x; g.ldb = ldx;
    g.M = N; g.N = K; g.K = M;                                    // dW[N,K] = dy^T[N,M] * x[M,K], contraction over M
    g.C = dw; g.ldc = lddw; g.c_bf16 = 0; g.accumulate = accumulate;
    return run_gemm(g, dy_dtype, x_dtype, true, true, ws, ws_bytes, (cudaStream_t)stream, "linear_bwd_w");
}

}  // extern "C"
