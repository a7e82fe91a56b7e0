// attn.cu -- latent-token attention: flash-style forward and backward on tcgen05 tensor cores
// with TMEM accumulators (BF16 operands, FP32 accumulation / softmax statistics).
// Replaces rotary_emb + F.scaled_dot_product_attention of reference src/model/layers/attn.py:110-128
// (fp32 SDPA -> mem-efficient/math backends; no Blackwell tensor path reachable with fp32 inputs).
//
// Forward  : CTA = 128 query rows of one (batch, head); loop over 128-key tiles:
//            S = Q K^T (tcgen05.mma, M128 N128 K16 x d/16) -> TMEM; each of the 128 threads owns one
//            row (tcgen05.ld 32x32b): online softmax in registers in the log2 domain, P -> bf16
//            chunk-major shared tile; O_tile = P V (V consumed as an MN-major operand straight
//            from its natural [key][d] layout) -> TMEM -> registers, O = O*alpha + O_tile.
//            2 CTAs/SM overlap one CTA's MMAs with the other's exponentials (the kernel is
//            MUFU-bound at d = 32: S^2*h exps vs 4*S^2*H flops).
// Backward : CTA = 128 keys of one (batch, head); loop over 128-query tiles.  S^T = K Q^T and
//            dP^T = V dO^T land in TMEM; threads (one key row each, two column halves) rebuild
//            P^T = exp2(S^T*c - lse) and dS^T = P^T (dP^T - D) * scale, store both bf16 chunk-major;
//            dV += P^T dO and dK += dS^T Q accumulate in TMEM across the whole query loop,
//            dQ_tile = dS K (the dS^T tile re-read as an MN-major A operand) is added to global dQ
//            with vector reductions.
#include "common.cuh"
#include "tc05.cuh"

namespace gaot {

using bf16 = __nv_bfloat16;

// --------------------------------------------------------------------------- prep kernels
// fp32 [B,S,nh*d] (token-major projection output) -> bf16 [B,nh,S,d], optional RoPE.
template <bool ROPE>
__device__ __forceinline__ void rope8(float (&x)[8], int s, int c0, const float* __restrict__ freqs, bool inverse) {
    if (!ROPE) return;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float ang = (float)s * freqs[(c0 >> 1) + p];
        float sn, cs;
        sincosf(ang, &sn, &cs);
        if (inverse) sn = -sn;
        const float a = x[2 * p], b = x[2 * p + 1];
        x[2 * p] = a * cs - b * sn;
        x[2 * p + 1] = b * cs + a * sn;
    }
}

__global__ void __launch_bounds__(256)
attn_prep_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t B, int64_t S, int nh, int d,
                 const float* __restrict__ freqs /* NULL: no rope */) {
    const int cpr = d >> 3;                                   // 8-element chunks per head row
    const int64_t total = B * S * nh * cpr;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % cpr);
    const int h = (int)((idx / cpr) % nh);
    const int64_t s = (idx / ((int64_t)cpr * nh)) % S;
    const int64_t b = idx / ((int64_t)cpr * nh * S);
    const float* p = src + ((b * S + s) * nh + h) * d + ch * 8;
    const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
    float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (freqs) rope8<true>(x, (int)s, ch * 8, freqs, false);
    uint4 o;
    o.x = tc::pack_bf16(x[0], x[1]); o.y = tc::pack_bf16(x[2], x[3]);
    o.z = tc::pack_bf16(x[4], x[5]); o.w = tc::pack_bf16(x[6], x[7]);
    *reinterpret_cast<uint4*>(dst + ((b * nh + h) * S + s) * d + ch * 8) = o;
}

// backward prep: dO -> bf16 [B,H,S,d] and Dvec[b,h,s] = sum_c dO*O
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const float* __restrict__ dO, const float* __restrict__ O, bf16* __restrict__ dOb,
                     float* __restrict__ Dvec, int64_t B, int64_t S, int H, int d) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (b, s, h)
    if (idx >= B * S * H) return;
    const int h = (int)(idx % H);
    const int64_t s = (idx / H) % S, b = idx / ((int64_t)H * S);
    const float* pd = dO + idx * d;
    const float* po = O + idx * d;
    bf16* out = dOb + ((b * H + h) * S + s) * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(pd + c), a1 = *reinterpret_cast<const float4*>(pd + c + 4);
        const float4 o0 = *reinterpret_cast<const float4*>(po + c), o1 = *reinterpret_cast<const float4*>(po + c + 4);
        acc += a0.x * o0.x + a0.y * o0.y + a0.z * o0.z + a0.w * o0.w + a1.x * o1.x + a1.y * o1.y + a1.z * o1.z + a1.w * o1.w;
        uint4 o;
        o.x = tc::pack_bf16(a0.x, a0.y); o.y = tc::pack_bf16(a0.z, a0.w);
        o.z = tc::pack_bf16(a1.x, a1.y); o.w = tc::pack_bf16(a1.z, a1.w);
        *reinterpret_cast<uint4*>(out + c) = o;
    }
    Dvec[(b * H + h) * S + s] = acc;
}

// backward post: [B,H,S,d] fp32 per-head grads -> token-major [B,S,nh_out*d], summing the GQA group
// and undoing RoPE (gradient of a rotation = rotation by the negative angle).
__global__ void __launch_bounds__(256)
attn_bwd_post_kernel(const float* __restrict__ gh, float* __restrict__ out, int64_t B, int64_t S, int H,
                     int nh_out, int d, const float* __restrict__ freqs) {
    const int cpr = d >> 3;
    const int64_t total = B * S * nh_out * cpr;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % cpr);
    const int ho = (int)((idx / cpr) % nh_out);
    const int64_t s = (idx / ((int64_t)cpr * nh_out)) % S;
    const int64_t b = idx / ((int64_t)cpr * nh_out * S);
    const int rep = H / nh_out;
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < rep; ++r) {
        const float* p = gh + ((b * H + ho * rep + r) * S + s) * d + ch * 8;
        const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
        x[0] += v0.x; x[1] += v0.y; x[2] += v0.z; x[3] += v0.w; x[4] += v1.x; x[5] += v1.y; x[6] += v1.z; x[7] += v1.w;
    }
    if (freqs) rope8<true>(x, (int)s, ch * 8, freqs, true);
    float* o = out + ((b * S + s) * nh_out + ho) * d + ch * 8;
    *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
}

// --------------------------------------------------------------------------- tile helpers
template <int D>
__device__ __forceinline__ void load_row(const bf16* __restrict__ g, bool valid, uint4 (&r)[D / 8]) {
#pragma unroll
    for (int c = 0; c < D / 8; ++c)
        r[c] = valid ? *reinterpret_cast<const uint4*>(g + c * 8) : make_uint4(0u, 0u, 0u, 0u);
}
template <int D>
__device__ __forceinline__ void store_row(uint8_t* tile, int row, const uint4 (&r)[D / 8]) {
#pragma unroll
    for (int c = 0; c < D / 8; ++c) *reinterpret_cast<uint4*>(tile + c * (128 * 16) + row * 16) = r[c];
}

// --------------------------------------------------------------------------- forward
template <int D>
__global__ void __launch_bounds__(128, 2)
attn_fwd_kernel(const bf16* __restrict__ Qb, const bf16* __restrict__ Kb, const bf16* __restrict__ Vb,
                float* __restrict__ out, float* __restrict__ lse, int S, int H, int Hkv, float scale_log2) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    constexpr int TILE_B = 128 * D * 2;
    uint8_t* Qs = sm;
    uint8_t* Ks = sm + TILE_B;            // [2]
    uint8_t* Vs = sm + 3 * TILE_B;        // [2]
    uint8_t* Ps = sm + 5 * TILE_B;        // 128 x 128 bf16
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int kvh = h / (H / Hkv);
    const int q = q0 + tid;
    const bool valid_q = q < S;
    const bf16* Kbase = Kb + ((size_t)(b * Hkv + kvh) * S) * D;
    const bf16* Vbase = Vb + ((size_t)(b * Hkv + kvh) * S) * D;
    const int nkv = (S + 127) / 128;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { tc::mbar_init(&mbar[0], 1); tc::mbar_init(&mbar[1], 1); tc::mbar_fence_init(); }
    {
        uint4 r[D / 8];
        load_row<D>(Qb + ((size_t)(b * H + h) * S + (valid_q ? q : 0)) * D, valid_q, r);
        store_row<D>(Qs, tid, r);
        const bool vk = tid < S;
        load_row<D>(Kbase + (size_t)(vk ? tid : 0) * D, vk, r);
        store_row<D>(Ks, tid, r);
        load_row<D>(Vbase + (size_t)(vk ? tid : 0) * D, vk, r);
        store_row<D>(Vs, tid, r);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t sQ = tc::smem_u32(Qs), sP = tc::smem_u32(Ps);
    constexpr uint32_t idescS = tc::make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescPV = tc::make_idesc_bf16(128, D, 0, 1);

    float m = -INFINITY, l = 0.f;
    float O[D];
#pragma unroll
    for (int c = 0; c < D; ++c) O[c] = 0.f;
    uint32_t ph0 = 0, ph1 = 0;

    for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        const uint32_t sK = tc::smem_u32(Ks + buf * TILE_B), sV = tc::smem_u32(Vs + buf * TILE_B);
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < D / 16; ++s)
                tc::mma_bf16(tmem, tc::desc_kmajor(sQ, 128, s), tc::desc_kmajor(sK, 128, s), idescS, s > 0);
            tc::mma_commit(&mbar[0]);
        }
        // prefetch the next K/V rows into registers while the tensor core works
        uint4 kreg[D / 8], vreg[D / 8];
        const bool have_next = j + 1 < nkv;
        if (have_next) {
            const int kn = (j + 1) * 128 + tid;
            const bool vk = kn < S;
            load_row<D>(Kbase + (size_t)(vk ? kn : 0) * D, vk, kreg);
            load_row<D>(Vbase + (size_t)(vk ? kn : 0) * D, vk, vreg);
        }
        tc::mbar_wait(&mbar[0], ph0); ph0 ^= 1;
        tc::fence_after_sync();

        float sv[128];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            float t[32];
            tc::tmem_ld32(tlane + c4 * 32, t);
#pragma unroll
            for (int c = 0; c < 32; ++c) sv[c4 * 32 + c] = t[c] * scale_log2;
        }
        const int kvalid = S - j * 128;            // keys >= kvalid in this tile are padding
        if (kvalid < 128) {
#pragma unroll
            for (int c = 0; c < 128; ++c) if (c >= kvalid) sv[c] = -INFINITY;
        }
        float mx = m;
#pragma unroll
        for (int c = 0; c < 128; ++c) mx = fmaxf(mx, sv[c]);
        const float alpha = exp2f(m - mx);          // m = -inf on the first tile -> 0
        float rs = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 16; ++c8) {
            float p[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { p[c] = exp2f(sv[c8 * 8 + c] - mx); rs += p[c]; }
            uint4 o;
            o.x = tc::pack_bf16(p[0], p[1]); o.y = tc::pack_bf16(p[2], p[3]);
            o.z = tc::pack_bf16(p[4], p[5]); o.w = tc::pack_bf16(p[6], p[7]);
            *reinterpret_cast<uint4*>(Ps + c8 * (128 * 16) + tid * 16) = o;
        }
        l = l * alpha + rs;
        m = mx;
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem, tc::desc_kmajor(sP, 128, s), tc::desc_mnmajor(sV, 128, s), idescPV, s > 0);
            tc::mma_commit(&mbar[1]);
        }
        if (have_next) {
            store_row<D>(Ks + (buf ^ 1) * TILE_B, tid, kreg);
            store_row<D>(Vs + (buf ^ 1) * TILE_B, tid, vreg);
        }
        tc::mbar_wait(&mbar[1], ph1); ph1 ^= 1;
        tc::fence_after_sync();
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
            float t[32];
            tc::tmem_ld32(tlane + c0, t);
#pragma unroll
            for (int c = 0; c < 32; ++c) O[c0 + c] = O[c0 + c] * alpha + t[c];
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    if (valid_q) {
        const float inv = 1.0f / l;
        float* o = out + ((size_t)b * S + q) * (H * D) + h * D;
#pragma unroll
        for (int c = 0; c < D; c += 4)
            *reinterpret_cast<float4*>(o + c) = make_float4(O[c] * inv, O[c + 1] * inv, O[c + 2] * inv, O[c + 3] * inv);
        lse[((size_t)b * H + h) * S + q] = m + log2f(l);
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// --------------------------------------------------------------------------- backward
template <int D>
__global__ void __launch_bounds__(256, 1)
attn_bwd_kernel(const bf16* __restrict__ Qb, const bf16* __restrict__ Kb, const bf16* __restrict__ Vb,
                const bf16* __restrict__ dOb, const float* __restrict__ lse, const float* __restrict__ Dvec,
                float* __restrict__ dQacc, float* __restrict__ dKh, float* __restrict__ dVh,
                int S, int H, int Hkv, float scale, float scale_log2) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float lse_s[128], D_s[128];
    constexpr int TILE_B = 128 * D * 2;
    uint8_t* Kt = sm;
    uint8_t* Vt = sm + TILE_B;
    uint8_t* Qs = sm + 2 * TILE_B;
    uint8_t* dOs = sm + 3 * TILE_B;
    uint8_t* PTs = sm + 4 * TILE_B;                 // P^T  [128 keys x 128 q] bf16
    uint8_t* dSs = PTs + 128 * 128 * 2;             // dS^T [128 keys x 128 q] bf16
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 128;
    const int kvh = h / (H / Hkv);
    const int key = k0 + row;
    const bool valid_k = key < S;
    const size_t head_off = ((size_t)(b * H + h) * S) * D;
    const int nq = (S + 127) / 128;
    constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_DV = 256, TM_DK = 256 + D, TM_DQ = 256 + 2 * D;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&mbar[0], 1); tc::mbar_init(&mbar[1], 1); tc::mbar_fence_init(); }
    {
        uint4 r[D / 8];
        const bf16* base = (half == 0 ? Kb : Vb) + ((size_t)(b * Hkv + kvh) * S + (valid_k ? key : 0)) * D;
        load_row<D>(base, valid_k, r);
        store_row<D>(half == 0 ? Kt : Vt, row, r);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sK = tc::smem_u32(Kt), sV = tc::smem_u32(Vt), sQ = tc::smem_u32(Qs), sdO = tc::smem_u32(dOs);
    const uint32_t sPT = tc::smem_u32(PTs), sdS = tc::smem_u32(dSs);
    constexpr uint32_t idesc128 = tc::make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescKM = tc::make_idesc_bf16(128, D, 0, 1);     // A K-major, B MN-major
    constexpr uint32_t idescMM = tc::make_idesc_bf16(128, D, 1, 1);     // A MN-major, B MN-major
    uint32_t ph0 = 0, ph1 = 0;

    for (int i = 0; i < nq; ++i) {
        const int q0 = i * 128;
        {
            const int qq = q0 + row;
            const bool vq = qq < S;
            uint4 r[D / 8];
            load_row<D>((half == 0 ? Qb : dOb) + head_off + (size_t)(vq ? qq : 0) * D, vq, r);
            store_row<D>(half == 0 ? Qs : dOs, row, r);
            if (half == 0) lse_s[row] = vq ? lse[((size_t)b * H + h) * S + qq] : INFINITY;
            else D_s[row] = vq ? Dvec[((size_t)b * H + h) * S + qq] : 0.f;
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int s = 0; s < D / 16; ++s)
                tc::mma_bf16(tmem + TM_ST, tc::desc_kmajor(sK, 128, s), tc::desc_kmajor(sQ, 128, s), idesc128, s > 0);
#pragma unroll
            for (int s = 0; s < D / 16; ++s)
                tc::mma_bf16(tmem + TM_DPT, tc::desc_kmajor(sV, 128, s), tc::desc_kmajor(sdO, 128, s), idesc128, s > 0);
            tc::mma_commit(&mbar[0]);
        }
        tc::mbar_wait(&mbar[0], ph0); ph0 ^= 1;
        tc::fence_after_sync();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c0 = half * 64 + cc * 32;
            float st[32], dp[32];
            tc::tmem_ld32(tlane + TM_ST + c0, st);
            tc::tmem_ld32(tlane + TM_DPT + c0, dp);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float p[8], ds[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int qc = c0 + c8 * 8 + c;
                    const float pv = valid_k ? exp2f(st[c8 * 8 + c] * scale_log2 - lse_s[qc]) : 0.f;
                    p[c] = pv;
                    ds[c] = pv * (dp[c8 * 8 + c] - D_s[qc]) * scale;
                }
                uint4 o;
                o.x = tc::pack_bf16(p[0], p[1]); o.y = tc::pack_bf16(p[2], p[3]);
                o.z = tc::pack_bf16(p[4], p[5]); o.w = tc::pack_bf16(p[6], p[7]);
                const int chunk = (c0 >> 3) + c8;
                *reinterpret_cast<uint4*>(PTs + chunk * (128 * 16) + row * 16) = o;
                o.x = tc::pack_bf16(ds[0], ds[1]); o.y = tc::pack_bf16(ds[2], ds[3]);
                o.z = tc::pack_bf16(ds[4], ds[5]); o.w = tc::pack_bf16(ds[6], ds[7]);
                *reinterpret_cast<uint4*>(dSs + chunk * (128 * 16) + row * 16) = o;
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int s = 0; s < 8; ++s)   // dV[key,d] += P^T[key,q] dO[q,d]
                tc::mma_bf16(tmem + TM_DV, tc::desc_kmajor(sPT, 128, s), tc::desc_mnmajor(sdO, 128, s), idescKM, (i > 0) || (s > 0));
#pragma unroll
            for (int s = 0; s < 8; ++s)   // dK[key,d] += dS^T[key,q] Q[q,d]
                tc::mma_bf16(tmem + TM_DK, tc::desc_kmajor(sdS, 128, s), tc::desc_mnmajor(sQ, 128, s), idescKM, (i > 0) || (s > 0));
#pragma unroll
            for (int s = 0; s < 8; ++s)   // dQ[q,d] = dS[q,key] K[key,d]
                tc::mma_bf16(tmem + TM_DQ, tc::desc_mnmajor(sdS, 128, s), tc::desc_mnmajor(sK, 128, s), idescMM, s > 0);
            tc::mma_commit(&mbar[1]);
        }
        tc::mbar_wait(&mbar[1], ph1); ph1 ^= 1;
        tc::fence_after_sync();
        {
            const int qq = q0 + row;                 // dQ rows are queries
            float* dst = dQacc + head_off + (size_t)qq * D + half * (D / 2);
            if constexpr (D == 32) {
                float t[16];
                tc::tmem_ld16(tlane + TM_DQ + half * 16, t);
                if (qq < S) {
#pragma unroll
                    for (int c = 0; c < 16; c += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]));
                }
            } else {
                float t[32];
                tc::tmem_ld32(tlane + TM_DQ + half * 32, t);
                if (qq < S) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]));
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    // ---- epilogue: dK, dV of this key tile ----
    {
        float* dk = dKh + head_off + (size_t)key * D + half * (D / 2);
        float* dv = dVh + head_off + (size_t)key * D + half * (D / 2);
        if constexpr (D == 32) {
            float t[16];
            tc::tmem_ld16(tlane + TM_DK + half * 16, t);
            if (valid_k) for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(dk + c) = make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]);
            tc::tmem_ld16(tlane + TM_DV + half * 16, t);
            if (valid_k) for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(dv + c) = make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]);
        } else {
            float t[32];
            tc::tmem_ld32(tlane + TM_DK + half * 32, t);
            if (valid_k) for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dk + c) = make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]);
            tc::tmem_ld32(tlane + TM_DV + half * 32, t);
            if (valid_k) for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dv + c) = make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// --------------------------------------------------------------------------- host side
struct AttnWs {
    bf16 *Qb, *Kb, *Vb, *dOb;
    float *Dvec, *dQacc, *dKh, *dVh;
};
static size_t attn_ws_bytes(int64_t B, int64_t S, int H, int Hkv, int d) {
    const size_t qe = (size_t)B * H * S * d, ke = (size_t)B * Hkv * S * d;
    return align_up(qe * 2) * 2 + align_up(ke * 2) * 2 + align_up((size_t)B * H * S * 4) + 3 * align_up(qe * 4) + 1024;
}
static bool attn_carve(AttnWs& w, void* ws, size_t bytes, int64_t B, int64_t S, int H, int Hkv, int d) {
    Arena ar(ws, bytes);
    const size_t qe = (size_t)B * H * S * d, ke = (size_t)B * Hkv * S * d;
    w.Qb = ar.take<bf16>(qe); w.Kb = ar.take<bf16>(ke); w.Vb = ar.take<bf16>(ke); w.dOb = ar.take<bf16>(qe);
    w.Dvec = ar.take<float>((size_t)B * H * S);
    w.dQacc = ar.take<float>(qe); w.dKh = ar.take<float>(qe); w.dVh = ar.take<float>(qe);
    return ar.ok();
}
static int attn_check(int64_t B, int64_t S, int H, int Hkv, int d) {
    GAOT_CHECK_ARG(B >= 1 && S >= 1 && H >= 1 && Hkv >= 1 && H % Hkv == 0, "attn: bad shape");
    GAOT_CHECK_ARG(B <= 65535 && H <= 65535 && S < ((int64_t)1 << 24), "attn: shape too large");
    if (d != 32 && d != 64) { set_error("attn: head_dim %d unsupported (32 or 64)", d); return GAOT_ERR_UNSUPPORTED; }
    return GAOT_OK;
}
static inline unsigned nb256(int64_t n) { return (unsigned)((n + 255) / 256); }

static int attn_prep_all(const float* q, const float* k, const float* v, const AttnWs& w, int64_t B, int64_t S,
                         int H, int Hkv, int d, const float* freqs, cudaStream_t st) {
    attn_prep_kernel<<<nb256(B * S * H * (d / 8)), 256, 0, st>>>(q, w.Qb, B, S, H, d, freqs);
    GAOT_LAUNCH_CHECK();
    attn_prep_kernel<<<nb256(B * S * Hkv * (d / 8)), 256, 0, st>>>(k, w.Kb, B, S, Hkv, d, freqs);
    GAOT_LAUNCH_CHECK();
    attn_prep_kernel<<<nb256(B * S * Hkv * (d / 8)), 256, 0, st>>>(v, w.Vb, B, S, Hkv, d, nullptr);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // namespace gaot

using namespace gaot;

extern "C" {

size_t gaot_attn_workspace_bytes(int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d) {
    return attn_ws_bytes(B, S, H, Hkv, d);
}

int gaot_attn_forward(const float* q, const float* k, const float* v, int64_t B, int64_t S, int32_t H,
                      int32_t Hkv, int32_t d, const float* rope_freqs, void* ws, size_t ws_bytes, float* out,
                      float* lse, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = attn_check(B, S, H, Hkv, d);
    if (rc) return rc;
    AttnWs w;
    if (!attn_carve(w, ws, ws_bytes, B, S, H, Hkv, d)) { set_error("attn_forward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    rc = attn_prep_all(q, k, v, w, B, S, H, Hkv, d, rope_freqs, st);
    if (rc) return rc;
    const float scale_log2 = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
    dim3 grid((unsigned)((S + 127) / 128), (unsigned)H, (unsigned)B);
    GAOT_TIME_KERNEL("attn_fwd", st, 4.0 * (double)B * H * (double)S * (double)S * d);
    if (d == 32) {
        const size_t smem = 5 * 128 * 32 * 2 + 128 * 128 * 2;
        GAOT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_fwd_kernel<32><<<grid, 128, smem, st>>>(w.Qb, w.Kb, w.Vb, out, lse, (int)S, H, Hkv, scale_log2);
    } else {
        const size_t smem = 5 * 128 * 64 * 2 + 128 * 128 * 2;
        GAOT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_fwd_kernel<64><<<grid, 128, smem, st>>>(w.Qb, w.Kb, w.Vb, out, lse, (int)S, H, Hkv, scale_log2);
    }
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_attn_backward(const float* q, const float* k, const float* v, const float* out, const float* d_out,
                       const float* lse, int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d,
                       const float* rope_freqs, void* ws, size_t ws_bytes, float* dq, float* dk, float* dv,
                       void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = attn_check(B, S, H, Hkv, d);
    if (rc) return rc;
    AttnWs w;
    if (!attn_carve(w, ws, ws_bytes, B, S, H, Hkv, d)) { set_error("attn_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    rc = attn_prep_all(q, k, v, w, B, S, H, Hkv, d, rope_freqs, st);
    if (rc) return rc;
    attn_bwd_prep_kernel<<<nb256(B * S * H), 256, 0, st>>>(d_out, out, w.dOb, w.Dvec, B, S, H, d);
    GAOT_LAUNCH_CHECK();
    GAOT_CUDA(cudaMemsetAsync(w.dQacc, 0, (size_t)B * H * S * d * sizeof(float), st));
    const float scale = 1.0f / sqrtf((float)d), scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((unsigned)((S + 127) / 128), (unsigned)H, (unsigned)B);
    {
    GAOT_TIME_KERNEL("attn_bwd", st, 10.0 * (double)B * H * (double)S * (double)S * d);
    if (d == 32) {
        const size_t smem = 4 * 128 * 32 * 2 + 2 * 128 * 128 * 2;
        GAOT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_bwd_kernel<32><<<grid, 256, smem, st>>>(w.Qb, w.Kb, w.Vb, w.dOb, lse, w.Dvec, w.dQacc, w.dKh, w.dVh,
                                                     (int)S, H, Hkv, scale, scale_log2);
    } else {
        const size_t smem = 4 * 128 * 64 * 2 + 2 * 128 * 128 * 2;
        GAOT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_bwd_kernel<64><<<grid, 256, smem, st>>>(w.Qb, w.Kb, w.Vb, w.dOb, lse, w.Dvec, w.dQacc, w.dKh, w.dVh,
                                                     (int)S, H, Hkv, scale, scale_log2);
    }
    }
    GAOT_LAUNCH_CHECK();
    attn_bwd_post_kernel<<<nb256(B * S * H * (d / 8)), 256, 0, st>>>(w.dQacc, dq, B, S, H, H, d, rope_freqs);
    GAOT_LAUNCH_CHECK();
    attn_bwd_post_kernel<<<nb256(B * S * Hkv * (d / 8)), 256, 0, st>>>(w.dKh, dk, B, S, H, Hkv, d, rope_freqs);
    GAOT_LAUNCH_CHECK();
    attn_bwd_post_kernel<<<nb256(B * S * Hkv * (d / 8)), 256, 0, st>>>(w.dVh, dv, B, S, H, Hkv, d, nullptr);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // extern "C"
