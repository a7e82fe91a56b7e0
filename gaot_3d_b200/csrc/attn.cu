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
#include <cstdlib>

namespace gaot {

using bf16 = __nv_bfloat16;

// --------------------------------------------------------------------------- operand layouts
// Q, K, V and dO are all stored PRE-TILED in bf16: per (b,h) ceil(S/128) tiles of 128 rows, each
// tile already in the kernels' chunk-major shared-memory layout (tc05.cuh: 16-byte chunk c of row r at c*128*16 + r*16),
// rows >= S zero.  A tile is one contiguous block that the backward's loader moves with a single cp.async.bulk (TMA
// engine, async proxy: no generic->async proxy fence in front of the tensor core).
// Conventions that let the tensor core produce the softmax arguments and row sums directly:
//   * Q is stored multiplied by (1/sqrt(d)) * log2(e), so S = Q K^T is already in the scaled log2 domain;
//   * every tile carries ONE extra chunk (index d/8) per row: for Q it holds (-lse_hi, -lse_lo, 0...), for dO
//     (-D_hi, -D_lo, 0...) as bf16 hi/lo pairs, written by the backward prep kernel.  With K and V extended by a chunk
//     (1, 1, 0...) the products K' Q'^T = S - lse and V' dO'^T = dP - D need no per-element correction;
//   * V's extra chunk is (1, 0...) for real keys (0 for padding): P V' returns the row sum of P in column d.  K's is 0.
__host__ __device__ __forceinline__ int64_t pad128(int64_t S) { return (S + 127) / 128 * 128; }
__host__ __device__ __forceinline__ int64_t tile_elems(int d) { return 128 * (int64_t)(d + 8); }
__device__ __forceinline__ size_t tiled_off(int64_t head, int64_t S_pad, int64_t s, int d, int ch) {
    return ((size_t)head * (S_pad >> 7) + (size_t)(s >> 7)) * tile_elems(d) + (size_t)ch * (128 * 8) + (size_t)(s & 127) * 8;
}
__device__ __forceinline__ uint4 hilo_chunk(float x) {        // bf16 (hi, lo) split of x, rest of the chunk zero
    const bf16 hi = __float2bfloat16_rn(x);
    const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    return make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16), 0u, 0u, 0u);
}

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
                 const float* __restrict__ freqs /* NULL: no rope */, int tiled, float qscale,
                 int extra /* tile's extra chunk: 0 leave, 1 zeros, 2 (1,0...) for rows < S */) {
    const int cpr = d >> 3;                                   // 8-element chunks per head row
    const int64_t Sx = tiled ? pad128(S) : S;
    const int64_t total = B * Sx * nh * cpr;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % cpr);
    const int h = (int)((idx / cpr) % nh);
    const int64_t s = (idx / ((int64_t)cpr * nh)) % Sx;
    const int64_t b = idx / ((int64_t)cpr * nh * Sx);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (s < S) {
        const float* p = src + ((b * S + s) * nh + h) * d + ch * 8;
        const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
        float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (freqs) rope8<true>(x, (int)s, ch * 8, freqs, false);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] *= qscale;
        o.x = tc::pack_bf16(x[0], x[1]); o.y = tc::pack_bf16(x[2], x[3]);
        o.z = tc::pack_bf16(x[4], x[5]); o.w = tc::pack_bf16(x[6], x[7]);
    }
    bf16* out = tiled ? dst + tiled_off(b * nh + h, Sx, s, d, ch) : dst + ((b * nh + h) * S + s) * d + ch * 8;
    *reinterpret_cast<uint4*>(out) = o;
    if (tiled && extra && ch == 0)
        *reinterpret_cast<uint4*>(dst + tiled_off(b * nh + h, Sx, s, d, cpr)) = make_uint4((extra == 2 && s < S) ? 0x00003F80u : 0u, 0u, 0u, 0u);
}

// fused-projection variant: qkv bf16 [B*S, ld] (columns: H*d of q | Hkv*d of k | Hkv*d of v, the output of
// one [Wq;Wk;Wv] GEMM) -> Qb [B,H,S,d], Kb, Vb [B,Hkv,S,d] bf16 with RoPE on q and k; one launch.
__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__global__ void __launch_bounds__(256)
attn_pack_qkv_kernel(const bf16* __restrict__ qkv, int64_t ld, bf16* __restrict__ Qb, bf16* __restrict__ Kb,
                     bf16* __restrict__ Vb, int64_t B, int64_t S, int H, int Hkv, int d, const float* __restrict__ freqs, float qscale) {
    const int cpr = d >> 3, nh = H + 2 * Hkv;
    const int64_t Sp = pad128(S);
    const int64_t total = B * Sp * nh * cpr;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % cpr);
    const int hh = (int)((idx / cpr) % nh);
    const int64_t s = (idx / ((int64_t)cpr * nh)) % Sp;
    const int64_t b = idx / ((int64_t)cpr * nh * Sp);
    bf16* dst; bool rope; int extra;                           // extra chunk: 0 leave (Q: written by the backward prep), 1 zeros, 2 (1,0...)
    if (hh < H) { dst = Qb + tiled_off(b * H + hh, Sp, s, d, ch); rope = true; extra = 0; }
    else if (hh < H + Hkv) { dst = Kb + tiled_off(b * Hkv + (hh - H), Sp, s, d, ch); rope = true; extra = 1; }
    else { dst = Vb + tiled_off(b * Hkv + (hh - H - Hkv), Sp, s, d, ch); rope = false; extra = 2; }
    if (extra && ch == 0)
        *reinterpret_cast<uint4*>(dst + (size_t)cpr * (128 * 8)) = make_uint4((extra == 2 && s < S) ? 0x00003F80u : 0u, 0u, 0u, 0u);
    if (s >= S) {                                             // padding rows of the tiles are zero
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(qkv + (b * S + s) * ld + (int64_t)hh * d + ch * 8));
    uint4 o = raw;
    if ((rope && freqs) || hh < H) {
        float x[8];
        unpack_bf16x8(raw, x);
        if (rope && freqs) rope8<true>(x, (int)s, ch * 8, freqs, false);
        if (hh < H) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] *= qscale;
        }
        o.x = tc::pack_bf16(x[0], x[1]); o.y = tc::pack_bf16(x[2], x[3]);
        o.z = tc::pack_bf16(x[4], x[5]); o.w = tc::pack_bf16(x[6], x[7]);
    }
    *reinterpret_cast<uint4*>(dst) = o;
}

// fused-block variant of the backward prep: dO arrives token-major in bf16 (the epilogue of the o_proj dX GEMM), O as
// the forward's UNROUNDED fp32 copy.  D must cancel against the kernel's own dP = dO_bf16 . V (dS = P (dP - D)), so it
// is formed from the bf16 dO the MMA consumes and the fp32 O; a bf16-rounded O leaves a 2^-9 error along mean(K) that
// swamps dQ / dK wherever the attention is close to uniform.
__global__ void __launch_bounds__(256)
attn_bwd_prep_bf16_kernel(const bf16* __restrict__ dO, const float* __restrict__ O, bf16* __restrict__ dOb,
                          float* __restrict__ Dvec, bf16* __restrict__ Qb, const float* __restrict__ lse, int fold_D,
                          int64_t B, int64_t S, int H, int d) {
    const int64_t Sp = pad128(S);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (b, s, h) over the padded length
    if (idx >= B * Sp * H) return;
    const int h = (int)(idx % H);
    const int64_t s = (idx / H) % Sp, b = idx / ((int64_t)H * Sp);
    bf16* out = dOb + tiled_off(b * H + h, Sp, s, d, 0);
    bf16* qx = Qb + tiled_off(b * H + h, Sp, s, d, d >> 3);            // the tile's extra chunk of this row
    if (s >= S) {
        for (int c = 0; c <= d; c += 8) *reinterpret_cast<uint4*>(out + (c >> 3) * (128 * 8)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(qx) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const bf16* pd = dO + ((b * S + s) * H + h) * d;
    const float* po = O + ((b * S + s) * H + h) * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(pd + c));
        const float4 o0 = __ldg(reinterpret_cast<const float4*>(po + c)), o1 = __ldg(reinterpret_cast<const float4*>(po + c + 4));
        float fa[8];
        unpack_bf16x8(a, fa);
        acc += fa[0] * o0.x + fa[1] * o0.y + fa[2] * o0.z + fa[3] * o0.w + fa[4] * o1.x + fa[5] * o1.y + fa[6] * o1.z + fa[7] * o1.w;
        *reinterpret_cast<uint4*>(out + (c >> 3) * (128 * 8)) = a;
    }
    Dvec[(b * H + h) * S + s] = acc;
    *reinterpret_cast<uint4*>(out + (d >> 3) * (128 * 8)) = fold_D ? hilo_chunk(-acc) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(qx) = hilo_chunk(-lse[(b * H + h) * S + s]);
}

// fused-block backward post: per-head fp32 dQ / dK / dV -> ONE bf16 token-major [B*S, ld] gradient of the fused
// qkv projection output (GQA group summed, RoPE undone), ready to be the A operand of the dX / dW GEMMs.
__global__ void __launch_bounds__(256)
attn_bwd_post_qkv_kernel(const float* __restrict__ dQh, const float* __restrict__ dKh, const float* __restrict__ dVh,
                         bf16* __restrict__ out, int64_t ld, int64_t B, int64_t S, int H, int Hkv, int d,
                         const float* __restrict__ freqs) {
    const int cpr = d >> 3, nh = H + 2 * Hkv;
    const int64_t total = B * S * nh * cpr;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % cpr);
    const int hh = (int)((idx / cpr) % nh);
    const int64_t s = (idx / ((int64_t)cpr * nh)) % S;
    const int64_t b = idx / ((int64_t)cpr * nh * S);
    const float* src; int h0, rep; bool rope;
    if (hh < H) { src = dQh; h0 = hh; rep = 1; rope = true; }
    else if (hh < H + Hkv) { src = dKh; rep = H / Hkv; h0 = (hh - H) * rep; rope = true; }
    else { src = dVh; rep = H / Hkv; h0 = (hh - H - Hkv) * rep; rope = false; }
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < rep; ++r) {
        const float* p = src + ((b * H + h0 + r) * S + s) * d + ch * 8;
        const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
        x[0] += v0.x; x[1] += v0.y; x[2] += v0.z; x[3] += v0.w; x[4] += v1.x; x[5] += v1.y; x[6] += v1.z; x[7] += v1.w;
    }
    if (rope && freqs) rope8<true>(x, (int)s, ch * 8, freqs, true);
    uint4 o;
    o.x = tc::pack_bf16(x[0], x[1]); o.y = tc::pack_bf16(x[2], x[3]);
    o.z = tc::pack_bf16(x[4], x[5]); o.w = tc::pack_bf16(x[6], x[7]);
    *reinterpret_cast<uint4*>(out + (b * S + s) * ld + (int64_t)hh * d + ch * 8) = o;
}

// backward prep: dO -> bf16 [B,H,S,d] and Dvec[b,h,s] = sum_c dO*O
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const float* __restrict__ dO, const float* __restrict__ O, bf16* __restrict__ dOb,
                     float* __restrict__ Dvec, bf16* __restrict__ Qb, const float* __restrict__ lse, int fold_D,
                          int64_t B, int64_t S, int H, int d) {
    const int64_t Sp = pad128(S);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (b, s, h) over the padded length
    if (idx >= B * Sp * H) return;
    const int h = (int)(idx % H);
    const int64_t s = (idx / H) % Sp, b = idx / ((int64_t)H * Sp);
    bf16* out = dOb + tiled_off(b * H + h, Sp, s, d, 0);
    bf16* qx = Qb + tiled_off(b * H + h, Sp, s, d, d >> 3);            // the tile's extra chunk of this row
    if (s >= S) {
        for (int c = 0; c <= d; c += 8) *reinterpret_cast<uint4*>(out + (c >> 3) * (128 * 8)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(qx) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const float* pd = dO + ((b * S + s) * H + h) * d;
    const float* po = O + ((b * S + s) * H + h) * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(pd + c), a1 = *reinterpret_cast<const float4*>(pd + c + 4);
        const float4 o0 = *reinterpret_cast<const float4*>(po + c), o1 = *reinterpret_cast<const float4*>(po + c + 4);
        acc += a0.x * o0.x + a0.y * o0.y + a0.z * o0.z + a0.w * o0.w + a1.x * o1.x + a1.y * o1.y + a1.z * o1.z + a1.w * o1.w;
        uint4 o;
        o.x = tc::pack_bf16(a0.x, a0.y); o.y = tc::pack_bf16(a0.z, a0.w);
        o.z = tc::pack_bf16(a1.x, a1.y); o.w = tc::pack_bf16(a1.z, a1.w);
        *reinterpret_cast<uint4*>(out + (c >> 3) * (128 * 8)) = o;
    }
    Dvec[(b * H + h) * S + s] = acc;
    *reinterpret_cast<uint4*>(out + (d >> 3) * (128 * 8)) = fold_D ? hilo_chunk(-acc) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(qx) = hilo_chunk(-lse[(b * H + h) * S + s]);
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
// row `s` of a pre-tiled operand (head base `g`, see tiled_off)
template <int D>
__device__ __forceinline__ void load_row_tiled(const bf16* __restrict__ g, int64_t s, bool valid, uint4 (&r)[D / 8]) {
    const bf16* p = g + (size_t)(s >> 7) * tile_elems(D) + (size_t)(s & 127) * 8;
#pragma unroll
    for (int c = 0; c < D / 8; ++c)
        r[c] = valid ? *reinterpret_cast<const uint4*>(p + c * (128 * 8)) : make_uint4(0u, 0u, 0u, 0u);
}
template <int D>
__device__ __forceinline__ void store_row(uint8_t* tile, int row, const uint4 (&r)[D / 8]) {
#pragma unroll
    for (int c = 0; c < D / 8; ++c) *reinterpret_cast<uint4*>(tile + c * (128 * 16) + row * 16) = r[c];
}

// --------------------------------------------------------------------------- forward
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], minimax cubic for 2^f
// (max rel. error 1.0e-4, far below the bf16 rounding of P), n added into the exponent field.  x is clamped at -125
// (-inf scores of padding keys give 2^-125 ~ 2e-38).  Used for a compile-time fraction of the softmax elements so that the
// MUFU pipe (16 ex2 / clk / SM: the floor of this kernel at head_dim 32) and the FMA pipe share the exponentials.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: n lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05500893f, 0.24221096f);
    p = fmaf(p, f, 0.69328293f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Counter-based dropout mask shared by forward and backward (and by oracle/attn.py::dropout_keep):
// keep(b,h,q,k) = lowbias32(rowkey(b,h,q) ^ k * 0x85EBCA6B) >= thresh, rowkey = lowbias32(seed ^ row * 0x9E3779B1) + seed_hi
struct DropCfg { uint32_t thresh; uint32_t seed_lo, seed_hi; float inv_keep; };
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t drop_rowkey(const DropCfg& dc, uint32_t global_row) {
    return lowbias32(dc.seed_lo ^ (global_row * 0x9E3779B1u)) + dc.seed_hi;
}
__device__ __forceinline__ bool drop_keep(const DropCfg& dc, uint32_t rowkey, uint32_t k) {
    return lowbias32(rowkey ^ (k * 0x85EBCA6Bu)) >= dc.thresh;
}

// V tiles carry 16 extra columns: column D is all ones, so the P*V tensor-core product also
// returns the row sum of the (bf16-rounded) probabilities -- no per-element FADD in the softmax.
// Software pipeline (one mbarrier wait + one CTA barrier per key tile): after the softmax of tile j
// the issuing thread queues P_j*V_j AND S_{j+1} = Q K_{j+1}^T back to back; the O update with the
// P_{j-1}*V_{j-1} result is deferred to the top of the next iteration.
template <int D, bool DROP>
__global__ void __launch_bounds__(128, 2)
attn_fwd_kernel(const bf16* __restrict__ Qb, const bf16* __restrict__ Kb, const bf16* __restrict__ Vb,
                void* __restrict__ out_v, int out_bf16, float* __restrict__ out32, float* __restrict__ lse, int S, int H, int Hkv,
                const DropCfg dc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    constexpr int TILE_B = 128 * D * 2;
    constexpr int VTILE_B = 128 * (D + 16) * 2;
    constexpr uint32_t TM_S = 0, TM_PV = 128;
    uint8_t* Qs = sm;
    uint8_t* Ks = sm + TILE_B;                          // [2]
    uint8_t* Vs = sm + 3 * TILE_B;                      // [2] x (D+16 columns)
    uint8_t* Ps = sm + 3 * TILE_B + 2 * VTILE_B;        // 128 x 128 bf16
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int kvh = h / (H / Hkv);
    const int q = q0 + tid;
    const bool valid_q = q < S;
    const bf16* Kbase = Kb + (size_t)(b * Hkv + kvh) * (pad128(S) >> 7) * tile_elems(D);   // pre-tiled operands
    const bf16* Vbase = Vb + (size_t)(b * Hkv + kvh) * (pad128(S) >> 7) * tile_elems(D);
    const int nkv = (S + 127) / 128;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    uint4 kreg[D / 8], vreg[D / 8];                     // register-staged K/V rows of a future tile
    {
        uint4 r[D / 8];
        load_row_tiled<D>(Qb + (size_t)(b * H + h) * (pad128(S) >> 7) * tile_elems(D), valid_q ? q : 0, valid_q, r);
        store_row<D>(Qs, tid, r);
        const bool vk = tid < S;
        load_row_tiled<D>(Kbase, vk ? tid : 0, vk, r);
        store_row<D>(Ks, tid, r);
        load_row_tiled<D>(Vbase, vk ? tid : 0, vk, r);
        store_row<D>(Vs, tid, r);
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {                // constant ones / zero chunks (bf16 1.0 = 0x3F80)
            *reinterpret_cast<uint4*>(Vs + bb * VTILE_B + (D / 8) * (128 * 16) + tid * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(Vs + bb * VTILE_B + (D / 8 + 1) * (128 * 16) + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (nkv > 1) {
            const int kn = 128 + tid;
            const bool v1 = kn < S;
            load_row_tiled<D>(Kbase, v1 ? kn : 0, v1, kreg);
            load_row_tiled<D>(Vbase, v1 ? kn : 0, v1, vreg);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t sQ = tc::smem_u32(Qs), sP = tc::smem_u32(Ps);
    constexpr uint32_t idescS = tc::make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescPV = tc::make_idesc_bf16(128, D + 16, 0, 1);
    const tc::Desc dQ_ = tc::kmajor(sQ, 128), dP_ = tc::kmajor(sP, 128);
    const tc::Desc dK0 = tc::kmajor(tc::smem_u32(Ks), 128), dK1 = tc::kmajor(tc::smem_u32(Ks + TILE_B), 128);
    const tc::Desc dV0 = tc::mnmajor(tc::smem_u32(Vs), 128), dV1 = tc::mnmajor(tc::smem_u32(Vs + VTILE_B), 128);
    constexpr uint32_t KS = tc::kstep_kmajor(128);
    if (warp == 0) {
        if (tc::elect_one()) {
#pragma unroll
            for (int s = 0; s < D / 16; ++s)
                tc::mma_bf16(tmem + TM_S, dQ_.adv(s * KS).u64(), dK0.adv(s * KS).u64(), idescS, s > 0);
            tc::mma_commit(&mbar);
        }
        __syncwarp();
    }

    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;     // m in the scaled log2 domain
    float rs_prev = 0.f;                                // DROP: row sum of the un-dropped probabilities of the previous tile
    const uint32_t rowkey = DROP ? drop_rowkey(dc, (uint32_t)((b * H + h) * S + q)) : 0u;
    float O[D];
#pragma unroll
    for (int c = 0; c < D; ++c) O[c] = 0.f;
    uint32_t ph = 0;

    for (int j = 0; j <= nkv; ++j) {
        tc::mbar_wait(&mbar, ph); ph ^= 1;              // S_j (j < nkv) and P_{j-1} V_{j-1} (j > 0) are complete
        tc::fence_after_sync();
        if (j > 0) {                                    // deferred O / l update with tile j-1
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
                float t[32];
                tc::tmem_ld32(tlane + TM_PV + c0, t);
#pragma unroll
                for (int c = 0; c < 32; ++c) O[c0 + c] = fmaf(O[c0 + c], alpha_prev, t[c]);
            }
            if (DROP) {
                l = fmaf(l, alpha_prev, rs_prev);
            } else {
                float t[16];
                tc::tmem_ld16(tlane + TM_PV + D, t);
                l = fmaf(l, alpha_prev, t[0]);
            }
        }
        if (j == nkv) break;
        const int buf = j & 1;
        if (j + 1 < nkv) {                              // K/V of tile j+1 -> the other buffer (its readers finished)
            store_row<D>(Ks + (buf ^ 1) * TILE_B, tid, kreg);
            store_row<D>(Vs + (buf ^ 1) * VTILE_B, tid, vreg);
        }
        if (j + 2 < nkv) {                              // start fetching tile j+2
            const int kn = (j + 2) * 128 + tid;
            const bool vk = kn < S;
            load_row_tiled<D>(Kbase, vk ? kn : 0, vk, kreg);
            load_row_tiled<D>(Vbase, vk ? kn : 0, vk, vreg);
        }
        float sv[128];
        {
            uint32_t r0[32], r1[32], r2[32], r3[32];
            tc::tmem_ld32_nowait(tlane + TM_S, r0);
            tc::tmem_ld32_nowait(tlane + TM_S + 32, r1);
            tc::tmem_ld32_nowait(tlane + TM_S + 64, r2);
            tc::tmem_ld32_nowait(tlane + TM_S + 96, r3);
            tc::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                sv[c] = __uint_as_float(r0[c]); sv[32 + c] = __uint_as_float(r1[c]);
                sv[64 + c] = __uint_as_float(r2[c]); sv[96 + c] = __uint_as_float(r3[c]);
            }
        }
        const int kvalid = S - j * 128;                 // keys >= kvalid in this tile are padding
        if (kvalid < 128) {
#pragma unroll
            for (int c = 0; c < 128; ++c) if (c >= kvalid) sv[c] = -INFINITY;
        }
        float mr = fmax3(sv[0], sv[1], sv[2]);
#pragma unroll
        for (int c = 3; c + 1 < 128; c += 2) mr = fmax3(mr, sv[c], sv[c + 1]);
        mr = fmaxf(mr, sv[127]);
        const float mx = fmaxf(m, mr);
        alpha_prev = ex2_approx(m - mx);                // m = -inf on the first tile -> 0
        const float nmx = -mx;
        float rs = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 16; ++c8) {
            float p[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                p[c] = ex2_approx(sv[c8 * 8 + c] + nmx);       // Q is pre-scaled: the scores are already in the log2 domain
                if (DROP) {
                    rs += p[c];
                    p[c] = drop_keep(dc, rowkey, (uint32_t)(j * 128 + c8 * 8 + c)) ? p[c] * dc.inv_keep : 0.f;
                }
            }
            uint4 o;
            o.x = tc::pack_bf16(p[0], p[1]); o.y = tc::pack_bf16(p[2], p[3]);
            o.z = tc::pack_bf16(p[4], p[5]); o.w = tc::pack_bf16(p[6], p[7]);
            *reinterpret_cast<uint4*>(Ps + c8 * (128 * 16) + tid * 16) = o;
        }
        m = mx;
        rs_prev = rs;
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const tc::Desc dV = buf ? dV1 : dV0;
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    tc::mma_bf16(tmem + TM_PV, dP_.adv(s * KS).u64(), dV.adv(s * tc::KSTEP_MN).u64(), idescPV, s > 0);
                if (j + 1 < nkv) {
                    const tc::Desc dK = buf ? dK0 : dK1;
#pragma unroll
                    for (int s = 0; s < D / 16; ++s)
                        tc::mma_bf16(tmem + TM_S, dQ_.adv(s * KS).u64(), dK.adv(s * KS).u64(), idescS, s > 0);
                }
                tc::mma_commit(&mbar);
            }
            __syncwarp();
        }
    }
    if (valid_q) {
        const float inv = 1.0f / l;
        const size_t ooff = ((size_t)b * S + q) * (H * D) + h * D;
        if (out_bf16) {
            bf16* o = reinterpret_cast<bf16*>(out_v) + ooff;
#pragma unroll
            for (int c = 0; c < D; c += 8) {
                uint4 pk;
                pk.x = tc::pack_bf16(O[c] * inv, O[c + 1] * inv); pk.y = tc::pack_bf16(O[c + 2] * inv, O[c + 3] * inv);
                pk.z = tc::pack_bf16(O[c + 4] * inv, O[c + 5] * inv); pk.w = tc::pack_bf16(O[c + 6] * inv, O[c + 7] * inv);
                *reinterpret_cast<uint4*>(o + c) = pk;
            }
            if (out32) {            // unrounded copy for the backward's D = rowsum(dO * O) (see attn_bwd_prep_bf16_kernel)
                float* o32 = out32 + ooff;
#pragma unroll
                for (int c = 0; c < D; c += 4)
                    *reinterpret_cast<float4*>(o32 + c) = make_float4(O[c] * inv, O[c + 1] * inv, O[c + 2] * inv, O[c + 3] * inv);
            }
        } else {
            float* o = reinterpret_cast<float*>(out_v) + ooff;
#pragma unroll
            for (int c = 0; c < D; c += 4)
                *reinterpret_cast<float4*>(o + c) = make_float4(O[c] * inv, O[c + 1] * inv, O[c + 2] * inv, O[c + 3] * inv);
        }
        lse[((size_t)b * H + h) * S + q] = m + log2f(l);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// --------------------------------------------------------------------------- forward, warp-specialised (head_dim 32)
// One CTA per SM owns 128 queries of one (batch, head) and walks the 128-key tiles.  The scores of a tile live in one
// of TWO TMEM stages, so the tensor core computes S of tile j+2 while the softmax warps work on tile j+1 and the P V
// products of tile j drain behind them; no CTA-wide barrier in the loop, the roles talk through mbarriers.
//   softmax warps (16): warp = (lane quarter q4, key group g): each thread owns ONE query row and the 32 keys of group g
//       of every tile -- four independent online-softmax streams per row, so no cross-warp max exchange.  Per tile:
//       tcgen05.ld 32 scores -> max -> P = exp2(S - m) -> bf16 pairs -> tcgen05.st INTO THE SAME TMEM COLUMNS (P aliases
//       S: only this thread ever reads those S values) -> arrive.  No shared-memory traffic, no proxy fence.
//       The stream's running output O_g stays in TMEM for the whole kernel; the reference max m is only raised when the
//       tile max exceeds it by more than 2^8 (the result is invariant to m; P <= 256 is harmless in bf16 / fp32), in
//       which case the warp rescales its O_g rows in place (tcgen05.ld / st) after the previous P V has completed.
//   MMA warp (1 lane): wait softmax(j) -> O_g += P_g V_g (A operand from TMEM, V' carries a ones column: column 32 of
//       O_g is the row sum) -> commit(pv) -> S(j+2) = Q K^T into the stage just consumed -> commit(S).
//   loader warp: one cp.async.bulk per pre-tiled K / V tile, three tiles deep, released by the P V barrier.
// End: the four streams of a row are merged through shared memory (max, rescale, sum), out = O / l, lse = m + log2 l.
// TMEM (512 columns): S / P stage s at 128 s (P_g at +32 g, 16 packed columns); O_g at 256 + 48 g.
namespace fw2 {
constexpr int D = 32, NST = 3;
constexpr int KT_B = 128 * D * 2;                   // 8 KB  K tile (the data chunks only)
constexpr int VT_B = 128 * (D + 16) * 2;            // 12 KB V' tile: 32 + ones chunk + zero chunk
constexpr int T_G = 128 * (D + 8) * 2;              // 10 KB: a pre-tiled operand tile in global memory
constexpr int OFF_Q = 0, OFF_K = KT_B, OFF_V = OFF_K + NST * KT_B, END_RING = OFF_V + NST * VT_B;   // 69632
constexpr int MERGE_B = 4 * 128 * 33 * 4 + 128 * 4 * 4;                                             // 69632
constexpr int SMEM = (END_RING > KT_B + MERGE_B ? END_RING : KT_B + MERGE_B);
constexpr uint32_t TM_O = 256;
constexpr float TAU = 8.0f;
constexpr int NSW = 16;                             // softmax warps
constexpr int W_MMA = NSW, W_LOAD = NSW + 1, THREADS = (NSW + 2) * 32;
}

// EMU = how many of every 32 softmax elements take the polynomial exponential (0, 4, 8, 12)
// PREF: the scores of tile j+1 are requested from TMEM as soon as tile j's registers are consumed (their load latency runs under
// the P store / fence / arrive tail of tile j)
template <bool DROP, int EMU, bool PREF = false>
__global__ void __launch_bounds__(fw2::THREADS, 1)
attn_fwd2_kernel(const bf16* __restrict__ Qb, const bf16* __restrict__ Kb, const bf16* __restrict__ Vb,
                 void* __restrict__ out_v, int out_bf16, float* __restrict__ out32, float* __restrict__ lse,
                 int S, int H, int Hkv, const DropCfg dc) {
    using namespace fw2;
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar_S[2], bar_sm[2], bar_pv, bar_load[NST];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int kvh = h / (H / Hkv);
    const int nkv = (S + 127) / 128;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&bar_S[i], 1); tc::mbar_init(&bar_sm[i], NSW); }
        tc::mbar_init(&bar_pv, 1);
#pragma unroll
        for (int i = 0; i < NST; ++i) tc::mbar_init(&bar_load[i], 1);
        tc::mbar_fence_init();
    }
    if (tid < 128 * NST) {                            // the zero chunk of every V' buffer (never overwritten)
        const int row = tid & 127, bf = tid >> 7;
        *reinterpret_cast<uint4*>(sm + OFF_V + bf * VT_B + 5 * (128 * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sbase = tc::smem_u32(sm);

    if (warp == W_LOAD) {
        // ------------------------------------------------------------------ loader
        if (lane == 0) {
            const bf16* qsrc = Qb + ((size_t)(b * H + h) * nkv + blockIdx.x) * (T_G / 2);
            const bf16* ksrc = Kb + (size_t)(b * Hkv + kvh) * nkv * (T_G / 2);
            const bf16* vsrc = Vb + (size_t)(b * Hkv + kvh) * nkv * (T_G / 2);
            for (int t = 0; t < nkv; ++t) {
                const int buf = t % NST;
                if (t >= NST) tc::mbar_wait(&bar_pv, (uint32_t)((t - NST) & 1));     // P V of the tile that used the buffer is done
                tc::mbar_arrive_expect_tx(&bar_load[buf], (t == 0 ? KT_B : 0) + KT_B + T_G);
                if (t == 0) tc::bulk_copy_g2s(sbase + OFF_Q, qsrc, KT_B, &bar_load[buf]);
                tc::bulk_copy_g2s(sbase + OFF_K + buf * KT_B, ksrc + (size_t)t * (T_G / 2), KT_B, &bar_load[buf]);
                tc::bulk_copy_g2s(sbase + OFF_V + buf * VT_B, vsrc + (size_t)t * (T_G / 2), T_G, &bar_load[buf]);
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ tensor-core issue (one lane)
        if (tc::elect_one()) {
            constexpr uint32_t idescS = tc::make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idescPV = tc::make_idesc_bf16(128, D + 16, 0, 1);   // A from TMEM (K-major), B = V' MN-major
            constexpr uint32_t KS = tc::kstep_kmajor(128);
            const tc::Desc dQ_ = tc::kmajor(sbase + OFF_Q, 128);
            const tc::Desc dK_ = tc::kmajor(sbase + OFF_K, 128);
            const tc::Desc dV_ = tc::mnmajor(sbase + OFF_V, 128);
            auto issue_S = [&](int j) {
                const uint32_t off = (j % NST) * KT_B;
#pragma unroll
                for (int k = 0; k < D / 16; ++k)
                    tc::mma_bf16(tmem + (uint32_t)(j & 1) * 128u, dQ_.adv(k * KS).u64(), dK_.adv(off + k * KS).u64(), idescS, k > 0);
            };
            tc::mbar_wait(&bar_load[0], 0);
            issue_S(0); tc::mma_commit(&bar_S[0]);
            if (nkv > 1) { tc::mbar_wait(&bar_load[1], 0); issue_S(1); tc::mma_commit(&bar_S[1]); }
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                tc::mbar_wait(&bar_sm[s], (uint32_t)((j >> 1) & 1));             // P(j) is in TMEM, the stage's S has been read
                tc::fence_after_sync();
                const uint32_t voff = (j % NST) * VT_B;
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int k = 0; k < 2; ++k)     // O_g[q, 0..47] += P_g[q, 16 keys] V'[16 keys, 0..47]
                        tc::mma_bf16_ts(tmem + TM_O + 48u * g, tmem + (uint32_t)s * 128u + 32u * g + 8u * k,
                                        dV_.adv(voff + (g * 2 + k) * tc::KSTEP_MN).u64(), idescPV, (j | k) != 0);
                tc::mma_commit(&bar_pv);
                if (j + 2 < nkv) {
                    tc::mbar_wait(&bar_load[(j + 2) % NST], (uint32_t)(((j + 2) / NST) & 1));
                    issue_S(j + 2);                 // overwrites the stage P(j) lives in: ordered behind the P V above
                    tc::mma_commit(&bar_S[s]);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax warps
        const int q4 = warp & 3, g = warp >> 2;
        const int row = q4 * 32 + lane, q = q0 + row;
        const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
        const uint32_t tO = tlane + TM_O + 48u * g;
        const uint32_t rowkey = DROP ? drop_rowkey(dc, (uint32_t)((b * H + h) * S + q)) : 0u;
        float m_used = -INFINITY, l_reg = 0.f;
        uint32_t svr[32];
        auto request_scores = [&](int j) {                // wait for S(j), then issue (not await) this thread's 32 columns
            tc::mbar_wait(&bar_S[j & 1], (uint32_t)((j >> 1) & 1));
            tc::fence_after_sync();
            tc::tmem_ld32_nowait(tlane + (uint32_t)(j & 1) * 128u + 32u * g, svr);
        };
        if (PREF) request_scores(0);
        for (int j = 0; j < nkv; ++j) {
            const int s = j & 1;
            const uint32_t tS = tlane + (uint32_t)s * 128u + 32u * g;
            if (!PREF) request_scores(j);
            tc::tmem_wait_ld();
            float sv[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) sv[c] = __uint_as_float(svr[c]);
            const int kvalid = S - (j * 128 + g * 32);   // keys >= kvalid of this group are padding
            if (kvalid < 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c) if (c >= kvalid) sv[c] = -INFINITY;
            }
            float mt = fmax3(sv[0], sv[1], sv[2]);
#pragma unroll
            for (int c = 3; c + 1 < 32; c += 2) mt = fmax3(mt, sv[c], sv[c + 1]);
            mt = fmaxf(mt, sv[31]);
            const bool raise = mt > m_used + TAU;        // also true on the first tile with a real key (m_used = -inf)
            if (__any_sync(0xffffffffu, raise)) {
                if (j > 0) {                             // rescale this warp's O_g rows in place (rare after the first tiles)
                    tc::mbar_wait(&bar_pv, (uint32_t)((j - 1) & 1));
                    tc::fence_after_sync();
                    const float f = raise ? ex2_approx(m_used - mt) : 1.0f;      // m_used = -inf -> 0 (O_g is still all zero)
                    uint32_t o[32], o32;
                    tc::tmem_ld32_nowait(tO, o);
                    tc::tmem_ld1_nowait(tO + 32, o32);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
                    tc::tmem_st32(tO, o);
                    tc::tmem_st1(tO + 32, __float_as_uint(__uint_as_float(o32) * f));
                    if (DROP) l_reg *= f;
                }
                if (raise) m_used = mt;
            }
            const float nm = (m_used == -INFINITY) ? 0.f : -m_used;             // all keys so far padding: exp2(-inf) = 0
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                // pairs picked evenly: pair i is emulated when floor((i+1) k / 16) > floor(i k / 16), k = EMU / 2
                const bool emu = ((c / 2 + 1) * (EMU / 2)) / 16 > ((c / 2) * (EMU / 2)) / 16;
                float p0 = emu ? ex2_poly(sv[c] + nm) : ex2_approx(sv[c] + nm);
                float p1 = emu ? ex2_poly(sv[c + 1] + nm) : ex2_approx(sv[c + 1] + nm);
                if (DROP) {
                    l_reg += p0 + p1;
                    const uint32_t kk = (uint32_t)(j * 128 + g * 32 + c);
                    p0 = drop_keep(dc, rowkey, kk) ? p0 * dc.inv_keep : 0.f;
                    p1 = drop_keep(dc, rowkey, kk + 1) ? p1 * dc.inv_keep : 0.f;
                }
                pk[c >> 1] = tc::pack_bf16(p0, p1);
            }
            if (PREF && j + 1 < nkv) request_scores(j + 1);          // sv is dead from here on
            tc::tmem_st16(tS, pk);
            tc::tmem_wait_st();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_sm[s]);
        }
        // ---- merge the four streams of each row ----
        tc::mbar_wait(&bar_pv, (uint32_t)((nkv - 1) & 1));
        tc::fence_after_sync();
        float* mbuf = reinterpret_cast<float*>(sm + KT_B);              // [128][4] stream maxima
        float* obuf = mbuf + 128 * 4;                                   // [4][128][33] rescaled stream outputs (+ row sum)
        mbuf[row * 4 + g] = m_used;
        asm volatile("bar.sync 1, %0;\n" ::"n"(NSW * 32) : "memory");
        const float4 m4 = *reinterpret_cast<const float4*>(mbuf + row * 4);
        const float m = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));
        const float f = (m_used == -INFINITY) ? 0.f : ex2_approx(m_used - m);
        {
            uint32_t o[32], o32;
            tc::tmem_ld32_nowait(tO, o);
            tc::tmem_ld1_nowait(tO + 32, o32);
            tc::tmem_wait_ld();
            float* dst = obuf + (size_t)(g * 128 + row) * 33;
            if (m_used == -INFINITY) {                                  // the stream never saw a real key: O_g was never written
#pragma unroll
                for (int c = 0; c < 33; ++c) dst[c] = 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) dst[c] = __uint_as_float(o[c]) * f;
                dst[32] = (DROP ? l_reg : __uint_as_float(o32)) * f;
            }
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(NSW * 32) : "memory");
        if (q < S) {
            float l = 0.f, acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                const float* src = obuf + (size_t)(gg * 128 + row) * 33;
                l += src[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] += src[g * 8 + c];
            }
            const float inv = 1.0f / l;
            const size_t ooff = ((size_t)b * S + q) * (H * D) + h * D + g * 8;
            if (out_bf16) {
                uint4 pk;
                pk.x = tc::pack_bf16(acc[0] * inv, acc[1] * inv); pk.y = tc::pack_bf16(acc[2] * inv, acc[3] * inv);
                pk.z = tc::pack_bf16(acc[4] * inv, acc[5] * inv); pk.w = tc::pack_bf16(acc[6] * inv, acc[7] * inv);
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out_v) + ooff) = pk;
                if (out32) {
                    *reinterpret_cast<float4*>(out32 + ooff) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
                    *reinterpret_cast<float4*>(out32 + ooff + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
                }
            } else {
                float* o = reinterpret_cast<float*>(out_v) + ooff;
                *reinterpret_cast<float4*>(o) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
                *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
            }
            if (g == 0) lse[((size_t)b * H + h) * S + q] = m + log2f(l);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// --------------------------------------------------------------------------- backward
// CTA = 128 keys of one (batch, head), 256 threads, loop over 128-query tiles processed as two
// 64-query halves so that the TMEM footprint (S^T 64 + dP^T 64 + dV + dK + dQ columns) fits 256
// columns and TWO CTAs share an SM.  Software pipeline: one mbarrier wait + one CTA barrier per half;
// after the exp/FMA phase of a half the issuing thread queues dV, dK (and dQ on the second half) of
// THIS half and S^T, dP^T of the NEXT half back to back; Q / dO tiles are double buffered in shared
// memory and staged through registers one tile ahead.
template <int D, bool DROP>
__global__ void __launch_bounds__(256, 2)
attn_bwd_kernel(const bf16* __restrict__ Qb, const bf16* __restrict__ Kb, const bf16* __restrict__ Vb,
                const bf16* __restrict__ dOb, const float* __restrict__ lse, const float* __restrict__ Dvec,
                float* __restrict__ dQacc, float* __restrict__ dKh, float* __restrict__ dVh,
                int S, int H, int Hkv, float scale, float scale_dk, const DropCfg dc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float nlse_s[2][128];    // -lse (log2 domain); -inf for padding queries
    __shared__ __align__(16) float D_s[2][128];
    __shared__ __align__(16) uint32_t rk_s[2][128];   // DROP: per-query row keys of the tile
    constexpr int TILE_B = 128 * D * 2;
    constexpr uint32_t TM_COLS = (128 + 3 * D <= 256) ? 256 : 512;
    uint8_t* Kt = sm;
    uint8_t* Vt = sm + TILE_B;
    uint8_t* Qs = sm + 2 * TILE_B;                  // [2]
    uint8_t* dOs = sm + 4 * TILE_B;                 // [2]
    uint8_t* PTs = sm + 6 * TILE_B;                 // P^T of the current half  [128 keys x 64 q] bf16
    uint8_t* dSs = PTs + 128 * 64 * 2;              // dS^T of the whole tile   [128 keys x 128 q] bf16
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 128;
    const int kvh = h / (H / Hkv);
    const int key = k0 + row;
    const bool valid_k = key < S;
    const size_t head_off = ((size_t)(b * H + h) * S) * D;
    const size_t stat_off = ((size_t)b * H + h) * S;
    const int nq = (S + 127) / 128;
    constexpr uint32_t TM_ST = 0, TM_DPT = 64, TM_DV = 128, TM_DK = 128 + D, TM_DQ = 128 + 2 * D;
    const bf16* tile_src = (half == 0 ? Qb : dOb) + (size_t)(b * H + h) * (pad128(S) >> 7) * tile_elems(D);   // pre-tiled operands
    const float* stat_src = (half == 0 ? lse : Dvec) + stat_off;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, TM_COLS);
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    uint4 nreg[D / 8];
    float nstat = 0.f;
    {
        uint4 r[D / 8];
        load_row_tiled<D>((half == 0 ? Kb : Vb) + (size_t)(b * Hkv + kvh) * (pad128(S) >> 7) * tile_elems(D), valid_k ? key : 0, valid_k, r);
        store_row<D>(half == 0 ? Kt : Vt, row, r);
        const bool vq = row < S;                     // query tile 0 -> buffer 0
        load_row_tiled<D>(tile_src, vq ? row : 0, vq, r);
        store_row<D>(half == 0 ? Qs : dOs, row, r);
        const float st0 = vq ? stat_src[row] : (half == 0 ? INFINITY : 0.f);
        if (half == 0) nlse_s[0][row] = -st0; else D_s[0][row] = st0;
        if (DROP && half == 0) rk_s[0][row] = drop_rowkey(dc, (uint32_t)(stat_off + row));
        if (nq > 1) {                                // tile 1 -> registers
            const int qn = 128 + row;
            const bool v1 = qn < S;
            load_row_tiled<D>(tile_src, v1 ? qn : 0, v1, nreg);
            nstat = v1 ? stat_src[qn] : (half == 0 ? INFINITY : 0.f);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sK = tc::smem_u32(Kt), sV = tc::smem_u32(Vt), sQ0 = tc::smem_u32(Qs), sdO0 = tc::smem_u32(dOs);
    const uint32_t sPT = tc::smem_u32(PTs), sdS = tc::smem_u32(dSs);
    constexpr uint32_t idesc64 = tc::make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idescKM = tc::make_idesc_bf16(128, D, 0, 1);     // A K-major, B MN-major
    constexpr uint32_t idescMM = tc::make_idesc_bf16(128, D, 1, 1);     // A MN-major, B MN-major
    uint32_t ph = 0;

    constexpr uint32_t KS = tc::kstep_kmajor(128);
    const tc::Desc kK = tc::kmajor(sK, 128), kV = tc::kmajor(sV, 128), mK = tc::mnmajor(sK, 128);
    const tc::Desc kPT = tc::kmajor(sPT, 128), kdS = tc::kmajor(sdS, 128), mdS = tc::mnmajor(sdS, 128);
    const tc::Desc kQ = tc::kmajor(sQ0, 128), kdO = tc::kmajor(sdO0, 128), mQ = tc::mnmajor(sQ0, 128), mdO = tc::mnmajor(sdO0, 128);
    // queues S^T and dP^T of (tile buffer `bq`, half `hq`)
    auto issue_scores = [&](int bq, int hq) {
        const uint32_t off = bq * TILE_B + hq * 64 * 16;
#pragma unroll
        for (int s = 0; s < D / 16; ++s)
            tc::mma_bf16(tmem + TM_ST, kK.adv(s * KS).u64(), kQ.adv(off + s * KS).u64(), idesc64, s > 0);
#pragma unroll
        for (int s = 0; s < D / 16; ++s)
            tc::mma_bf16(tmem + TM_DPT, kV.adv(s * KS).u64(), kdO.adv(off + s * KS).u64(), idesc64, s > 0);
    };
    // dQ of the previous tile (rows = queries) -> global accumulation.  Each warp transposes its
    // 32-row x D/2-column block through a private shared slice so that every vector reduction
    // covers fully used 32-byte sectors (a thread's own row would touch half a sector per request).
    // The staging bytes are carved from the part of the P^T tile that only this warp writes in its
    // next exp phase (chunks half*4 .. half*4+3, rows 32*(warp&3) ..): 4 pieces of 8 rows x 64 B.
    const int lane = tid & 31;
    uint8_t* xbase = PTs + (half * 4) * (128 * 16) + ((warp & 3) * 32) * 16;
    auto xaddr = [&](int r, int c) -> float* {       // r: row 0..31 of the warp block, c: float column 0..15
        return reinterpret_cast<float*>(xbase + (r >> 3) * (128 * 16) + (r & 7) * 64) + c;
    };
    auto dq_epilogue = [&](int q0) {
        if constexpr (D != 32) {                      // wide heads: one row per thread (1 CTA/SM path anyway)
            const int qq = q0 + row;
            float t[32];
            tc::tmem_ld32(tlane + TM_DQ + half * 32, t);
            if (qq < S) {
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    atomicAdd(reinterpret_cast<float4*>(dQacc + head_off + (size_t)qq * D + half * 32 + c),
                              make_float4(t[c] * scale, t[c + 1] * scale, t[c + 2] * scale, t[c + 3] * scale));
            }
            return;
        }
        float t[16];
        tc::tmem_ld16(tlane + TM_DQ + half * 16, t);
#pragma unroll
        for (int c = 0; c < 16; c += 4)
            *reinterpret_cast<float4*>(xaddr(lane, c)) = make_float4(t[c] * scale, t[c + 1] * scale, t[c + 2] * scale, t[c + 3] * scale);
        __syncwarp();
        {
#pragma unroll
            for (int it = 0; it < 4; ++it) {          // 8 rows x 64 B per warp instruction
                const int r = it * 8 + (lane >> 2), ch = lane & 3;
                const int qq = q0 + (warp & 3) * 32 + r;
                if (qq < S) {
                    const float4 v = *reinterpret_cast<const float4*>(xaddr(r, ch * 4));
                    atomicAdd(reinterpret_cast<float4*>(dQacc + head_off + (size_t)qq * D + half * 16 + ch * 4), v);
                }
            }
        }
        __syncwarp();
    };

    if (warp == 0) {
        if (tc::elect_one()) { issue_scores(0, 0); tc::mma_commit(&mbar); }
        __syncwarp();
    }

    for (int i = 0; i < nq; ++i) {
        const int bq = i & 1;
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
            tc::mbar_wait(&mbar, ph); ph ^= 1;        // scores of (i, hq) ready; every earlier MMA has completed
            tc::fence_after_sync();
            if (hq == 0 && i > 0) dq_epilogue((i - 1) * 128);
            {
                const int c0 = half * 32;             // this thread's 32 query columns of the half
                float st[32], dp[32];
                {
                    uint32_t r0[32], r1[32];
                    tc::tmem_ld32_nowait(tlane + TM_ST + c0, r0);
                    tc::tmem_ld32_nowait(tlane + TM_DPT + c0, r1);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; ++c) { st[c] = __uint_as_float(r0[c]); dp[c] = __uint_as_float(r1[c]); }
                }
                if (!valid_k) {                       // padding keys: exp2(-inf) = 0 -> P = dS = 0
#pragma unroll
                    for (int c = 0; c < 32; ++c) st[c] = -INFINITY;
                }
                const float* nl_p = &nlse_s[bq][hq * 64 + c0];
                const float* dd_p = &D_s[bq][hq * 64 + c0];
                const uint32_t* rk_p = &rk_s[bq][hq * 64 + c0];
                const uint32_t kterm = (uint32_t)key * 0x85EBCA6Bu;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const float4 l0 = *reinterpret_cast<const float4*>(nl_p + c8 * 8);
                    const float4 l1 = *reinterpret_cast<const float4*>(nl_p + c8 * 8 + 4);
                    const float4 d0 = *reinterpret_cast<const float4*>(dd_p + c8 * 8);
                    const float4 d1 = *reinterpret_cast<const float4*>(dd_p + c8 * 8 + 4);
                    const float nl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                    float p[8], ds[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        p[c] = ex2_approx(st[c8 * 8 + c] + nl[c]);   // Q is pre-scaled: scores already in the log2 domain
                        if (DROP) {
                            const float mk = (lowbias32(rk_p[c8 * 8 + c] ^ kterm) >= dc.thresh) ? dc.inv_keep : 0.f;
                            ds[c] = p[c] * fmaf(dp[c8 * 8 + c], mk, -dd[c]);
                            p[c] *= mk;                                   // P^T tile feeds dV with the dropped probabilities
                        } else {
                            ds[c] = p[c] * (dp[c8 * 8 + c] - dd[c]);      // unscaled; `scale` is applied to dQ / dK at the end
                        }
                    }
                    uint4 o;
                    o.x = tc::pack_bf16(p[0], p[1]); o.y = tc::pack_bf16(p[2], p[3]);
                    o.z = tc::pack_bf16(p[4], p[5]); o.w = tc::pack_bf16(p[6], p[7]);
                    *reinterpret_cast<uint4*>(PTs + (half * 4 + c8) * (128 * 16) + row * 16) = o;
                    o.x = tc::pack_bf16(ds[0], ds[1]); o.y = tc::pack_bf16(ds[2], ds[3]);
                    o.z = tc::pack_bf16(ds[4], ds[5]); o.w = tc::pack_bf16(ds[6], ds[7]);
                    *reinterpret_cast<uint4*>(dSs + (hq * 8 + half * 4 + c8) * (128 * 16) + row * 16) = o;
                }
            }
            if (hq == 0 && i + 1 < nq) {              // tile i+1: registers -> the other buffer; fetch tile i+2
                store_row<D>((half == 0 ? Qs : dOs) + (bq ^ 1) * TILE_B, row, nreg);
                if (half == 0) nlse_s[bq ^ 1][row] = -nstat; else D_s[bq ^ 1][row] = nstat;
                if (DROP && half == 0) rk_s[bq ^ 1][row] = drop_rowkey(dc, (uint32_t)(stat_off + (i + 1) * 128 + row));
                if (i + 2 < nq) {
                    const int qn = (i + 2) * 128 + row;
                    const bool vq = qn < S;
                    load_row_tiled<D>(tile_src, vq ? qn : 0, vq, nreg);
                    nstat = vq ? stat_src[qn] : (half == 0 ? INFINITY : 0.f);
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    const uint32_t off = bq * TILE_B + hq * 64 * 16;
#pragma unroll
                    for (int s = 0; s < 4; ++s)   // dV[key,d] += P^T[key, 64 q] dO[64 q, d]
                        tc::mma_bf16(tmem + TM_DV, kPT.adv(s * KS).u64(), mdO.adv(off + s * tc::KSTEP_MN).u64(), idescKM,
                                     (i > 0) || (hq > 0) || (s > 0));
#pragma unroll
                    for (int s = 0; s < 4; ++s)   // dK[key,d] += dS^T[key, 64 q] Q[64 q, d]
                        tc::mma_bf16(tmem + TM_DK, kdS.adv((hq * 4 + s) * KS).u64(), mQ.adv(off + s * tc::KSTEP_MN).u64(), idescKM,
                                     (i > 0) || (hq > 0) || (s > 0));
                    if (hq == 1) {
#pragma unroll
                        for (int s = 0; s < 8; ++s)   // dQ[q,d] = dS[q, key] K[key, d] over all 128 queries of the tile
                            tc::mma_bf16(tmem + TM_DQ, mdS.adv(s * tc::KSTEP_MN).u64(), mK.adv(s * tc::KSTEP_MN).u64(), idescMM, s > 0);
                    }
                    if (hq == 0) issue_scores(bq, 1);
                    else if (i + 1 < nq) issue_scores(bq ^ 1, 0);
                    tc::mma_commit(&mbar);
                }
                __syncwarp();
            }
        }
    }
    tc::mbar_wait(&mbar, ph); ph ^= 1;
    tc::fence_after_sync();
    dq_epilogue((nq - 1) * 128);
    // ---- epilogue: dK (scaled), dV of this key tile ----
    {
        float* dk = dKh + head_off + (size_t)key * D + half * (D / 2);
        float* dv = dVh + head_off + (size_t)key * D + half * (D / 2);
        float t[D / 2];
        if constexpr (D == 32) tc::tmem_ld16(tlane + TM_DK + half * 16, t);
        else tc::tmem_ld32(tlane + TM_DK + half * 32, t);
        if (valid_k) for (int c = 0; c < D / 2; c += 4)
            *reinterpret_cast<float4*>(dk + c) = make_float4(t[c] * scale_dk, t[c + 1] * scale_dk, t[c + 2] * scale_dk, t[c + 3] * scale_dk);
        if constexpr (D == 32) tc::tmem_ld16(tlane + TM_DV + half * 16, t);
        else tc::tmem_ld32(tlane + TM_DV + half * 32, t);
        if (valid_k) for (int c = 0; c < D / 2; c += 4)
            *reinterpret_cast<float4*>(dv + c) = make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TM_COLS);
}

// --------------------------------------------------------------------------- backward, warp-specialised (head_dim 32)
// One CTA per SM owns 128 keys of one (batch, head) and walks the query tiles in 64-query steps; the step's scores live
// in one of TWO TMEM stages, so the tensor core computes the scores of step n+2 while the compute warps are busy with
// step n+1, and the accumulator products of step n (dV, dK, dQ) drain behind them.  Nothing in the loop is a CTA-wide
// barrier; the roles talk through mbarriers:
//   compute warps (16): wait scores(n) -> tcgen05.ld S'^T, dP'^T (16 columns each) -> P^T = exp2(S'^T),
//                       dS^T = P^T dP'^T -> packed bf16 pairs written back INTO the same TMEM columns (A operands of
//                       dV / dK); dS^T also to a chunk-major shared tile for dQ -> arrive(cmp[n&1])
//   MMA warp (1 lane) : wait cmp[n&1] -> dV += P^T dO, dK += dS^T Q (A from TMEM) -> scores(n+2) into the same stage ->
//                       commit(S[n&1]); (second half) dQ = dS K -> commit(acc[n&1])
//   drain warps (4)   : wait acc of a tile's second step -> tcgen05.ld dQ (128 queries x 32) -> arrive(dq) -> scaled rows
//                       to a linear 4 KB staging block -> ONE cp.reduce.async.bulk (.add.f32) per warp into global dQ
//   loader warp (1)   : one cp.async.bulk per pre-tiled Q / dO tile (10 KB each), three tiles deep, released by the
//                       accumulator barrier of the tile that used the buffer
// The operands carry the softmax statistics (see "operand layouts"): S' = K'Q'^T = scaled scores - lse and
// dP' = V'dO'^T = dP - D come out of the tensor core (contraction length 48 = 32 + one hi/lo chunk + a zero chunk), so
// an element costs one ex2, one multiply and two halves of a pack.  With dropout D is subtracted by hand (dO' carries
// zeros) because the mask multiplies dP first.
// TMEM (512 columns): NSTG = 3 score stages: S'^T at 128 s, dP'^T at 128 s + 64; dV 384; dK 416; dQ (two tiles in flight) 448 + 32 b.
// Three stages (round 2; two before): the scores of step n+3 are issued when step n's P^T / dS^T are written, so the compute warps have
// two steps of scores ahead of them and the ~500-cycle chain [last warp arrives -> dV, dK -> scores -> commit] no longer sits between
// consecutive steps (ncu r01g: 14 % of the compute warps' samples were the wait for the next scores, tensor pipe 33 %, XU 44 %).
namespace bw2 {
constexpr int D = 32;
constexpr int NLB = 3;                              // Q / dO tile buffers
constexpr int TILE_B = 128 * (D + 16) * 2;          // 12 KB: 128 rows x (32 + extra chunk + zero chunk) bf16
constexpr int TILE_G = 128 * (D + 8) * 2;           // 10 KB: what a tile occupies in global memory (no zero chunk)
constexpr int PT_B = 128 * 64 * 2;                  // 16 KB: P^T of one step
constexpr int DS_B = 128 * 128 * 2;                 // 32 KB: dS^T of one query tile
constexpr int OFF_K = 0, OFF_V = TILE_B, OFF_Q = 2 * TILE_B, OFF_DO = (2 + NLB) * TILE_B, OFF_PT = (2 + 2 * NLB) * TILE_B;
constexpr int OFF_DS = OFF_PT + 2 * PT_B, OFF_STG = OFF_DS + 2 * DS_B;
constexpr int SMEM = OFF_STG + 4 * 4096;            // 212992 B
}

// PADK: the sequence has padding keys (S % 128 != 0) -> per-element validity select; EMU: of every 32 exponentials this many
// run as the FMA-pipe polynomial (ex2_poly)
// DBG (timing experiments only, results are wrong when non-zero; NOT instantiated in the library -- profiles/r02e_attn_bwd_ablation.txt was
// measured with a scratch launcher at commit 1001e31): 1 = no dQ products, 2 = no dS^T shared-memory stores, 4 = no MUFU (multiply instead
// of ex2), 8 = no dV / dK products, 16 = no dQ drain
// PREF: the compute warps request the scores of step n+1 from TMEM as soon as the registers of step n are consumed, so the
// tcgen05.ld latency runs under the store / fence / arrive tail of step n instead of in front of step n+1's exponentials
template <int NCW, bool DROP, bool PADK = true, int EMU = 0, int NSTG = 2, int DBG = 0, bool PREF = false>
__global__ void __launch_bounds__((NCW + 6) * 32, 1)
attn_bwd2_kernel(const bf16* __restrict__ Qb, const bf16* __restrict__ Kb, const bf16* __restrict__ Vb,
                 const bf16* __restrict__ dOb, const float* __restrict__ Dvec,
                 float* __restrict__ dQacc, float* __restrict__ dKh, float* __restrict__ dVh,
                 int S, int H, int Hkv, float scale, float scale_dk, const DropCfg dc) {
    using namespace bw2;
    constexpr int CG = NCW / 4;                      // column groups of a 64-query step
    constexpr int CPT = 64 / CG;                     // score columns per compute thread
    static_assert(NCW == 16, "a thread's 16 queries must be exactly one K16 step of the TMEM A operands");
    constexpr int CQ = D / CG;                       // dK / dV columns per compute thread in the epilogue
    constexpr int W_DRAIN = NCW, W_MMA = NCW + 4, W_LOAD = NCW + 5;
    // NSTG score stages of 128 TMEM columns (S'^T | dP'^T of one 64-query step), then dV, dK and two dQ tiles: 512 columns at NSTG = 3
    constexpr uint32_t TM_DV = 128u * NSTG, TM_DK = TM_DV + 32u, TM_DQ = TM_DV + 64u;
    static_assert(NSTG == 2 || NSTG == 3, "score stages");
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar_S[NSTG], bar_acc[2], bar_cmp[NSTG], bar_load[NLB], bar_dq[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float D_s[DROP ? NLB : 1][128];       // DROP only (zero-filled for padding queries)
    __shared__ __align__(16) uint32_t rk_s[DROP ? NLB : 1][128];   // DROP only: per-query row keys
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 128;
    const int kvh = h / (H / Hkv);
    const size_t head_off = ((size_t)(b * H + h) * S) * D;
    const size_t stat_off = ((size_t)b * H + h) * S;
    const int nq = (S + 127) / 128, nsteps = 2 * nq;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&bar_acc[i], 1); tc::mbar_init(&bar_dq[i], 4); }
#pragma unroll
        for (int i = 0; i < NSTG; ++i) { tc::mbar_init(&bar_S[i], 1); tc::mbar_init(&bar_cmp[i], NCW); }
#pragma unroll
        for (int i = 0; i < NLB; ++i) tc::mbar_init(&bar_load[i], DROP ? 34 : 1);
        tc::mbar_fence_init();
    }
    if (tid < 256) {                                  // K' / V' tiles of this CTA (resident for the whole kernel)
        const int row = tid & 127, which = tid >> 7;
        const bool vk = k0 + row < S;
        uint4 r[D / 8];
        load_row_tiled<D>((which == 0 ? Kb : Vb) + (size_t)(b * Hkv + kvh) * nq * tile_elems(D), vk ? k0 + row : 0, vk, r);
        uint8_t* tile = sm + (which == 0 ? OFF_K : OFF_V);
        store_row<D>(tile, row, r);
        *reinterpret_cast<uint4*>(tile + 4 * (128 * 16) + row * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);   // (1, 1, 0...)
        *reinterpret_cast<uint4*>(tile + 5 * (128 * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    } else if (tid < 256 + 128 * 2) {                 // the zero chunk of every Q' / dO' buffer (never overwritten)
        const int row = tid & 127, which = (tid - 256) >> 7;
#pragma unroll
        for (int bf = 0; bf < NLB; ++bf)
            *reinterpret_cast<uint4*>(sm + (which == 0 ? OFF_Q : OFF_DO) + bf * TILE_B + 5 * (128 * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sbase = tc::smem_u32(sm);

    if (warp == W_LOAD) {
        // ------------------------------------------------------------------ loader (fire and forget)
        const bf16* qsrc = Qb + (size_t)(b * H + h) * nq * (TILE_G / 2);
        const bf16* dsrc = dOb + (size_t)(b * H + h) * nq * (TILE_G / 2);
        for (int t = 0; t < nq; ++t) {
            const int buf = t % NLB;
            if (t >= NLB) tc::mbar_wait(&bar_acc[1], (uint32_t)((t - NLB) & 1));   // every MMA that read this buffer has completed
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&bar_load[buf], 2 * TILE_G);
                tc::bulk_copy_g2s(sbase + OFF_Q + buf * TILE_B, qsrc + (size_t)t * (TILE_G / 2), TILE_G, &bar_load[buf]);
                tc::bulk_copy_g2s(sbase + OFF_DO + buf * TILE_B, dsrc + (size_t)t * (TILE_G / 2), TILE_G, &bar_load[buf]);
            }
            if (DROP) {                               // D and the row keys of the tile for the compute warps
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 32 + lane, q = t * 128 + r;
                    const bool ok = q < S;
                    tc::cp_async4(tc::smem_u32(&D_s[buf][r]), ok ? (const void*)(Dvec + stat_off + q) : (const void*)Dvec, ok);
                    rk_s[buf][r] = drop_rowkey(dc, (uint32_t)(stat_off + q));
                }
                tc::cp_async_mbar_arrive_noinc(&bar_load[buf]);
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_load[buf]);     // releases the row-key stores
            }
        }
        if (DROP) tc::cp_async_wait_all();
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ tensor-core issue (one lane)
        if (tc::elect_one()) {
            constexpr uint32_t idesc64 = tc::make_idesc_bf16(128, 64, 0, 0);
            constexpr uint32_t idescKM = tc::make_idesc_bf16(128, D, 0, 1);     // A K-major, B MN-major
            constexpr uint32_t idescMM = tc::make_idesc_bf16(128, D, 1, 1);     // A MN-major, B MN-major
            constexpr uint32_t KS = tc::kstep_kmajor(128);
            const tc::Desc kK = tc::kmajor(sbase + OFF_K, 128), kV = tc::kmajor(sbase + OFF_V, 128), mK = tc::mnmajor(sbase + OFF_K, 128);
            const tc::Desc kQ = tc::kmajor(sbase + OFF_Q, 128), kdO = tc::kmajor(sbase + OFF_DO, 128);
            const tc::Desc mQ = tc::mnmajor(sbase + OFF_Q, 128), mdO = tc::mnmajor(sbase + OFF_DO, 128);
            const tc::Desc mdS = tc::mnmajor(sbase + OFF_DS, 128);
            auto issue_scores = [&](int n) {          // contraction over 48 columns: 32 + the statistics chunk + a zero chunk
                const uint32_t off = ((n >> 1) % NLB) * TILE_B + (n & 1) * 64 * 16;
                const uint32_t tS = tmem + (uint32_t)(n % NSTG) * 128u;
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    tc::mma_bf16(tS, kK.adv(s * KS).u64(), kQ.adv(off + s * KS).u64(), idesc64, s > 0);
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    tc::mma_bf16(tS + 64, kV.adv(s * KS).u64(), kdO.adv(off + s * KS).u64(), idesc64, s > 0);
            };
            for (int n = 0; n < NSTG && n < nsteps; ++n) {                      // the first NSTG steps' scores
                if ((n & 1) == 0) { tc::mbar_wait(&bar_load[(n >> 1) % NLB], 0); tc::fence_after_sync(); }
                issue_scores(n); tc::mma_commit(&bar_S[n]);
            }
            for (int n = 0; n < nsteps; ++n) {
                const int i = n >> 1, hq = n & 1, s = n % NSTG;
                tc::mbar_wait(&bar_cmp[s], (uint32_t)((n / NSTG) & 1));         // P^T / dS^T of step n written, stage s drained
                tc::fence_after_sync();
                const uint32_t off = (i % NLB) * TILE_B + hq * 64 * 16;
                const uint32_t tS = tmem + (uint32_t)s * 128u;
#pragma unroll
                for (int k = 0; k < ((DBG & 8) ? 0 : 4); ++k)     // dV[key,d] += P^T[key, 16 q] dO[16 q, d]   (A = packed P^T in the stage's S' columns)
                    tc::mma_bf16_ts(tmem + TM_DV, tS + 16u * k, mdO.adv(off + k * tc::KSTEP_MN).u64(), idescKM, (n | k) != 0);
#pragma unroll
                for (int k = 0; k < ((DBG & 8) ? 0 : 4); ++k)     // dK[key,d] += dS^T[key, 16 q] Q'[16 q, d]  (A = packed dS^T in the stage's dP' columns)
                    tc::mma_bf16_ts(tmem + TM_DK, tS + 64u + 16u * k, mQ.adv(off + k * tc::KSTEP_MN).u64(), idescKM, (n | k) != 0);
                if (n + NSTG < nsteps) {          // overwrites the stage the two products above read: ordered behind them
                    const int n2 = n + NSTG, i2 = n2 >> 1;
                    if ((n2 & 1) == 0) { tc::mbar_wait(&bar_load[i2 % NLB], (uint32_t)((i2 / NLB) & 1)); tc::fence_after_sync(); }
                    issue_scores(n2);
                    tc::mma_commit(&bar_S[s]);
                }
                if (hq == 1) {
                    if (i >= 2) { tc::mbar_wait(&bar_dq[i & 1], (uint32_t)(((i - 2) >> 1) & 1)); tc::fence_after_sync(); }
#pragma unroll
                    for (int k = 0; k < ((DBG & 1) ? 0 : 8); ++k)   // dQ[q,d] = dS[q, key] K[key, d] over the 128 queries of the tile
                        tc::mma_bf16(tmem + TM_DQ + (uint32_t)(i & 1) * 32u, mdS.adv((i & 1) * DS_B + k * tc::KSTEP_MN).u64(),
                                     mK.adv(k * tc::KSTEP_MN).u64(), idescMM, k > 0);
                }
                tc::mma_commit(&bar_acc[hq]);
            }
        }
        __syncwarp();
    } else if (warp >= W_DRAIN) {
        // ------------------------------------------------------------------ dQ drain: TMEM -> staging -> bulk reduce-add
        const int quarter = warp & 3;
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        float* stg = reinterpret_cast<float*>(sm + OFF_STG + quarter * 4096);
        for (int i = 0; i < nq; ++i) {
            tc::mbar_wait(&bar_acc[1], (uint32_t)(i & 1));
            tc::fence_after_sync();
            float t[32];
            tc::tmem_ld32(tlane + TM_DQ + (uint32_t)(i & 1) * 32u, t);
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&bar_dq[i & 1]); tc::bulk_wait_read0(); }   // staging block free again
            __syncwarp();
            if (DBG & 16) continue;
            {
                // Row `lane` is 128 contiguous bytes (the bulk reduce needs the block linear), so a straight float4 store
                // would put all lanes of a quarter-warp on the same four banks (8-way conflict, ~1000 smem cycles per
                // tile that also stall the tensor core's operand reads).  Lane l instead writes its chunks in the order
                // (j + l) mod 8: the register array is rotated by l with three conditional-move stages.
                float4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = make_float4(t[4 * c] * scale, t[4 * c + 1] * scale, t[4 * c + 2] * scale, t[4 * c + 3] * scale);
                const int rot = lane & 7;
#pragma unroll
                for (int sh = 1; sh < 8; sh <<= 1) {
                    const bool on = (rot & sh) != 0;
                    float4 w[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) w[c] = on ? v[(c + sh) & 7] : v[c];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = w[c];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((j + rot) & 7) * 4) = v[j];
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                const int q0 = i * 128 + quarter * 32;
                const int rows = min(32, S - q0);
                if (rows > 0) tc::bulk_reduce_add_f32(dQacc + head_off + (size_t)q0 * D, stg, (uint32_t)rows * (D * 4));
                tc::bulk_commit();
            }
        }
        if (lane == 0) tc::bulk_wait0();
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ compute warps
        const int quarter = warp & 3, cg = warp >> 2;
        const int row = quarter * 32 + lane, key = k0 + row;
        const bool valid_k = PADK ? key < S : true;
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        const int c0 = cg * CPT;
        const uint32_t kterm = (uint32_t)key * 0x85EBCA6Bu;
        uint8_t* const ds0 = sm + OFF_DS + (cg * (CPT / 8)) * (128 * 16) + row * 16;
        uint32_t r0[CPT], r1[CPT];
        auto request_scores = [&](int n) {            // wait for the scores of step n, then issue (not await) their TMEM loads
            const int s = n % NSTG;
            tc::mbar_wait(&bar_S[s], (uint32_t)((n / NSTG) & 1));
            tc::fence_after_sync();
            if constexpr (CPT == 16) {
                tc::tmem_ld16_nowait(tlane + (uint32_t)s * 128u + c0, r0);
                tc::tmem_ld16_nowait(tlane + (uint32_t)s * 128u + 64u + c0, r1);
            } else {
                tc::tmem_ld32_nowait(tlane + (uint32_t)s * 128u + c0, r0);
                tc::tmem_ld32_nowait(tlane + (uint32_t)s * 128u + 64u + c0, r1);
            }
        };
        if (PREF) request_scores(0);
        for (int n = 0; n < nsteps; ++n) {
            const int i = n >> 1, hq = n & 1, s = n % NSTG;
            if (DROP && hq == 0) tc::mbar_wait(&bar_load[i % NLB], (uint32_t)((i / NLB) & 1));   // D / row keys of the tile
            if (!PREF) request_scores(n);
            tc::tmem_wait_ld();
            uint4 pk[CPT / 8], dk[CPT / 8];
#pragma unroll
            for (int c8 = 0; c8 < CPT / 8; ++c8) {
                float p[8], ds[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int cc = c8 * 8 + c;
                    const bool emu = ((cc + 1) * EMU) / 32 > (cc * EMU) / 32;
                    const float e = (DBG & 4) ? __uint_as_float(r0[cc]) * 0.001f : (emu ? ex2_poly(__uint_as_float(r0[cc])) : ex2_approx(__uint_as_float(r0[cc])));
                    p[c] = valid_k ? e : 0.f;                              // padding keys: P = dS = 0
                    if (DROP) {
                        const int qi = hq * 64 + c0 + cc;
                        const float mk = (lowbias32(rk_s[i % NLB][qi] ^ kterm) >= dc.thresh) ? dc.inv_keep : 0.f;
                        ds[c] = p[c] * fmaf(__uint_as_float(r1[cc]), mk, -D_s[i % NLB][qi]);
                        p[c] *= mk;                                       // P^T feeds dV with the dropped probabilities
                    } else {
                        ds[c] = p[c] * __uint_as_float(r1[cc]);           // unscaled; `scale` is applied to dQ / dK at the end
                    }
                }
                pk[c8].x = tc::pack_bf16(p[0], p[1]); pk[c8].y = tc::pack_bf16(p[2], p[3]);
                pk[c8].z = tc::pack_bf16(p[4], p[5]); pk[c8].w = tc::pack_bf16(p[6], p[7]);
                dk[c8].x = tc::pack_bf16(ds[0], ds[1]); dk[c8].y = tc::pack_bf16(ds[2], ds[3]);
                dk[c8].z = tc::pack_bf16(ds[4], ds[5]); dk[c8].w = tc::pack_bf16(ds[6], ds[7]);
            }
            if (PREF && n + 1 < nsteps) request_scores(n + 1);      // r0 / r1 are dead from here on
            // P^T and dS^T go back INTO the TMEM columns their scores came from (this thread owns them) as packed bf16
            // pairs: they are the A operands of dV / dK.  dS^T also goes to shared memory for dQ = dS K (transposed use).
            {
                uint32_t a[CPT / 2];
#pragma unroll
                for (int c8 = 0; c8 < CPT / 8; ++c8) { a[c8 * 4] = pk[c8].x; a[c8 * 4 + 1] = pk[c8].y; a[c8 * 4 + 2] = pk[c8].z; a[c8 * 4 + 3] = pk[c8].w; }
                tc::tmem_st8(tlane + (uint32_t)s * 128u + c0, a);
#pragma unroll
                for (int c8 = 0; c8 < CPT / 8; ++c8) { a[c8 * 4] = dk[c8].x; a[c8 * 4 + 1] = dk[c8].y; a[c8 * 4 + 2] = dk[c8].z; a[c8 * 4 + 3] = dk[c8].w; }
                tc::tmem_st8(tlane + (uint32_t)s * 128u + 64u + c0, a);
            }
            if (n >= 2) tc::mbar_wait(&bar_acc[hq], (uint32_t)(((n >> 1) - 1) & 1));  // dQ of tile i-2 has read this dS^T buffer
            uint8_t* dst = ds0 + (i & 1) * DS_B + hq * 8 * (128 * 16);
#pragma unroll
            if (!(DBG & 2)) {
#pragma unroll
                for (int c8 = 0; c8 < CPT / 8; ++c8) *reinterpret_cast<uint4*>(dst + c8 * (128 * 16)) = dk[c8];
            }
            tc::tmem_wait_st();
            if (!(DBG & 2)) tc::fence_async_smem();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_cmp[s]);
        }
        // ---- epilogue: dK (scaled), dV of this key tile ----
        tc::mbar_wait(&bar_acc[1], (uint32_t)(((nsteps - 1) >> 1) & 1));
        tc::fence_after_sync();
        if constexpr (CQ == 8) {
            float t[8];
            tc::tmem_ld8(tlane + TM_DK + cg * 8, t);
            if (valid_k) {
                float* dkp = dKh + head_off + (size_t)key * D + cg * 8;
                *reinterpret_cast<float4*>(dkp) = make_float4(t[0] * scale_dk, t[1] * scale_dk, t[2] * scale_dk, t[3] * scale_dk);
                *reinterpret_cast<float4*>(dkp + 4) = make_float4(t[4] * scale_dk, t[5] * scale_dk, t[6] * scale_dk, t[7] * scale_dk);
            }
            tc::tmem_ld8(tlane + TM_DV + cg * 8, t);
            if (valid_k) {
                float* dvp = dVh + head_off + (size_t)key * D + cg * 8;
                *reinterpret_cast<float4*>(dvp) = make_float4(t[0], t[1], t[2], t[3]);
                *reinterpret_cast<float4*>(dvp + 4) = make_float4(t[4], t[5], t[6], t[7]);
            }
        } else {
            float t[16];
            tc::tmem_ld16(tlane + TM_DK + cg * 16, t);
            if (valid_k) {
                float* dkp = dKh + head_off + (size_t)key * D + cg * 16;
#pragma unroll
                for (int c = 0; c < 16; c += 4)
                    *reinterpret_cast<float4*>(dkp + c) = make_float4(t[c] * scale_dk, t[c + 1] * scale_dk, t[c + 2] * scale_dk, t[c + 3] * scale_dk);
            }
            tc::tmem_ld16(tlane + TM_DV + cg * 16, t);
            if (valid_k) {
                float* dvp = dVh + head_off + (size_t)key * D + cg * 16;
#pragma unroll
                for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(dvp + c) = make_float4(t[c], t[c + 1], t[c + 2], t[c + 3]);
            }
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
    const size_t qe = (size_t)B * H * S * d, ke = (size_t)B * Hkv * pad128(S) * (d + 8), qp = (size_t)B * H * pad128(S) * (d + 8);
    return align_up(qp * 2) * 2 + align_up(ke * 2) * 2 + align_up((size_t)B * H * S * 4) + 3 * align_up(qe * 4) + 1024;
}
static bool attn_carve(AttnWs& w, void* ws, size_t bytes, int64_t B, int64_t S, int H, int Hkv, int d) {
    Arena ar(ws, bytes);
    const size_t qe = (size_t)B * H * S * d, ke = (size_t)B * Hkv * pad128(S) * (d + 8), qp = (size_t)B * H * pad128(S) * (d + 8);
    w.Qb = ar.take<bf16>(qp); w.Kb = ar.take<bf16>(ke); w.Vb = ar.take<bf16>(ke); w.dOb = ar.take<bf16>(qp);
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
    attn_prep_kernel<<<nb256(B * pad128(S) * H * (d / 8)), 256, 0, st>>>(q, w.Qb, B, S, H, d, freqs, 1, (1.0f / sqrtf((float)d)) * 1.4426950408889634f, 0);
    GAOT_LAUNCH_CHECK();
    attn_prep_kernel<<<nb256(B * pad128(S) * Hkv * (d / 8)), 256, 0, st>>>(k, w.Kb, B, S, Hkv, d, freqs, 1, 1.0f, 1);
    GAOT_LAUNCH_CHECK();
    attn_prep_kernel<<<nb256(B * pad128(S) * Hkv * (d / 8)), 256, 0, st>>>(v, w.Vb, B, S, Hkv, d, nullptr, 1, 1.0f, 2);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // namespace gaot

using namespace gaot;

extern "C" {

size_t gaot_attn_workspace_bytes(int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d) {
    return attn_ws_bytes(B, S, H, Hkv, d);
}

static DropCfg make_drop(float p, uint64_t seed) {
    DropCfg dc;
    double t = (double)p * 4294967296.0;
    dc.thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    dc.seed_lo = (uint32_t)seed; dc.seed_hi = (uint32_t)(seed >> 32);
    dc.inv_keep = p < 1.0f ? 1.0f / (1.0f - p) : 0.f;
    return dc;
}

static int attn_launch_fwd(const AttnWs& w, int64_t B, int64_t S, int H, int Hkv, int d, float dropout_p, uint64_t seed,
                           void* out, int out_bf16, float* out32, float* lse, cudaStream_t st) {
    GAOT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "attn: dropout_p must be in [0,1)");
    GAOT_CHECK_ARG((int64_t)B * H * S < ((int64_t)1 << 32), "attn: B*H*S too large for the dropout counter");
    const DropCfg dc = make_drop(dropout_p, seed);
    const bool drop = dropout_p > 0.f;
    dim3 grid((unsigned)((S + 127) / 128), (unsigned)H, (unsigned)B);
    GAOT_TIME_KERNEL("attn_fwd", st, 4.0 * (double)B * H * (double)S * (double)S * d);
    const char* dbg_env = getenv("GAOT_ATTN_DEBUG");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    if (d == 32 && !(dbg & 64)) {                    // warp-specialised kernel (GAOT_ATTN_DEBUG bit 6 selects the older one)
        // GAOT_ATTN_EMU = 0 / 4 / 8 / 12 of every 32 exponentials on the FMA pipe (default: the measured best)
        static const int emu = getenv("GAOT_ATTN_EMU") ? atoi(getenv("GAOT_ATTN_EMU")) : 4;   // measured at S = 16384: 0.688 / 0.670 / 0.683 / 0.713 ms for 0 / 4 / 8 / 12
#define GAOT_FWD2(DR, EM)                                                                                              \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<DR, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw2::SMEM)); \
         attn_fwd2_kernel<DR, EM><<<grid, fw2::THREADS, fw2::SMEM, st>>>(w.Qb, w.Kb, w.Vb, out, out_bf16, out32, lse, (int)S, H, Hkv, dc); } while (0)
        // score prefetch measured SLOWER in the forward (0.854 vs 0.659 ms at S = 16384, r02f): kept for A/B only
        static const int fpref = getenv("GAOT_ATTN_FWD_PREF") ? atoi(getenv("GAOT_ATTN_FWD_PREF")) : 0;
#define GAOT_FWD2P(EM)                                                                                                 \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<false, EM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fw2::SMEM)); \
         attn_fwd2_kernel<false, EM, true><<<grid, fw2::THREADS, fw2::SMEM, st>>>(w.Qb, w.Kb, w.Vb, out, out_bf16, out32, lse, (int)S, H, Hkv, dc); } while (0)
        if (drop) { if (emu >= 8) GAOT_FWD2(true, 8); else GAOT_FWD2(true, 0); }          // the dropout variant keeps all-MUFU unless asked
        else if (fpref) { if (emu >= 8) GAOT_FWD2P(8); else if (emu >= 4) GAOT_FWD2P(4); else GAOT_FWD2P(0); }
        else if (emu >= 12) GAOT_FWD2(false, 12);
        else if (emu >= 8) GAOT_FWD2(false, 8);
        else if (emu >= 4) GAOT_FWD2(false, 4);
        else GAOT_FWD2(false, 0);
#undef GAOT_FWD2P
#undef GAOT_FWD2
        GAOT_LAUNCH_CHECK();
        return GAOT_OK;
    }
#define GAOT_FWD_LAUNCH(DD, DR, SM)                                                                                   \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DD, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM))); \
         attn_fwd_kernel<DD, DR><<<grid, 128, (SM), st>>>(w.Qb, w.Kb, w.Vb, out, out_bf16, out32, lse, (int)S, H, Hkv, dc); } while (0)
    if (d == 32) {
        const size_t smem = 3 * 128 * 32 * 2 + 2 * 128 * 48 * 2 + 128 * 128 * 2;
        if (drop) GAOT_FWD_LAUNCH(32, true, smem); else GAOT_FWD_LAUNCH(32, false, smem);
    } else {
        const size_t smem = 3 * 128 * 64 * 2 + 2 * 128 * 80 * 2 + 128 * 128 * 2;
        if (drop) GAOT_FWD_LAUNCH(64, true, smem); else GAOT_FWD_LAUNCH(64, false, smem);
    }
#undef GAOT_FWD_LAUNCH
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_attn_forward(const float* q, const float* k, const float* v, int64_t B, int64_t S, int32_t H,
                      int32_t Hkv, int32_t d, const float* rope_freqs, float dropout_p, uint64_t seed,
                      void* ws, size_t ws_bytes, float* out, float* lse, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = attn_check(B, S, H, Hkv, d);
    if (rc) return rc;
    AttnWs w;
    if (!attn_carve(w, ws, ws_bytes, B, S, H, Hkv, d)) { set_error("attn_forward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    rc = attn_prep_all(q, k, v, w, B, S, H, Hkv, d, rope_freqs, st);
    if (rc) return rc;
    return attn_launch_fwd(w, B, S, H, Hkv, d, dropout_p, seed, out, 0, nullptr, lse, st);
}

// core backward on prepared operands: w.Qb/Kb/Vb/dOb/Dvec filled -> w.dQacc / dKh / dVh (fp32, per head)
static int attn_launch_bwd(const AttnWs& w, const float* lse, int64_t B, int64_t S, int H, int Hkv, int d,
                           float dropout_p, uint64_t seed, cudaStream_t st) {
    GAOT_CUDA(cudaMemsetAsync(w.dQacc, 0, (size_t)B * H * S * d * sizeof(float), st));
    // Q is stored pre-multiplied by scale * log2(e): dK = scale dS^T Q = ln(2) dS^T Q'
    const float scale = 1.0f / sqrtf((float)d), scale_dk = 0.6931471805599453f;
    const char* dbg_env = getenv("GAOT_ATTN_DEBUG");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    dim3 grid((unsigned)((S + 127) / 128), (unsigned)H, (unsigned)B);
    GAOT_TIME_KERNEL("attn_bwd", st, 10.0 * (double)B * H * (double)S * (double)S * d);
    const DropCfg dc = make_drop(dropout_p, seed);
    const bool drop = dropout_p > 0.f;
    if (d == 32 && !(dbg & 128)) {                   // warp-specialised kernel (GAOT_ATTN_DEBUG bit 7 selects the older one)
#define GAOT_BWD2_LAUNCH(NCW, DR)                                                                                      \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<NCW, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bw2::SMEM)); \
         attn_bwd2_kernel<NCW, DR><<<grid, (NCW + 6) * 32, bw2::SMEM, st>>>(w.Qb, w.Kb, w.Vb, w.dOb, w.Dvec, w.dQacc, w.dKh, w.dVh, \
                                                                           (int)S, H, Hkv, scale, scale_dk, dc); } while (0)
        // GAOT_ATTN_BWD_EMU = 0 / 4 / 8 (of every 32 exponentials on the FMA pipe); sequences without padding keys skip the
        // per-element validity select
        static const int bemu = getenv("GAOT_ATTN_BWD_EMU") ? atoi(getenv("GAOT_ATTN_BWD_EMU")) : 0;
        // measured at S = 16384 (profiles/r02f_attn_variants.txt): 2 stages 1.164 ms; 3 stages 1.188; 2 stages + score prefetch 1.429;
        // 3 stages + score prefetch 1.109 (default).  GAOT_ATTN_BWD_STAGES / GAOT_ATTN_BWD_PREF pick the others for A/B timing.
        static const int bstg = getenv("GAOT_ATTN_BWD_STAGES") ? atoi(getenv("GAOT_ATTN_BWD_STAGES")) : 3;
        static const int bpref = getenv("GAOT_ATTN_BWD_PREF") ? atoi(getenv("GAOT_ATTN_BWD_PREF")) : (bstg == 3 ? 1 : 0);
#define GAOT_BWD2_LAUNCH6(NS)                                                                                          \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<16, false, false, 0, NS, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bw2::SMEM)); \
         attn_bwd2_kernel<16, false, false, 0, NS, 0, true><<<grid, (16 + 6) * 32, bw2::SMEM, st>>>(w.Qb, w.Kb, w.Vb, w.dOb, w.Dvec, w.dQacc, w.dKh, w.dVh, \
                                                                                                   (int)S, H, Hkv, scale, scale_dk, dc); } while (0)
#define GAOT_BWD2_LAUNCH4(PK, EM, NS)                                                                                  \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<16, false, PK, EM, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bw2::SMEM)); \
         attn_bwd2_kernel<16, false, PK, EM, NS><<<grid, (16 + 6) * 32, bw2::SMEM, st>>>(w.Qb, w.Kb, w.Vb, w.dOb, w.Dvec, w.dQacc, w.dKh, w.dVh, \
                                                                                        (int)S, H, Hkv, scale, scale_dk, dc); } while (0)
        if (drop) GAOT_BWD2_LAUNCH(16, true);
        else if (S % 128 != 0) GAOT_BWD2_LAUNCH(16, false);
        else if (bpref) { if (bstg == 3) GAOT_BWD2_LAUNCH6(3); else GAOT_BWD2_LAUNCH6(2); }
        else if (bstg == 3) GAOT_BWD2_LAUNCH4(false, 0, 3);
        else if (bemu >= 8) GAOT_BWD2_LAUNCH4(false, 8, 2);
        else if (bemu >= 4) GAOT_BWD2_LAUNCH4(false, 4, 2);
        else GAOT_BWD2_LAUNCH4(false, 0, 2);
#undef GAOT_BWD2_LAUNCH6
#undef GAOT_BWD2_LAUNCH4
#undef GAOT_BWD2_LAUNCH
        GAOT_LAUNCH_CHECK();
        return GAOT_OK;
    }
#define GAOT_BWD_LAUNCH(DD, DR, SM)                                                                                   \
    do { GAOT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DD, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM))); \
         attn_bwd_kernel<DD, DR><<<grid, 256, (SM), st>>>(w.Qb, w.Kb, w.Vb, w.dOb, lse, w.Dvec, w.dQacc, w.dKh, w.dVh,    \
                                                         (int)S, H, Hkv, scale, scale_dk, dc); } while (0)
    if (d == 32) {
        const size_t smem = 6 * 128 * 32 * 2 + 128 * 64 * 2 + 128 * 128 * 2;
        if (drop) GAOT_BWD_LAUNCH(32, true, smem); else GAOT_BWD_LAUNCH(32, false, smem);
    } else {
        const size_t smem = 6 * 128 * 64 * 2 + 128 * 64 * 2 + 128 * 128 * 2;
        if (drop) GAOT_BWD_LAUNCH(64, true, smem); else GAOT_BWD_LAUNCH(64, false, smem);
    }
#undef GAOT_BWD_LAUNCH
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_attn_backward(const float* q, const float* k, const float* v, const float* out, const float* d_out,
                       const float* lse, int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d,
                       const float* rope_freqs, float dropout_p, uint64_t seed, void* ws, size_t ws_bytes,
                       float* dq, float* dk, float* dv, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = attn_check(B, S, H, Hkv, d);
    if (rc) return rc;
    AttnWs w;
    if (!attn_carve(w, ws, ws_bytes, B, S, H, Hkv, d)) { set_error("attn_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    rc = attn_prep_all(q, k, v, w, B, S, H, Hkv, d, rope_freqs, st);
    if (rc) return rc;
    attn_bwd_prep_kernel<<<nb256(B * pad128(S) * H), 256, 0, st>>>(d_out, out, w.dOb, w.Dvec, w.Qb, lse, dropout_p > 0.f ? 0 : 1, B, S, H, d);
    GAOT_LAUNCH_CHECK();
    rc = attn_launch_bwd(w, lse, B, S, H, Hkv, d, dropout_p, seed, st);
    if (rc) return rc;
    attn_bwd_post_kernel<<<nb256(B * S * H * (d / 8)), 256, 0, st>>>(w.dQacc, dq, B, S, H, H, d, rope_freqs);
    GAOT_LAUNCH_CHECK();
    attn_bwd_post_kernel<<<nb256(B * S * Hkv * (d / 8)), 256, 0, st>>>(w.dKh, dk, B, S, H, Hkv, d, rope_freqs);
    GAOT_LAUNCH_CHECK();
    attn_bwd_post_kernel<<<nb256(B * S * Hkv * (d / 8)), 256, 0, st>>>(w.dVh, dv, B, S, H, Hkv, d, nullptr);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

// ---- fused-block flavour: operands stay in the kernels' own bf16 per-head layout between forward and backward ----
size_t gaot_attn_packed_bytes(int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d) {
    return align_up((size_t)B * H * pad128(S) * (d + 8) * 2) + 2 * align_up((size_t)B * Hkv * pad128(S) * (d + 8) * 2);
}

static void attn_packed_ptrs(AttnWs& w, void* packed, int64_t B, int64_t S, int H, int Hkv, int d) {
    char* p = (char*)packed;
    w.Qb = (bf16*)p; p += align_up((size_t)B * H * pad128(S) * (d + 8) * 2);
    w.Kb = (bf16*)p; p += align_up((size_t)B * Hkv * pad128(S) * (d + 8) * 2);
    w.Vb = (bf16*)p;
}

int gaot_attn_fused_forward(const void* qkv, int64_t ld, int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d,
                            const float* rope_freqs, float dropout_p, uint64_t seed,
                            void* packed, void* out, float* out_f32, float* lse, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = attn_check(B, S, H, Hkv, d);
    if (rc) return rc;
    GAOT_CHECK_ARG(qkv && packed && out && out_f32 && lse && ld % 8 == 0, "attn_fused_forward: bad pointer / ld");
    AttnWs w{};
    attn_packed_ptrs(w, packed, B, S, H, Hkv, d);
    attn_pack_qkv_kernel<<<nb256(B * pad128(S) * (H + 2 * Hkv) * (d / 8)), 256, 0, st>>>((const bf16*)qkv, ld, w.Qb, w.Kb, w.Vb, B, S, H, Hkv, d, rope_freqs, (1.0f / sqrtf((float)d)) * 1.4426950408889634f);
    GAOT_LAUNCH_CHECK();
    return attn_launch_fwd(w, B, S, H, Hkv, d, dropout_p, seed, out, 1, out_f32, lse, st);
}

size_t gaot_attn_fused_backward_workspace_bytes(int64_t B, int64_t S, int32_t H, int32_t d) {
    const size_t qe = (size_t)B * H * S * d, qp = (size_t)B * H * pad128(S) * (d + 8);
    return align_up(qp * 2) + align_up((size_t)B * H * S * 4) + 3 * align_up(qe * 4) + 1024;
}

int gaot_attn_fused_backward(const void* packed, const float* out, const void* d_out, const float* lse,
                             int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d, const float* rope_freqs,
                             float dropout_p, uint64_t seed, void* ws, size_t ws_bytes,
                             void* d_qkv, int64_t ld, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = attn_check(B, S, H, Hkv, d);
    if (rc) return rc;
    GAOT_CHECK_ARG(packed && out && d_out && lse && d_qkv && ld % 8 == 0, "attn_fused_backward: bad pointer / ld");
    AttnWs w{};
    attn_packed_ptrs(w, const_cast<void*>(packed), B, S, H, Hkv, d);
    Arena ar(ws, ws_bytes);
    const size_t qe = (size_t)B * H * S * d, qp = (size_t)B * H * pad128(S) * (d + 8);
    w.dOb = ar.take<bf16>(qp); w.Dvec = ar.take<float>((size_t)B * H * S);
    w.dQacc = ar.take<float>(qe); w.dKh = ar.take<float>(qe); w.dVh = ar.take<float>(qe);
    if (!ar.ok()) { set_error("attn_fused_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    attn_bwd_prep_bf16_kernel<<<nb256(B * pad128(S) * H), 256, 0, st>>>((const bf16*)d_out, out, w.dOb, w.Dvec, w.Qb, lse, dropout_p > 0.f ? 0 : 1, B, S, H, d);
    GAOT_LAUNCH_CHECK();
    rc = attn_launch_bwd(w, lse, B, S, H, Hkv, d, dropout_p, seed, st);
    if (rc) return rc;
    attn_bwd_post_qkv_kernel<<<nb256(B * S * (H + 2 * Hkv) * (d / 8)), 256, 0, st>>>(w.dQacc, w.dKh, w.dVh, (bf16*)d_qkv, ld, B, S, H, Hkv, d, rope_freqs);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // extern "C"
