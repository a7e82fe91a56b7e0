// gno_fwd.cu -- fused GNO forward (FP32 CUDA-core path):
//   gather(y_pos[src], x_pos[qry], f_y[src]) -> per-edge kernel MLP (erf-GELU) -> (* f_y[src])
//   -> CSR-ordered segmented mean, one kernel, no [E,*] intermediates in HBM.
// Replaces reference src/model/layers/integral_transform.py:114-171 (3 index gathers, cat,
// 4 cuBLAS SGEMMs, 3 GELUs, mul, torch_scatter mean = ~20 kernels, ~4.5 KB/edge of HBM traffic).
//
// Tile = 128 consecutive CSR edges per CTA iteration (persistent grid, weights resident in
// shared memory, transposed so a thread's 4 output weights are one LDS.128).  Activations are
// kept feature-major [width][128] in shared memory; each thread owns an 8-edge x 4-output
// register tile (32 FMA per 3 LDS.128).  Source feature rows are staged with cp.async while the
// MLP runs.  The reduction walks the tile's query segments (edges are query-sorted): one warp
// per segment, lanes = channels, fixed order -> deterministic, no atomics.  Segments that cross
// a tile boundary are finished by a tiny fix-up kernel in tile order.
#include "gno_common.cuh"

namespace gaot {

constexpr int FTE = 128;
constexpr int FTHREADS = 256;
constexpr int ACT_ROWS = 72;

struct FwdSmemLayout {
    int wT_off[GNO_MAX_LAYERS];   // float offsets
    int b_off[GNO_MAX_LAYERS];
    int np[GNO_MAX_LAYERS];       // padded output width of each layer
    int actA, actB, fsm, ints;
    int total_floats;
};

static FwdSmemLayout fwd_layout(const GnoArgs& a) {
    FwdSmemLayout L;
    int off = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        L.np[l] = (a.dims[l + 1] + 3) / 4 * 4;
        L.wT_off[l] = off; off += a.dims[l] * L.np[l];
    }
    for (int l = 0; l < a.n_layers; ++l) { L.b_off[l] = off; off += L.np[l]; }
    off = (off + 3) / 4 * 4;
    L.actA = off; off += ACT_ROWS * FTE;
    L.actB = off; off += ACT_ROWS * FTE;
    L.fsm = off; off += (a.f_y ? FTE * a.c_f : 0);
    L.ints = off; off += 3 * FTE + 16;
    L.total_floats = off;
    return L;
}

__global__ void __launch_bounds__(FTHREADS, 1)
gno_fwd_fp32_kernel(const GnoArgs a, const FwdSmemLayout L, float* __restrict__ out,
                    float* __restrict__ head_partial) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* actA = smem + L.actA;
    float* actB = smem + L.actB;
    float* fsm = smem + L.fsm;
    int* s_src = reinterpret_cast<int*>(smem + L.ints);
    int* s_qry = s_src + FTE;
    int* seg_first = s_qry + FTE;          // [FTE + 1]
    int* s_misc = seg_first + FTE + 1;     // [0..3] warp head ballots, [4] nseg

    // ---- stage weights (transposed, zero padded) and biases once per CTA ----
    for (int l = 0; l < a.n_layers; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1], NP = L.np[l];
        float* wT = smem + L.wT_off[l];
        const float* W = a.params + a.w_off[l];
        for (int idx = tid; idx < K * NP; idx += FTHREADS) {
            const int i = idx / NP, j = idx - i * NP;
            wT[idx] = j < N ? W[j * K + i] : 0.0f;
        }
        float* bs = smem + L.b_off[l];
        for (int j = tid; j < NP; j += FTHREADS) bs[j] = j < N ? a.params[a.b_off[l] + j] : 0.0f;
    }
    __syncthreads();

    const int Cout = a.dims[a.n_layers];
    const int CP = Cout + 4;                       // padded row of the edge-major value tile
    const bool use_f_mul = (a.transform == 0 || a.transform == 1);
    const bool f_in_mlp = (a.transform == 1 || a.transform == 2);
    const int cf4 = a.c_f >> 2;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int e0 = tile * FTE;
        const int ne = min(FTE, a.E - e0);
        if (tid < FTE) {
            const bool valid = tid < ne;
            s_src[tid] = valid ? a.csr_src[e0 + tid] : 0;
            s_qry[tid] = valid ? a.csr_qry[e0 + tid] : -1;
        }
        __syncthreads();
        // ---- async stage of the source feature rows ----
        if (a.f_y) {
            for (int idx = tid; idx < FTE * cf4; idx += FTHREADS) {
                const int e = idx / cf4, ch = idx - e * cf4;
                if (e < ne) cp_async16(fsm + e * a.c_f + ch * 4, a.f_y + (size_t)s_src[e] * a.c_f + ch * 4);
                else *reinterpret_cast<float4*>(fsm + e * a.c_f + ch * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            cp_async_commit();
        }
        // ---- coordinates -> first-layer input rows 0..5, segment heads ----
        if (tid < FTE) {
            const bool valid = tid < ne;
            const float* p = a.y_pos + (size_t)s_src[tid] * 3;
            actA[0 * FTE + tid] = valid ? p[0] : 0.f;
            actA[1 * FTE + tid] = valid ? p[1] : 0.f;
            actA[2 * FTE + tid] = valid ? p[2] : 0.f;
            const bool head = valid && (tid == 0 || s_qry[tid - 1] != s_qry[tid]);
            const unsigned b = __ballot_sync(0xffffffffu, head);
            if (lane == 0) s_misc[warp] = (int)b;
        } else {
            const int e = tid - FTE;
            const bool valid = e < ne;
            const float* p = a.x_pos + (size_t)(valid ? s_qry[e] : 0) * 3;
            actA[3 * FTE + e] = valid ? p[0] : 0.f;
            actA[4 * FTE + e] = valid ? p[1] : 0.f;
            actA[5 * FTE + e] = valid ? p[2] : 0.f;
        }
        __syncthreads();
        if (tid < FTE) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += __popc((unsigned)s_misc[w]);
            const unsigned mine = (unsigned)s_misc[warp];
            if ((mine >> lane) & 1u) seg_first[before + __popc(mine & ((1u << lane) - 1u))] = tid;
            if (tid == 0) {
                const int nseg = __popc((unsigned)s_misc[0]) + __popc((unsigned)s_misc[1]) +
                                 __popc((unsigned)s_misc[2]) + __popc((unsigned)s_misc[3]);
                s_misc[4] = nseg;
                seg_first[nseg] = ne;
            }
        }
        if (f_in_mlp) {
            cp_async_wait_all();
            __syncthreads();
            for (int idx = tid; idx < FTE * a.c_f; idx += FTHREADS) {
                const int c = idx / FTE, e = idx - c * FTE;
                actA[(6 + c) * FTE + e] = fsm[e * a.c_f + c];
            }
        }
        __syncthreads();

        // ---- MLP ----
        const int eb = (tid & 15) * 4;
        const int j0 = (tid >> 4) * 4;
        for (int l = 0; l < a.n_layers; ++l) {
            const int K = a.dims[l], NP = L.np[l];
            const float* in = (l & 1) ? actB : actA;
            float* outb = (l & 1) ? actA : actB;
            const float* wT = smem + L.wT_off[l];
            const bool last = (l == a.n_layers - 1);
            if (last && a.f_y) cp_async_wait_all();
            float acc[8][4];
            const bool active = j0 < NP;
            if (active) {
                const float4 bv = *reinterpret_cast<const float4*>(smem + L.b_off[l] + j0);
#pragma unroll
                for (int m = 0; m < 8; ++m) { acc[m][0] = bv.x; acc[m][1] = bv.y; acc[m][2] = bv.z; acc[m][3] = bv.w; }
#pragma unroll 4
                for (int i = 0; i < K; ++i) {
                    const float4 a0 = *reinterpret_cast<const float4*>(in + i * FTE + eb);
                    const float4 a1 = *reinterpret_cast<const float4*>(in + i * FTE + 64 + eb);
                    const float4 w = *reinterpret_cast<const float4*>(wT + i * NP + j0);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int m = 0; m < 8; ++m)
#pragma unroll
                        for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(av[m], wv[n], acc[m][n]);
                }
            }
            if (!last) {
                if (active) {
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        float4 o0, o1;
                        o0.x = gelu_exact(acc[0][n]); o0.y = gelu_exact(acc[1][n]);
                        o0.z = gelu_exact(acc[2][n]); o0.w = gelu_exact(acc[3][n]);
                        o1.x = gelu_exact(acc[4][n]); o1.y = gelu_exact(acc[5][n]);
                        o1.z = gelu_exact(acc[6][n]); o1.w = gelu_exact(acc[7][n]);
                        *reinterpret_cast<float4*>(outb + (j0 + n) * FTE + eb) = o0;
                        *reinterpret_cast<float4*>(outb + (j0 + n) * FTE + 64 + eb) = o1;
                    }
                }
                __syncthreads();
            } else {
                // epilogue: (* f_y[src]) and edge-major store for the segmented reduction
                if (a.f_y) __syncthreads();            // cp.async data of all threads visible
                if (active && j0 < Cout) {
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int e = (m < 4) ? (eb + m) : (64 + eb + m - 4);
                        float4 v = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
                        if (use_f_mul) {
                            const float4 f = *reinterpret_cast<const float4*>(fsm + e * a.c_f + j0);
                            v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
                        }
                        if (a.edge_w && e < ne) {          // attention weight of the edge (integral_transform.py:161-162)
                            const float w = a.edge_w[e0 + e];
                            v.x *= w; v.y *= w; v.z *= w; v.w *= w;
                        }
                        *reinterpret_cast<float4*>(outb + e * CP + j0) = v;
                    }
                }
                __syncthreads();
                // ---- segmented reduction over the tile's query segments ----
                const float* val = outb;
                const int nseg = s_misc[4];
                for (int s = warp; s < nseg; s += FTHREADS / 32) {
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
}

// finishes segments that span several tiles: sequential (tile order) => deterministic
__global__ void __launch_bounds__(64)
gno_fwd_fixup_kernel(const GnoArgs a, float* __restrict__ out, const float* __restrict__ head_partial) {
    const int t = blockIdx.x + 1;
    if (t >= a.ntiles) return;
    const int e0 = t * FTE;
    const int q = a.csr_qry[e0];
    const int rb = a.rowptr[q], re = a.rowptr[q + 1];
    if (rb >= e0) return;                       // tile starts on a fresh segment
    if (rb < e0 - FTE) return;                  // not the first continuation tile of this segment
    const int Cout = a.dims[a.n_layers];
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
        float acc = out[(size_t)q * Cout + c];
        for (int tt = t; tt < a.ntiles; ++tt) {
            acc += head_partial[(size_t)tt * Cout + c];
            if (re <= (tt + 1) * FTE) break;
        }
        out[(size_t)q * Cout + c] = a.reduce == 0 ? acc / (float)(re - rb) : acc;
    }
}

int gno_fwd_fixup(const GnoArgs& a, float* out, const float* head_partial, cudaStream_t st) {
    if (a.ntiles > 1) {
        gno_fwd_fixup_kernel<<<a.ntiles - 1, 64, 0, st>>>(a, out, head_partial);
        GAOT_LAUNCH_CHECK();
    }
    return GAOT_OK;
}

int gno_forward_fp32(const GnoArgs& a, void* ws, size_t ws_bytes, float* out, cudaStream_t st) {
    const int Cout = a.dims[a.n_layers];
    GAOT_CUDA(cudaMemsetAsync(out, 0, (size_t)a.nq * Cout * sizeof(float), st));
    if (a.E == 0) return GAOT_OK;
    if (a.f_y && (a.c_f & 3)) { set_error("gno: feature width must be a multiple of 4"); return GAOT_ERR_UNSUPPORTED; }
    Arena ar(ws, ws_bytes);
    float* head_partial = ar.take<float>((size_t)a.ntiles * Cout);
    if (!ar.ok()) { set_error("gno_forward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const FwdSmemLayout L = fwd_layout(a);
    const size_t smem = (size_t)L.total_floats * sizeof(float);
    if (smem > 227 * 1024) { set_error("gno_forward: MLP too large for shared memory (%zu B)", smem); return GAOT_ERR_UNSUPPORTED; }
    GAOT_CUDA(cudaFuncSetAttribute(gno_fwd_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
    {
        // algorithmic bytes: E*(16 + 4D + 4C_f) + nq*(4D + 4C_out)
        GAOT_TIME_KERNEL("gno_fwd", st, (double)a.E * (16.0 + 12.0 + 4.0 * a.c_f) + (double)a.nq * (12.0 + 4.0 * Cout));
        gno_fwd_fp32_kernel<<<grid, FTHREADS, smem, st>>>(a, L, out, head_partial);
    }
    GAOT_LAUNCH_CHECK();
    if (a.ntiles > 1) {
        gno_fwd_fixup_kernel<<<a.ntiles - 1, 64, 0, st>>>(a, out, head_partial);
        GAOT_LAUNCH_CHECK();
    }
    return GAOT_OK;
}

size_t gno_forward_ws_bytes(int64_t E, int Cout) {
    return align_up((size_t)((E + FTE - 1) / FTE + 1) * Cout * sizeof(float)) + 256;
}

}  // namespace gaot
