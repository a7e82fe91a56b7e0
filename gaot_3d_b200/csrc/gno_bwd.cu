// gno_bwd.cu -- fused GNO backward (FP32 CUDA-core path).
// Recomputes the per-edge MLP forward of a 64-edge CSR tile in shared memory (no [E,*]
// activations were saved by the forward), then back-propagates:
//   g_e   = d_out[qry_e] / cnt[qry_e]                      (scatter-mean backward = gather)
//   d f_y[src_e] += g_e * k_e                              (vectorised red.global.add.v4.f32)
//   dZ_L  = g_e * f_y[src_e]  ->  dW_l += dZ_l a_l^T,  db_l += sum_e dZ_l,
//   dZ_{l-1} = (W_l^T dZ_l) * gelu'(z_{l-1})
// Weight gradients are accumulated in registers across all tiles of a persistent CTA, written
// once per CTA and summed over CTAs in a fixed order by a second kernel (deterministic).
// Replaces autograd through reference integral_transform.py:114-171 (which keeps ~2 KB/edge alive).
#include "gno_common.cuh"

namespace gaot {

constexpr int BTE = 64;            // edges per tile
constexpr int BTHREADS = 256;
constexpr int RS = BTE + 4;        // padded row stride of feature-major tiles

struct BwdSmemLayout {
    int W_off[GNO_MAX_LAYERS], wT_off[GNO_MAX_LAYERS], b_off[GNO_MAX_LAYERS];
    int np[GNO_MAX_LAYERS], kp[GNO_MAX_LAYERS];
    int h0, z[GNO_MAX_LAYERS], hs, dZa, dZb, fsm, ksm, ints;
    int total_floats;
};

static BwdSmemLayout bwd_layout(const GnoArgs& a) {
    BwdSmemLayout L;
    int off = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        L.np[l] = (a.dims[l + 1] + 3) / 4 * 4;
        L.kp[l] = (a.dims[l] + 3) / 4 * 4;
        L.W_off[l] = off; off += a.dims[l + 1] * L.kp[l];
        L.wT_off[l] = off; off += a.dims[l] * L.np[l];
    }
    for (int l = 0; l < a.n_layers; ++l) { L.b_off[l] = off; off += L.np[l]; }
    off = (off + 3) / 4 * 4;
    L.h0 = off; off += L.kp[0] * RS;
    for (int l = 1; l < a.n_layers; ++l) { L.z[l] = off; off += GNO_MAXW * RS; }
    L.z[0] = 0;
    L.hs = off; off += GNO_MAXW * RS;
    L.dZa = off; off += GNO_MAXW * RS;
    L.dZb = off; off += GNO_MAXW * RS;
    L.fsm = off; off += (a.f_y ? BTE * a.c_f : 0);
    L.ksm = off; off += 0;
    L.ints = off; off += 2 * BTE + 8;
    L.total_floats = off;
    return L;
}

template <int NL>
__global__ void __launch_bounds__(BTHREADS, 1)
gno_bwd_fp32_kernel(const GnoArgs a, const BwdSmemLayout L, const float* __restrict__ d_out,
                    float* __restrict__ d_f, float* __restrict__ partial /* [grid][n_params] */) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    float* h0 = smem + L.h0;
    float* hs = smem + L.hs;
    float* dZa = smem + L.dZa;
    float* dZb = smem + L.dZb;
    float* fsm = smem + L.fsm;
    int* s_src = reinterpret_cast<int*>(smem + L.ints);
    int* s_qry = s_src + BTE;

    // ---- stage W (row-major, padded rows), W^T and biases ----
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1], NP = L.np[l], KP = L.kp[l];
        const float* W = a.params + a.w_off[l];
        float* Ws = smem + L.W_off[l];
        float* wT = smem + L.wT_off[l];
        for (int idx = tid; idx < N * KP; idx += BTHREADS) {
            const int j = idx / KP, i = idx - j * KP;
            Ws[idx] = i < K ? W[j * K + i] : 0.f;
        }
        for (int idx = tid; idx < K * NP; idx += BTHREADS) {
            const int i = idx / NP, j = idx - i * NP;
            wT[idx] = j < N ? W[j * K + i] : 0.f;
        }
        float* bs = smem + L.b_off[l];
        for (int j = tid; j < NP; j += BTHREADS) bs[j] = j < N ? a.params[a.b_off[l] + j] : 0.f;
    }
    __syncthreads();

    const int Cout = a.dims[NL];
    const bool use_f_mul = (a.transform == 0 || a.transform == 1);
    const bool f_in_mlp = (a.transform == 1 || a.transform == 2);
    const int cf4 = a.c_f >> 2;

    // persistent gradient accumulators: dW[l][n][m] <-> W_l[j = tj + 16 n][i = ti + 16 m]
    float dW[NL][4][4];
    float db[NL][4];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            db[l][n] = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) dW[l][n][m] = 0.f;
        }

    const int t16 = tid & 15, u16 = tid >> 4;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int e0 = tile * BTE;
        const int ne = min(BTE, a.E - e0);
        if (tid < BTE) {
            const bool valid = tid < ne;
            s_src[tid] = valid ? a.csr_src[e0 + tid] : 0;
            s_qry[tid] = valid ? a.csr_qry[e0 + tid] : -1;
        }
        __syncthreads();
        if (a.f_y) {
            for (int idx = tid; idx < BTE * cf4; idx += BTHREADS) {
                const int e = idx / cf4, ch = idx - e * cf4;
                if (e < ne) cp_async16(fsm + e * a.c_f + ch * 4, a.f_y + (size_t)s_src[e] * a.c_f + ch * 4);
                else *reinterpret_cast<float4*>(fsm + e * a.c_f + ch * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            cp_async_commit();
        }
        if (tid < BTE) {
            const bool valid = tid < ne;
            const float* p = a.y_pos + (size_t)s_src[tid] * 3;
            h0[0 * RS + tid] = valid ? p[0] : 0.f;
            h0[1 * RS + tid] = valid ? p[1] : 0.f;
            h0[2 * RS + tid] = valid ? p[2] : 0.f;
        } else if (tid < 2 * BTE) {
            const int e = tid - BTE;
            const bool valid = e < ne;
            const float* p = a.x_pos + (size_t)(valid ? s_qry[e] : 0) * 3;
            h0[3 * RS + e] = valid ? p[0] : 0.f;
            h0[4 * RS + e] = valid ? p[1] : 0.f;
            h0[5 * RS + e] = valid ? p[2] : 0.f;
        } else {
            // zero the padding rows of h0 (rows K0..KP0-1) once per tile (cheap)
            const int K0 = a.dims[0], KP0 = L.kp[0];
            for (int idx = tid - 2 * BTE; idx < (KP0 - K0) * BTE; idx += BTHREADS - 2 * BTE) {
                const int r = K0 + idx / BTE, e = idx % BTE;
                h0[r * RS + e] = 0.f;
            }
        }
        if (a.f_y) { cp_async_wait_all(); }
        __syncthreads();
        if (f_in_mlp) {
            for (int idx = tid; idx < BTE * a.c_f; idx += BTHREADS) {
                const int c = idx / BTE, e = idx - c * BTE;
                h0[(6 + c) * RS + e] = fsm[e * a.c_f + c];
            }
            __syncthreads();
        }

        // =============== forward recompute ===============
        // thread tile: edges eF..eF+3 (t16), outputs jF..jF+3 (u16)
        const int eF = t16 * 4, jF = u16 * 4;
        float kacc[4][4];      // last layer output k[e][c] for this thread's tile
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const int K = a.dims[l], NP = L.np[l];
            const float* in = (l == 0) ? h0 : hs;
            const float* wT = smem + L.wT_off[l];
            float acc[4][4];
            const bool active = jF < NP;
            if (active) {
                const float4 bv = *reinterpret_cast<const float4*>(smem + L.b_off[l] + jF);
#pragma unroll
                for (int m = 0; m < 4; ++m) { acc[m][0] = bv.x; acc[m][1] = bv.y; acc[m][2] = bv.z; acc[m][3] = bv.w; }
#pragma unroll 4
                for (int i = 0; i < K; ++i) {
                    const float4 av = *reinterpret_cast<const float4*>(in + i * RS + eF);
                    const float4 w = *reinterpret_cast<const float4*>(wT + i * NP + jF);
                    const float ae[4] = {av.x, av.y, av.z, av.w};
                    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(ae[m], wv[n], acc[m][n]);
                }
            }
            if (l < NL - 1) {
                __syncthreads();                       // everyone finished reading hs
                if (active) {
                    float* zl = smem + L.z[l + 1];
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const float4 zv = make_float4(acc[0][n], acc[1][n], acc[2][n], acc[3][n]);
                        *reinterpret_cast<float4*>(zl + (jF + n) * RS + eF) = zv;
                        const float4 hv = make_float4(gelu_exact(zv.x), gelu_exact(zv.y), gelu_exact(zv.z), gelu_exact(zv.w));
                        *reinterpret_cast<float4*>(hs + (jF + n) * RS + eF) = hv;
                    }
                }
                __syncthreads();
            } else {
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 4; ++n) kacc[m][n] = active ? acc[m][n] : 0.f;
            }
        }

        // =============== output-side gradients ===============
        // g = d_out[q] / cnt;  dZ_L = g (* f);  d_f += g * k
        {
            const bool active = jF < Cout;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int e = eF + m;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active && e < ne) {
                    const int q = s_qry[e];
                    g = *reinterpret_cast<const float4*>(d_out + (size_t)q * Cout + jF);
                    if (a.edge_w) {
                        if (a.d_edge_w) {                  // d loss / d w_e = <d_out[q], k_e (* f_e)>: 4 of the Cout terms here
                            float4 kf = make_float4(kacc[m][0], kacc[m][1], kacc[m][2], kacc[m][3]);
                            if (use_f_mul) {
                                const float4 f = *reinterpret_cast<const float4*>(fsm + e * a.c_f + jF);
                                kf.x *= f.x; kf.y *= f.y; kf.z *= f.z; kf.w *= f.w;
                            }
                            atomicAdd(a.d_edge_w + e0 + e, g.x * kf.x + g.y * kf.y + g.z * kf.z + g.w * kf.w);
                        }
                        const float w = a.edge_w[e0 + e];
                        g.x *= w; g.y *= w; g.z *= w; g.w *= w;
                    } else if (a.reduce == 0) {
                        const float inv = 1.0f / (float)(a.rowptr[q + 1] - a.rowptr[q]);
                        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
                    }
                    if (use_f_mul) {
                        if (d_f) {
                            float4 v = make_float4(g.x * kacc[m][0], g.y * kacc[m][1], g.z * kacc[m][2], g.w * kacc[m][3]);
                            atomicAdd(reinterpret_cast<float4*>(d_f + (size_t)s_src[e] * a.c_f + jF), v);
                        }
                        const float4 f = *reinterpret_cast<const float4*>(fsm + e * a.c_f + jF);
                        g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
                    }
                }
                kacc[m][0] = g.x; kacc[m][1] = g.y; kacc[m][2] = g.z; kacc[m][3] = g.w;   // now dZ_L
            }
            if (jF < GNO_MAXW) {
#pragma unroll
                for (int n = 0; n < 4; ++n)
                    *reinterpret_cast<float4*>(dZa + (jF + n) * RS + eF) =
                        make_float4(kacc[0][n], kacc[1][n], kacc[2][n], kacc[3][n]);
            }
        }
        __syncthreads();

        // =============== backward through the layers ===============
#pragma unroll
        for (int l = NL - 1; l >= 0; --l) {
            const int K = a.dims[l], N = a.dims[l + 1], KP = L.kp[l];
            const float* dZ = ((NL - 1 - l) & 1) ? dZb : dZa;
            float* dZn = ((NL - 1 - l) & 1) ? dZa : dZb;
            // a_l: h0 for l == 0, gelu(z_l) otherwise (hs still holds a_{NL-1} after the forward)
            if (l >= 1 && l < NL - 1) {
                const float* zl = smem + L.z[l];
                for (int idx = tid; idx < K * (BTE / 4); idx += BTHREADS) {
                    const int r = idx / (BTE / 4), c4 = (idx - r * (BTE / 4)) * 4;
                    const float4 zv = *reinterpret_cast<const float4*>(zl + r * RS + c4);
                    *reinterpret_cast<float4*>(hs + r * RS + c4) =
                        make_float4(gelu_exact(zv.x), gelu_exact(zv.y), gelu_exact(zv.z), gelu_exact(zv.w));
                }
                __syncthreads();
            }
            const float* al = (l == 0) ? h0 : hs;
            // ---- dW_l[j][i] += sum_e dZ[j][e] a_l[i][e];  i = t16 + 16 m, j = u16 + 16 n ----
            {
                float4 dzv[4], av[4];
                for (int e = 0; e < BTE; e += 4) {
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const int j = u16 + 16 * n;
                        dzv[n] = j < N ? *reinterpret_cast<const float4*>(dZ + j * RS + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int i = t16 + 16 * m;
                        av[m] = i < K ? *reinterpret_cast<const float4*>(al + i * RS + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        db[l][n] += (dzv[n].x + dzv[n].y) + (dzv[n].z + dzv[n].w);
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            float s = dW[l][n][m];
                            s = fmaf(dzv[n].x, av[m].x, s); s = fmaf(dzv[n].y, av[m].y, s);
                            s = fmaf(dzv[n].z, av[m].z, s); s = fmaf(dzv[n].w, av[m].w, s);
                            dW[l][n][m] = s;
                        }
                    }
                }
            }
            // ---- dA_l[i][e] = sum_j W_l[j][i] dZ[j][e]  (needed for l >= 1, or l == 0 with f in the MLP) ----
            if (l >= 1 || f_in_mlp) {
                const int iB = t16 * 4, eB = u16 * 4;
                const float* Ws = smem + L.W_off[l];
                float acc[4][4];
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < 4; ++n) acc[m][n] = 0.f;
                const bool active = iB < KP;
                if (active) {
#pragma unroll 4
                    for (int j = 0; j < N; ++j) {
                        const float4 w = *reinterpret_cast<const float4*>(Ws + j * KP + iB);
                        const float4 d = *reinterpret_cast<const float4*>(dZ + j * RS + eB);
                        const float wv[4] = {w.x, w.y, w.z, w.w};
                        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(wv[m], dv[n], acc[m][n]);
                    }
                }
                if (l >= 1) {
                    if (active) {
                        const float* zl = smem + L.z[l];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const float4 zv = *reinterpret_cast<const float4*>(zl + (iB + m) * RS + eB);
                            *reinterpret_cast<float4*>(dZn + (iB + m) * RS + eB) =
                                make_float4(acc[m][0] * gelu_grad(zv.x), acc[m][1] * gelu_grad(zv.y),
                                            acc[m][2] * gelu_grad(zv.z), acc[m][3] * gelu_grad(zv.w));
                        }
                    }
                } else if (d_f) {
                    // l == 0, f in the MLP input: rows 6.. of dA_0 are d f_y[src]
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int i = iB + m;
                        if (i >= 6 && i < K) {
#pragma unroll
                            for (int n = 0; n < 4; ++n) {
                                const int e = eB + n;
                                if (e < ne) atomicAdd(d_f + (size_t)s_src[e] * a.c_f + (i - 6), acc[m][n]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- flush per-CTA weight-gradient partials ----
    float* mine = partial + (size_t)blockIdx.x * a.n_params;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = a.dims[l], N = a.dims[l + 1];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int j = u16 + 16 * n;
            if (j < N) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int i = t16 + 16 * m;
                    if (i < K) mine[a.w_off[l] + j * K + i] = dW[l][n][m];
                }
                if (t16 == 0) mine[a.b_off[l] + j] = db[l][n];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
gno_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int n_params, float* __restrict__ d_params) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_params) return;
    float s = 0.f;
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * n_params + p];
    d_params[p] = s;
}

int gno_bwd_reduce(const float* partial, int nparts, int n_params, float* d_params, cudaStream_t st) {
    gno_bwd_reduce_kernel<<<(n_params + 255) / 256, 256, 0, st>>>(partial, nparts, n_params, d_params);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

size_t gno_backward_ws_bytes(int n_params) { return align_up((size_t)2 * kNumSMs * n_params * sizeof(float)) + 256; }   // <= 2 CTAs per SM

int gno_backward_fp32(const GnoArgs& a_in, const float* d_out, void* ws, size_t ws_bytes,
                      float* d_params, float* d_f, cudaStream_t st) {
    GnoArgs a = a_in;
    a.ntiles = (a.E + BTE - 1) / BTE;
    if (d_f) GAOT_CUDA(cudaMemsetAsync(d_f, 0, (size_t)a.n_src * a.c_f * sizeof(float), st));
    if (a.d_edge_w && a.E > 0) GAOT_CUDA(cudaMemsetAsync(a.d_edge_w, 0, (size_t)a.E * sizeof(float), st));
    if (a.E == 0) {
        GAOT_CUDA(cudaMemsetAsync(d_params, 0, (size_t)a.n_params * sizeof(float), st));
        return GAOT_OK;
    }
    if (a.f_y && (a.c_f & 3)) { set_error("gno: feature width must be a multiple of 4"); return GAOT_ERR_UNSUPPORTED; }
    for (int l = 1; l <= a.n_layers; ++l)
        if (a.dims[l] & 3) { set_error("gno_backward: layer widths must be multiples of 4"); return GAOT_ERR_UNSUPPORTED; }
    if (a.dims[0] > GNO_MAXW) { set_error("gno_backward: MLP input width %d > %d", a.dims[0], GNO_MAXW); return GAOT_ERR_UNSUPPORTED; }
    Arena ar(ws, ws_bytes);
    const int grid = a.ntiles < kNumSMs ? a.ntiles : kNumSMs;
    float* partial = ar.take<float>((size_t)grid * a.n_params);
    if (!ar.ok()) { set_error("gno_backward: workspace too small"); return GAOT_ERR_WORKSPACE; }
    const BwdSmemLayout L = bwd_layout(a);
    const size_t smem = (size_t)L.total_floats * sizeof(float);
    if (smem > 227 * 1024) { set_error("gno_backward: MLP too large for shared memory (%zu B)", smem); return GAOT_ERR_UNSUPPORTED; }
    const int CoutB = a.dims[a.n_layers];
    GAOT_TIME_KERNEL("gno_bwd", st, (double)a.E * (16.0 + 12.0 + 4.0 * a.c_f) + (double)a.nq * (12.0 + 8.0 * CoutB) + (double)a.n_src * 4.0 * a.c_f);
#define GAOT_BWD_CASE(NL)                                                                                   \
    case NL:                                                                                                \
        GAOT_CUDA(cudaFuncSetAttribute(gno_bwd_fp32_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gno_bwd_fp32_kernel<NL><<<grid, BTHREADS, smem, st>>>(a, L, d_out, d_f, partial);                   \
        break;
    switch (a.n_layers) {
        GAOT_BWD_CASE(1) GAOT_BWD_CASE(2) GAOT_BWD_CASE(3) GAOT_BWD_CASE(4) GAOT_BWD_CASE(5)
        default: set_error("gno_backward: n_layers %d unsupported (1..5)", a.n_layers); return GAOT_ERR_UNSUPPORTED;
    }
#undef GAOT_BWD_CASE
    GAOT_LAUNCH_CHECK();
    gno_bwd_reduce_kernel<<<(a.n_params + 255) / 256, 256, 0, st>>>(partial, grid, a.n_params, d_params);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // namespace gaot
