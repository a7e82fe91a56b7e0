// geo.cu -- statistical geometric embedding features in one pass over the query-major CSR.
// Replaces 5 scatters with atomics, an [E,3,3] outer-product tensor and a batched cuSOLVER
// eigvalsh (reference src/model/layers/geoembed.py:99-175) by: one thread per query accumulating
// the moments of (y - x) centred on the query point (magnitudes <= r, so the one-pass
// covariance keeps fp32 accuracy), then cov = E[dd^T] - delta delta^T and the eigenvalues of
// the 3x3 symmetric cov + 1e-6 I by cyclic Jacobi in fp64 registers.  The z-score over all
// queries (geoembed.py:177-180) is a separate two-stage reduction (global dependency).
#include "common.cuh"

namespace gaot {

__device__ __forceinline__ void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq) {
    // annihilate a_pq; r is the third index
    if (fabs(apq) < 1e-300) return;
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    app -= t * apq; aqq += t * apq; apq = 0.0;
    const double nrp = c * arp - s * arq, nrq = s * arp + c * arq;
    arp = nrp; arq = nrq;
}

// moments of (y - x) over one CSR row, centred on the query: m = {n, sum d, sum d^2, sum dx,dy,dz, sum xx,xy,xz,yy,yz,zz}.
// Sums (not means), so the partials of several physical-point shards add up (sharded encoder).
__device__ __forceinline__ void geo_row_moments(const float* __restrict__ src_pos, const int32_t* __restrict__ csr_src,
                                                int b, int e, float qx, float qy, float qz, float (&m)[12]) {
    float sd = 0.f, sd2 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    float cxx = 0.f, cxy = 0.f, cxz = 0.f, cyy = 0.f, cyz = 0.f, czz = 0.f;
    for (int p = b; p < e; ++p) {
        const int s = csr_src[p];
        const float dx = src_pos[(size_t)s * 3] - qx, dy = src_pos[(size_t)s * 3 + 1] - qy,
                    dz = src_pos[(size_t)s * 3 + 2] - qz;
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float d = sqrtf(d2);
        sd += d; sd2 += d * d;
        sx += dx; sy += dy; sz += dz;
        cxx += dx * dx; cxy += dx * dy; cxz += dx * dz; cyy += dy * dy; cyz += dy * dz; czz += dz * dz;
    }
    m[0] = (float)(e - b); m[1] = sd; m[2] = sd2; m[3] = sx; m[4] = sy; m[5] = sz;
    m[6] = cxx; m[7] = cxy; m[8] = cxz; m[9] = cyy; m[10] = cyz; m[11] = czz;
}

// moments -> [N_i, D_avg, D_var, delta(3), eigenvalues(3) descending]; zero row for an empty query
__device__ __forceinline__ void geo_finalize(const float (&m)[12], float* __restrict__ o) {
    const float n = m[0];
    if (n <= 0.f) {
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = 0.f;
        return;
    }
    const float inv = 1.0f / n;
    const float davg = m[1] * inv;
    float dvar = m[2] * inv - davg * davg;
    dvar = dvar > 0.f ? dvar : 0.f;
    const double mx = (double)m[3] / n, my = (double)m[4] / n, mz = (double)m[5] / n;
    double axx = (double)m[6] / n - mx * mx + 1e-6, ayy = (double)m[9] / n - my * my + 1e-6,
           azz = (double)m[11] / n - mz * mz + 1e-6;
    double axy = (double)m[7] / n - mx * my, axz = (double)m[8] / n - mx * mz, ayz = (double)m[10] / n - my * mz;
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
        jacobi_rot(axx, ayy, axy, axz, ayz);   // (p,q) = (x,y), r = z
        jacobi_rot(axx, azz, axz, axy, ayz);   // (x,z), r = y
        jacobi_rot(ayy, azz, ayz, axy, axz);   // (y,z), r = x
        if (fabs(axy) + fabs(axz) + fabs(ayz) < 1e-22) break;
    }
    double l0 = axx, l1 = ayy, l2 = azz, t;
    if (l0 < l1) { t = l0; l0 = l1; l1 = t; }
    if (l1 < l2) { t = l1; l1 = l2; l2 = t; }
    if (l0 < l1) { t = l0; l0 = l1; l1 = t; }
    o[0] = n; o[1] = davg; o[2] = dvar;
    o[3] = (float)mx; o[4] = (float)my; o[5] = (float)mz;
    o[6] = (float)l0; o[7] = (float)l1; o[8] = (float)l2;
}

__global__ void __launch_bounds__(128)
geo_stats_kernel(const float* __restrict__ src_pos, const float* __restrict__ qry_pos, int64_t nq,
                 const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_src,
                 float* __restrict__ feat) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    float m[12];
    geo_row_moments(src_pos, csr_src, rowptr[q], rowptr[q + 1], qry_pos[q * 3], qry_pos[q * 3 + 1], qry_pos[q * 3 + 2], m);
    geo_finalize(m, feat + q * 9);
}

// the two halves as separate kernels: per-shard moment partials [nq,12] -> (all-reduce) -> features
__global__ void __launch_bounds__(128)
geo_moments_kernel(const float* __restrict__ src_pos, const float* __restrict__ qry_pos, int64_t nq,
                   const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_src, float* __restrict__ mom) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    float m[12];
    geo_row_moments(src_pos, csr_src, rowptr[q], rowptr[q + 1], qry_pos[q * 3], qry_pos[q * 3 + 1], qry_pos[q * 3 + 2], m);
    float4* o = reinterpret_cast<float4*>(mom + q * 12);
    o[0] = make_float4(m[0], m[1], m[2], m[3]); o[1] = make_float4(m[4], m[5], m[6], m[7]); o[2] = make_float4(m[8], m[9], m[10], m[11]);
}
__global__ void __launch_bounds__(128)
geo_from_moments_kernel(const float* __restrict__ mom, int64_t nq, float* __restrict__ feat) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float4* i = reinterpret_cast<const float4*>(mom + q * 12);
    const float4 a = i[0], b = i[1], c = i[2];
    const float m[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    geo_finalize(m, feat + q * 9);
}

// ---- global z-score: per-feature mean and unbiased std over all queries (fp64 accumulation) ----
constexpr int ZS_BLOCKS = 296;
__global__ void __launch_bounds__(256)
zscore_partial_kernel(const float* __restrict__ feat, int64_t nq, int nfeat, double* __restrict__ part) {
    // part[block][2*nfeat]: sum, sum of squares (about 0; fp64 keeps 1e-5 relative accuracy)
    __shared__ double sh[256];
    for (int f = 0; f < nfeat; ++f) {
        double s = 0.0, s2 = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
            const double v = feat[i * nfeat + f];
            s += v; s2 += v * v;
        }
        for (int pass = 0; pass < 2; ++pass) {
            sh[threadIdx.x] = pass == 0 ? s : s2;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) part[(size_t)blockIdx.x * 2 * nfeat + 2 * f + pass] = sh[0];
            __syncthreads();
        }
    }
}
__global__ void zscore_final_kernel(const double* __restrict__ part, int nblocks, int64_t nq, int nfeat,
                                    float* __restrict__ mean_std) {
    const int f = threadIdx.x;
    if (f >= nfeat) return;
    double s = 0.0, s2 = 0.0;
    for (int b = 0; b < nblocks; ++b) { s += part[(size_t)b * 2 * nfeat + 2 * f]; s2 += part[(size_t)b * 2 * nfeat + 2 * f + 1]; }
    const double mean = s / (double)nq;
    double var = nq > 1 ? (s2 - (double)nq * mean * mean) / (double)(nq - 1) : 0.0;
    if (var < 0.0) var = 0.0;
    float sd = (float)sqrt(var);
    if (nq <= 1) sd = nanf("");                   // torch.std of one sample is NaN (NaN < 1e-6 is false)
    if (sd < 1e-6f) sd = 1.0f;
    mean_std[f] = (float)mean; mean_std[nfeat + f] = sd;
}
__global__ void __launch_bounds__(256)
zscore_apply_kernel(float* __restrict__ feat, int64_t total, int nfeat, const float* __restrict__ mean_std) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int f = (int)(i % nfeat);
    feat[i] = (feat[i] - mean_std[f]) / mean_std[nfeat + f];
}

}  // namespace gaot

using namespace gaot;

extern "C" {

int gaot_geo_stats(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq,
                   const int32_t* rowptr, const int32_t* csr_src, float* feat, void* stream) {
    (void)n_src;
    GAOT_CHECK_ARG(nq >= 0 && feat != nullptr, "geo_stats: bad arguments");
    if (nq == 0) return GAOT_OK;
    geo_stats_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, (cudaStream_t)stream>>>(src_pos, qry_pos, nq, rowptr, csr_src, feat);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_geo_moments(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq,
                     const int32_t* rowptr, const int32_t* csr_src, float* moments, void* stream) {
    (void)n_src;
    GAOT_CHECK_ARG(nq >= 0 && moments != nullptr, "geo_moments: bad arguments");
    if (nq == 0) return GAOT_OK;
    geo_moments_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, (cudaStream_t)stream>>>(src_pos, qry_pos, nq, rowptr, csr_src, moments);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_geo_from_moments(const float* moments, int64_t nq, float* feat, void* stream) {
    GAOT_CHECK_ARG(nq >= 0 && moments != nullptr && feat != nullptr, "geo_from_moments: bad arguments");
    if (nq == 0) return GAOT_OK;
    geo_from_moments_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, (cudaStream_t)stream>>>(moments, nq, feat);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

size_t gaot_geo_zscore_workspace_bytes(int64_t nq) {
    (void)nq;
    return align_up((size_t)ZS_BLOCKS * 2 * 32 * sizeof(double)) + align_up(64 * sizeof(float)) + 256;
}

int gaot_geo_zscore(float* feat, int64_t nq, int32_t nfeat, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(nfeat >= 1 && nfeat <= 32, "geo_zscore: nfeat must be in [1,32]");
    if (nq == 0) return GAOT_OK;
    Arena ar(ws, ws_bytes);
    double* part = ar.take<double>((size_t)ZS_BLOCKS * 2 * 32);
    float* ms = ar.take<float>(64);
    if (!ar.ok()) { set_error("geo_zscore: workspace too small"); return GAOT_ERR_WORKSPACE; }
    zscore_partial_kernel<<<ZS_BLOCKS, 256, 0, st>>>(feat, nq, nfeat, part);
    GAOT_LAUNCH_CHECK();
    zscore_final_kernel<<<1, 32, 0, st>>>(part, ZS_BLOCKS, nq, nfeat, ms);
    GAOT_LAUNCH_CHECK();
    const int64_t total = nq * nfeat;
    zscore_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, total, nfeat, ms);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // extern "C"
