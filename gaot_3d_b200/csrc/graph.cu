// graph.cu -- uniform-grid / cell-list radius and kNN search with a prefix-sum CSR edge
// emitter, coalesce, edge mask and CSR side-band.  Replaces torch_cluster radius/knn
// (reference src/model/layers/magno.py:183-200,:242-260), torch_geometric coalesce
// (:220,:293) and dropout_edge (:367).
//
// Pipeline (all stream-ordered, no host sync except the optional E read-back):
//   bbox(sources) -> grid params (device-resident) -> bin sources by cell (histogram,
//   scan, scatter into a packed {x,y,z,idx} array) -> bin queries by the same grid (so a
//   warp's 32 queries share cells: coherent loops, broadcast loads) -> one thread per
//   query walks its (2*reach+1)^3 block: count pass -> scan -> emit pass.
// HBM-bound integer/byte work: coalesced 16-byte source records, int32 internals, int64
// only at the edge_index boundary.
#include "common.cuh"
#include <cstdlib>
#include "graph_core.cuh"

namespace gaot {

constexpr int MAX_CELLS = 1 << 21;
constexpr int MAX_DIM = 1024;

// ---- float <-> order-preserving uint (for atomic min/max of a bounding box) ----
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void bbox_init_kernel(unsigned* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;       // mins
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;           // maxs
}

__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ pos, int64_t n, unsigned* __restrict__ bb) {
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = pos[i * 3 + a];
            mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bb[a], f2ord(mn[a]));
            atomicMax(&bb[3 + a], f2ord(mx[a]));
        }
    }
}

__global__ void grid_params_kernel(const unsigned* __restrict__ bb, int64_t n_src, float r, int mode,
                                   GridParams* __restrict__ gp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = ord2f(bb[a]); hi[a] = ord2f(bb[3 + a]); }
    *gp = compute_grid_params(lo, hi, n_src, r, mode, MAX_CELLS, MAX_DIM);
}

__global__ void __launch_bounds__(256)
cell_hist_kernel(const float* __restrict__ pos, int64_t n, const GridParams* __restrict__ gp,
                 int32_t* __restrict__ cell_of_pt, int32_t* __restrict__ cell_cnt) {
    const GridParams g = *gp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of(g, pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]);
    cell_of_pt[i] = c;
    atomicAdd(&cell_cnt[c], 1);
}

__global__ void __launch_bounds__(256)
cell_scatter_src_kernel(const float* __restrict__ pos, int64_t n, const int32_t* __restrict__ cell_of_pt,
                        const int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_fill,
                        SrcPoint* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of_pt[i];
    const int slot = cell_start[c] + atomicAdd(&cell_fill[c], 1);
    SrcPoint s; s.x = pos[i * 3]; s.y = pos[i * 3 + 1]; s.z = pos[i * 3 + 2]; s.idx = (int)i;
    out[slot] = s;
}

__global__ void __launch_bounds__(256)
cell_scatter_perm_kernel(int64_t n, const int32_t* __restrict__ cell_of_pt,
                         const int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_fill,
                         int32_t* __restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of_pt[i];
    perm[cell_start[c] + atomicAdd(&cell_fill[c], 1)] = (int)i;
}

template <int MAXCAP>
__global__ void __launch_bounds__(128)
radius_count_kernel(const float* __restrict__ y, int64_t ny, const int32_t* __restrict__ qperm,
                    const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                    const SrcPoint* __restrict__ pts, float r2, int cap, int32_t* __restrict__ counts) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ny) return;
    const GridParams g = *gp;
    const int q = qperm[t];
    counts[q] = radius_query<false>(g, cell_start, pts, y[(int64_t)q * 3], y[(int64_t)q * 3 + 1],
                                    y[(int64_t)q * 3 + 2], r2, cap, nullptr);
}

template <int MAXCAP>
__global__ void __launch_bounds__(128)
radius_emit_kernel(const float* __restrict__ y, int64_t ny, const int32_t* __restrict__ qperm,
                   const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                   const SrcPoint* __restrict__ pts, float r2, int cap, const int32_t* __restrict__ rowptr,
                   int64_t* __restrict__ out_y, int64_t* __restrict__ out_x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ny) return;
    const GridParams g = *gp;
    const int q = qperm[t];
    const int beg = rowptr[q];
    if (rowptr[q + 1] == beg) return;
    int list[MAXCAP];
    const int n = radius_query<true>(g, cell_start, pts, y[(int64_t)q * 3], y[(int64_t)q * 3 + 1],
                                     y[(int64_t)q * 3 + 2], r2, cap, list);
    for (int j = 0; j < n; ++j) {
        out_y[beg + j] = q;
        out_x[beg + j] = list[j];
    }
}

// ---- warp-per-query radius search: the encoder side (queries = latent tokens, sources = the physical cloud) has
// hundreds to thousands of candidates per query (582 at 500 K points, ~9 000 at 8 M) and few queries: one thread per query
// is a long serial scan on a fraction of the SMs.  Here the 32 lanes stride over the candidates of every cell column
// (coalesced 16-byte SrcPoint loads); COUNT is a warp sum; EMIT keeps the `cap` smallest matching source indices
// as a sorted list distributed over the lanes (lane i = i-th smallest) and inserts the rare improving candidates with one
// shuffle each (expected ~ cap (1 + ln(m / cap)) insertions for m matches).  Same result as radius_query<true>: the
// first `cap` matches by ascending source index, in ascending order (torch_cluster's CUDA semantics).  cap <= 32.
template <bool EMIT>
__global__ void __launch_bounds__(128)
radius_warp_kernel(const float* __restrict__ y, int64_t ny, const int32_t* __restrict__ qperm,
                   const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                   const SrcPoint* __restrict__ pts, float r2, int cap, int32_t* __restrict__ counts,
                   const int32_t* __restrict__ rowptr, int64_t* __restrict__ out_y, int64_t* __restrict__ out_x) {
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= ny) return;
    const GridParams g = *gp;
    const int q = qperm[t];
    int obeg = 0, nout = 0;
    if (EMIT) {
        obeg = rowptr[q];
        nout = rowptr[q + 1] - obeg;
        if (nout == 0) return;
    }
    const float qx = y[(int64_t)q * 3], qy = y[(int64_t)q * 3 + 1], qz = y[(int64_t)q * 3 + 2];
    const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx), cy = cell_coord(qy, g.oy, g.inv_h, g.ny), cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
    const int R = g.reach;
    const int x0 = clampi(cx - R, 0, g.nx - 1), x1 = clampi(cx + R, 0, g.nx - 1);
    const int y0 = clampi(cy - R, 0, g.ny - 1), y1 = clampi(cy + R, 0, g.ny - 1);
    const int z0 = clampi(cz - R, 0, g.nz - 1), z1 = clampi(cz + R, 0, g.nz - 1);
    int n = 0;
    int best = 0x7fffffff;                       // EMIT: lane i holds the i-th smallest matching index so far
    for (int ix = x0; ix <= x1; ++ix) {
        for (int iy = y0; iy <= y1; ++iy) {
            const int beg = cell_start[cell_id(g, ix, iy, z0)];
            const int end = cell_start[cell_id(g, ix, iy, z1) + 1];
            for (int base = beg; base < end; base += 32) {
                const int p = base + lane;
                bool match = false;
                int idx = 0;
                if (p < end) {
                    const SrcPoint s = pts[p];
                    match = dist2(s.x, s.y, s.z, qx, qy, qz) < r2;
                    idx = s.idx;
                }
                if (!EMIT) {
                    n += match ? 1 : 0;
                } else {
                    int thr = __shfl_sync(0xffffffffu, best, cap - 1);
                    unsigned mask = __ballot_sync(0xffffffffu, match && idx < thr);
                    while (mask) {
                        const int sl = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int v = __shfl_sync(0xffffffffu, idx, sl);
                        if (v >= thr) continue;                          // the threshold tightened since the ballot
                        int up = __shfl_up_sync(0xffffffffu, best, 1);
                        if (lane == 0) up = (int)0x80000000;
                        if (best > v) best = max(v, up);
                        thr = __shfl_sync(0xffffffffu, best, cap - 1);
                    }
                }
            }
        }
    }
    if (!EMIT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) counts[q] = n > cap ? cap : n;
    } else if (lane < nout) {
        out_y[obeg + lane] = q;
        out_x[obeg + lane] = best;
    }
}

template <int KMAX>
__global__ void __launch_bounds__(128)
knn_kernel(const float* __restrict__ y, int64_t ny, const int32_t* __restrict__ qperm,
           const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
           const SrcPoint* __restrict__ pts, int k, int64_t* __restrict__ out_y, int64_t* __restrict__ out_x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ny) return;
    const GridParams g = *gp;
    const int q = qperm[t];
    float bd[KMAX]; int bi[KMAX];
    const int n = knn_query<KMAX>(g, cell_start, pts, y[(int64_t)q * 3], y[(int64_t)q * 3 + 1],
                                  y[(int64_t)q * 3 + 2], k, bd, bi);
    for (int j = 0; j < k; ++j) {
        out_y[(int64_t)q * k + j] = q;
        out_x[(int64_t)q * k + j] = j < n ? bi[j] : -1;
    }
}

// ---------------------------------------------------------------- shared host-side plumbing
struct CellWs {
    GridParams* gp; unsigned* bb;
    int32_t* cell_start;      // [MAX_CELLS + 1]
    int32_t* cell_tmp;        // [MAX_CELLS + 1] histogram / fill counters
    int32_t* qcell_start;     // [MAX_CELLS + 1]
    int32_t* cell_of_pt;      // [max(nx, ny)]
    SrcPoint* pts;            // [nx]
    int32_t* qperm;           // [ny]
    int32_t* counts;          // [ny + 1]
    char* scan_ws; size_t scan_bytes;
};

static size_t cell_ws_bytes(int64_t nx, int64_t ny) {
    const int64_t nmax = nx > ny ? nx : ny;
    size_t b = 0;
    b += align_up(sizeof(GridParams)) + align_up(8 * sizeof(unsigned));
    b += 3 * align_up((size_t)(MAX_CELLS + 1) * sizeof(int32_t));
    b += align_up((size_t)nmax * sizeof(int32_t));
    b += align_up((size_t)nx * sizeof(SrcPoint));
    b += align_up((size_t)ny * sizeof(int32_t));
    b += align_up((size_t)(ny + 1) * sizeof(int32_t));
    const int64_t smax = (nmax + 1) > (MAX_CELLS + 1) ? (nmax + 1) : (MAX_CELLS + 1);
    b += align_up(scan_workspace_bytes(smax));
    return b + 1024;
}

static bool carve(CellWs& w, void* ws, size_t ws_bytes, int64_t nx, int64_t ny) {
    Arena ar(ws, ws_bytes);
    const int64_t nmax = nx > ny ? nx : ny;
    w.gp = ar.take<GridParams>(1);
    w.bb = ar.take<unsigned>(8);
    w.cell_start = ar.take<int32_t>(MAX_CELLS + 1);
    w.cell_tmp = ar.take<int32_t>(MAX_CELLS + 1);
    w.qcell_start = ar.take<int32_t>(MAX_CELLS + 1);
    w.cell_of_pt = ar.take<int32_t>((size_t)nmax);
    w.pts = ar.take<SrcPoint>((size_t)nx);
    w.qperm = ar.take<int32_t>((size_t)ny);
    w.counts = ar.take<int32_t>((size_t)ny + 1);
    const int64_t smax = (nmax + 1) > (MAX_CELLS + 1) ? (nmax + 1) : (MAX_CELLS + 1);
    w.scan_bytes = scan_workspace_bytes(smax);
    w.scan_ws = ar.take<char>(w.scan_bytes);
    return ar.ok();
}

static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// bbox + grid + source binning + query binning
static int build_cells(const float* x, int64_t nx, const float* y, int64_t ny, float r, int mode,
                       CellWs& w, cudaStream_t st) {
    bbox_init_kernel<<<1, 32, 0, st>>>(w.bb);
    GAOT_LAUNCH_CHECK();
    int gb = (int)nblk(nx, 256); if (gb > kNumSMs * 8) gb = kNumSMs * 8;
    bbox_kernel<<<gb, 256, 0, st>>>(x, nx, w.bb);
    GAOT_LAUNCH_CHECK();
    grid_params_kernel<<<1, 32, 0, st>>>(w.bb, nx, r, mode, w.gp);
    GAOT_LAUNCH_CHECK();
    // sources
    GAOT_CUDA(cudaMemsetAsync(w.cell_tmp, 0, (size_t)(MAX_CELLS + 1) * sizeof(int32_t), st));
    cell_hist_kernel<<<nblk(nx, 256), 256, 0, st>>>(x, nx, w.gp, w.cell_of_pt, w.cell_tmp);
    GAOT_LAUNCH_CHECK();
    int rc = exclusive_scan_i32(w.cell_tmp, w.cell_start, MAX_CELLS + 1, false, w.scan_ws, w.scan_bytes, st);
    if (rc) return rc;
    GAOT_CUDA(cudaMemsetAsync(w.cell_tmp, 0, (size_t)(MAX_CELLS + 1) * sizeof(int32_t), st));
    cell_scatter_src_kernel<<<nblk(nx, 256), 256, 0, st>>>(x, nx, w.cell_of_pt, w.cell_start, w.cell_tmp, w.pts);
    GAOT_LAUNCH_CHECK();
    // queries (same grid -> spatially coherent thread order)
    GAOT_CUDA(cudaMemsetAsync(w.cell_tmp, 0, (size_t)(MAX_CELLS + 1) * sizeof(int32_t), st));
    cell_hist_kernel<<<nblk(ny, 256), 256, 0, st>>>(y, ny, w.gp, w.cell_of_pt, w.cell_tmp);
    GAOT_LAUNCH_CHECK();
    rc = exclusive_scan_i32(w.cell_tmp, w.qcell_start, MAX_CELLS + 1, false, w.scan_ws, w.scan_bytes, st);
    if (rc) return rc;
    GAOT_CUDA(cudaMemsetAsync(w.cell_tmp, 0, (size_t)(MAX_CELLS + 1) * sizeof(int32_t), st));
    cell_scatter_perm_kernel<<<nblk(ny, 256), 256, 0, st>>>(ny, w.cell_of_pt, w.qcell_start, w.cell_tmp, w.qperm);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

static inline float r2_of(double r) { return (float)(r * r); }   // fl32(double(r)*double(r))

// ---------------------------------------------------------------- coalesce / mask / csr kernels
__global__ void __launch_bounds__(256)
pack_keys_kernel(const int64_t* __restrict__ r0, const int64_t* __restrict__ r1, int64_t n, int b1,
                 uint64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)r0[i] << b1) | (uint64_t)r1[i];
}
__global__ void __launch_bounds__(256)
unique_flag_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
unique_emit_kernel(const uint64_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ pos, int b1,
                   int64_t* __restrict__ o0, int64_t* __restrict__ o1, int64_t* __restrict__ e_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = (i == 0 || keys[i] != keys[i - 1]);
    if (head) {
        const uint64_t k = keys[i];
        o0[pos[i]] = (int64_t)(k >> b1);
        o1[pos[i]] = (int64_t)(k & (((uint64_t)1 << b1) - 1));
    }
    if (i == n - 1) *e_out = (int64_t)pos[i] + (head ? 1 : 0);
}

// Philox4x32-10 (counter-based): one 128-bit block per 4 edges
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t ctr, int lane4) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return (float)(c[lane4] >> 8) * (1.0f / 16777216.0f);     // [0,1), 24 bits like torch.rand(float32)
}
__global__ void __launch_bounds__(256)
mask_flag_kernel(int64_t n, float p, uint64_t seed, uint64_t offset, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = philox_uniform(seed, offset + (uint64_t)(i >> 2), (int)(i & 3)) >= p ? 1 : 0;
}
__global__ void __launch_bounds__(256)
mask_emit_kernel(const int64_t* __restrict__ r0, const int64_t* __restrict__ r1, int64_t n,
                 const int32_t* __restrict__ flag_pos /* exclusive scan of flags, [n+1] */,
                 int64_t* __restrict__ o0, int64_t* __restrict__ o1, int64_t* __restrict__ e_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag_pos[i + 1] != flag_pos[i]) { o0[flag_pos[i]] = r0[i]; o1[flag_pos[i]] = r1[i]; }
    if (i == n - 1) *e_out = flag_pos[n];
}

__global__ void __launch_bounds__(256)
csr_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ qry, int64_t n, int64_t n_src, int64_t nq,
                uint64_t* __restrict__ keys, int32_t* __restrict__ cnt, int32_t* __restrict__ status /* may be NULL: trusted */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t q = qry[i];
    if (status) {
        // edge lists that come from outside (precomputed edges, magno.py:506-516) are range-checked once here: every
        // later kernel gathers rows by these indices.  bit 0: query index out of [0,nq); bit 1: source index out of
        // [0,n_src); bit 2: claimed query-sorted but not monotone
        int bad = 0;
        if (q < 0 || q >= nq) bad |= 1;
        const int64_t s = src[i];
        if (s < 0 || s >= n_src) bad |= 2;
        if (!keys && i > 0 && qry[i - 1] > q) bad |= 4;
        if (bad) { atomicOr(status, bad); if (bad & 1) return; }
    }
    if (keys) keys[i] = ((uint64_t)q << 32) | (uint64_t)(uint32_t)i;
    atomicAdd(&cnt[q], 1);
}
__global__ void __launch_bounds__(256)
csr_gather_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ qry, int64_t n,
                  const uint64_t* __restrict__ keys /* may be NULL: identity */, int32_t* __restrict__ csr_src,
                  int32_t* __restrict__ csr_qry, int32_t* __restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t e = keys ? (int64_t)(uint32_t)(keys[i] & 0xffffffffu) : i;
    csr_src[i] = (int32_t)src[e];
    csr_qry[i] = (int32_t)qry[e];
    if (perm) perm[i] = (int32_t)e;
}

static inline int bits_for(int64_t maxval) { int b = 1; while (((int64_t)1 << b) <= maxval) ++b; return b; }

}  // namespace gaot

using namespace gaot;

// ================================================================ C ABI
// sources denser than queries (the encoder direction): one warp per query; GAOT_RADIUS_WARP=0/1 forces the choice
static bool radius_use_warp(int64_t nx, int64_t ny, int cap) {
    static const char* e = getenv("GAOT_RADIUS_WARP");
    if (cap > 32) return false;
    if (e) return atoi(e) != 0;
    return nx >= 2 * ny;
}

extern "C" {

size_t gaot_radius_workspace_bytes(int64_t nx, int64_t ny) { return cell_ws_bytes(nx, ny); }

int gaot_radius_count(const float* x, int64_t nx, const float* y, int64_t ny, double r, int cap,
                      void* ws, size_t ws_bytes, int32_t* rowptr, int64_t* E_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(nx >= 0 && ny >= 0 && nx < ((int64_t)1 << 31) && ny < ((int64_t)1 << 31) - 1, "radius: bad sizes");
    GAOT_CHECK_ARG(cap >= 1 && cap <= 128, "radius: max_num_neighbors must be in [1,128], got %d", cap);
    GAOT_CHECK_ARG(rowptr != nullptr, "radius: rowptr is null");
    if (nx == 0 || ny == 0) {
        GAOT_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(ny + 1) * sizeof(int32_t), st));
        if (E_host) *E_host = 0;
        return GAOT_OK;
    }
    CellWs w;
    if (!carve(w, ws, ws_bytes, nx, ny)) { set_error("radius: workspace too small"); return GAOT_ERR_WORKSPACE; }
    int rc = build_cells(x, nx, y, ny, (float)r, 0, w, st);
    if (rc) return rc;
    if (radius_use_warp(nx, ny, cap))
        radius_warp_kernel<false><<<nblk(ny * 32, 128), 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts, r2_of(r), cap,
                                                                       w.counts, nullptr, nullptr, nullptr);
    else
        radius_count_kernel<32><<<nblk(ny, 128), 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts,
                                                              r2_of(r), cap, w.counts);
    GAOT_LAUNCH_CHECK();
    rc = exclusive_scan_i32(w.counts, rowptr, ny, true, w.scan_ws, w.scan_bytes, st);
    if (rc) return rc;
    if (E_host) {
        int32_t e32 = 0;
        GAOT_CUDA(cudaMemcpyAsync(&e32, rowptr + ny, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GAOT_CUDA(cudaStreamSynchronize(st));
        *E_host = e32;
    }
    return GAOT_OK;
}

int gaot_radius_emit(const float* x, int64_t nx, const float* y, int64_t ny, double r, int cap,
                     void* ws, size_t ws_bytes, const int32_t* rowptr, int64_t* out_y, int64_t* out_x,
                     void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(cap >= 1 && cap <= 128, "radius: max_num_neighbors must be in [1,128]");
    if (nx == 0 || ny == 0) return GAOT_OK;
    CellWs w;
    if (!carve(w, ws, ws_bytes, nx, ny)) { set_error("radius: workspace too small"); return GAOT_ERR_WORKSPACE; }
    if (radius_use_warp(nx, ny, cap))
        radius_warp_kernel<true><<<nblk(ny * 32, 128), 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts, r2_of(r), cap,
                                                                      nullptr, rowptr, out_y, out_x);
    else if (cap <= 32)
        radius_emit_kernel<32><<<nblk(ny, 128), 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts,
                                                             r2_of(r), cap, rowptr, out_y, out_x);
    else
        radius_emit_kernel<128><<<nblk(ny, 128), 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts,
                                                              r2_of(r), cap, rowptr, out_y, out_x);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

size_t gaot_knn_workspace_bytes(int64_t nx, int64_t ny) { return cell_ws_bytes(nx, ny); }

int gaot_knn(const float* x, int64_t nx, const float* y, int64_t ny, int k, void* ws, size_t ws_bytes,
             int64_t* out_y, int64_t* out_x, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(nx >= 0 && ny >= 0 && nx < ((int64_t)1 << 31) && ny < ((int64_t)1 << 31), "knn: bad sizes");
    GAOT_CHECK_ARG(k >= 1 && k <= 128, "knn: k must be in [1,128], got %d", k);
    GAOT_CHECK_ARG(k <= nx || nx == 0, "knn: host must clamp k to nx");
    if (nx == 0 || ny == 0) return GAOT_OK;
    CellWs w;
    if (!carve(w, ws, ws_bytes, nx, ny)) { set_error("knn: workspace too small"); return GAOT_ERR_WORKSPACE; }
    int rc = build_cells(x, nx, y, ny, 0.f, 1, w, st);
    if (rc) return rc;
    const unsigned nb = nblk(ny, 128);
    GAOT_TIME_KERNEL("knn_search", st, 24.0 * (double)(nx + ny) + 8.0 * (double)(nx + ny) + 16.0 * (double)ny * k);
    if (k == 1)       knn_kernel<1><<<nb, 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts, k, out_y, out_x);
    else if (k <= 8)  knn_kernel<8><<<nb, 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts, k, out_y, out_x);
    else if (k <= 32) knn_kernel<32><<<nb, 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts, k, out_y, out_x);
    else              knn_kernel<128><<<nb, 128, 0, st>>>(y, ny, w.qperm, w.gp, w.cell_start, w.pts, k, out_y, out_x);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

size_t gaot_coalesce_workspace_bytes(int64_t E) {
    if (E < 1) E = 1;
    return 2 * align_up((size_t)E * sizeof(uint64_t)) + align_up((size_t)(E + 1) * sizeof(int32_t)) +
           align_up(sort_workspace_bytes(E)) + align_up(scan_workspace_bytes(E + 1)) + 1024;
}

int gaot_coalesce(const int64_t* row0, const int64_t* row1, int64_t E, int64_t max_row0, int64_t max_row1,
                  void* ws, size_t ws_bytes, int64_t* out0, int64_t* out1, int64_t* E_out_dev,
                  int64_t* E_out_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(E >= 0 && E < ((int64_t)1 << 31) - 1, "coalesce: bad E");
    GAOT_CHECK_ARG(E_out_dev != nullptr, "coalesce: E_out_dev is null");
    if (E == 0) {
        GAOT_CUDA(cudaMemsetAsync(E_out_dev, 0, sizeof(int64_t), st));
        if (E_out_host) *E_out_host = 0;
        return GAOT_OK;
    }
    const int b1 = bits_for(max_row1), b0 = bits_for(max_row0);
    GAOT_CHECK_ARG(b0 + b1 <= 64, "coalesce: index range too large");
    Arena ar(ws, ws_bytes);
    uint64_t* keys = ar.take<uint64_t>((size_t)E);
    uint64_t* tmp = ar.take<uint64_t>((size_t)E);
    int32_t* flag = ar.take<int32_t>((size_t)E + 1);
    const size_t sb = sort_workspace_bytes(E), cb = scan_workspace_bytes(E + 1);
    char* sort_ws = ar.take<char>(sb);
    char* scan_ws = ar.take<char>(cb);
    if (!ar.ok()) { set_error("coalesce: workspace too small"); return GAOT_ERR_WORKSPACE; }
    pack_keys_kernel<<<nblk(E, 256), 256, 0, st>>>(row0, row1, E, b1, keys);
    GAOT_LAUNCH_CHECK();
    int rc = radix_sort_u64(keys, tmp, E, 0, b0 + b1, sort_ws, sb, st);
    if (rc) return rc;
    unique_flag_kernel<<<nblk(E, 256), 256, 0, st>>>(keys, E, flag);
    GAOT_LAUNCH_CHECK();
    rc = exclusive_scan_i32(flag, flag, E, false, scan_ws, cb, st);
    if (rc) return rc;
    unique_emit_kernel<<<nblk(E, 256), 256, 0, st>>>(keys, E, flag, b1, out0, out1, E_out_dev);
    GAOT_LAUNCH_CHECK();
    if (E_out_host) {
        GAOT_CUDA(cudaMemcpyAsync(E_out_host, E_out_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        GAOT_CUDA(cudaStreamSynchronize(st));
    }
    return GAOT_OK;
}

size_t gaot_edge_mask_workspace_bytes(int64_t E) {
    if (E < 1) E = 1;
    return align_up((size_t)(E + 1) * sizeof(int32_t)) + align_up(scan_workspace_bytes(E + 1)) + 1024;
}

int gaot_edge_mask(const int64_t* row0, const int64_t* row1, int64_t E, double p_drop, uint64_t seed,
                   uint64_t offset, void* ws, size_t ws_bytes, int64_t* out0, int64_t* out1,
                   int64_t* E_out_dev, int64_t* E_out_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(E >= 0 && E < ((int64_t)1 << 31) - 1, "edge_mask: bad E");
    GAOT_CHECK_ARG(p_drop >= 0.0 && p_drop <= 1.0, "edge_mask: p must be in [0,1]");
    GAOT_CHECK_ARG(E_out_dev != nullptr, "edge_mask: E_out_dev is null");
    if (E == 0) {
        GAOT_CUDA(cudaMemsetAsync(E_out_dev, 0, sizeof(int64_t), st));
        if (E_out_host) *E_out_host = 0;
        return GAOT_OK;
    }
    Arena ar(ws, ws_bytes);
    int32_t* flag = ar.take<int32_t>((size_t)E + 1);
    const size_t cb = scan_workspace_bytes(E + 1);
    char* scan_ws = ar.take<char>(cb);
    if (!ar.ok()) { set_error("edge_mask: workspace too small"); return GAOT_ERR_WORKSPACE; }
    mask_flag_kernel<<<nblk(E, 256), 256, 0, st>>>(E, (float)p_drop, seed, offset, flag);
    GAOT_LAUNCH_CHECK();
    int rc = exclusive_scan_i32(flag, flag, E, true, scan_ws, cb, st);
    if (rc) return rc;
    mask_emit_kernel<<<nblk(E, 256), 256, 0, st>>>(row0, row1, E, flag, out0, out1, E_out_dev);
    GAOT_LAUNCH_CHECK();
    if (E_out_host) {
        GAOT_CUDA(cudaMemcpyAsync(E_out_host, E_out_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        GAOT_CUDA(cudaStreamSynchronize(st));
    }
    return GAOT_OK;
}

size_t gaot_csr_workspace_bytes(int64_t E, int64_t nq) {
    if (E < 1) E = 1;
    return 2 * align_up((size_t)E * sizeof(uint64_t)) + align_up(sort_workspace_bytes(E)) +
           align_up(scan_workspace_bytes(nq + 1)) + align_up(sizeof(int32_t)) + 1024;
}

int gaot_csr_from_edges(const int64_t* src, const int64_t* qry, int64_t E, int64_t n_src, int64_t nq,
                        int flags, void* ws, size_t ws_bytes, int32_t* rowptr, int32_t* csr_src,
                        int32_t* csr_qry, int32_t* perm, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(E >= 0 && E < ((int64_t)1 << 31) - 1, "csr: bad E");
    GAOT_CHECK_ARG(nq >= 0 && nq < ((int64_t)1 << 31) - 1 && n_src < ((int64_t)1 << 31), "csr: bad sizes");
    GAOT_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(nq + 1) * sizeof(int32_t), st));
    if (E == 0) return GAOT_OK;
    const bool sorted = (flags & 1) != 0, validate = (flags & 2) != 0;
    Arena ar(ws, ws_bytes);
    uint64_t* keys = ar.take<uint64_t>((size_t)E);
    uint64_t* tmp = ar.take<uint64_t>((size_t)E);
    const size_t sb = sort_workspace_bytes(E), cb = scan_workspace_bytes(nq + 1);
    char* sort_ws = ar.take<char>(sb);
    char* scan_ws = ar.take<char>(cb);
    int32_t* status = ar.take<int32_t>(1);
    if (!ar.ok()) { set_error("csr: workspace too small"); return GAOT_ERR_WORKSPACE; }
    if (validate) GAOT_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    csr_keys_kernel<<<nblk(E, 256), 256, 0, st>>>(src, qry, E, n_src, nq, sorted ? nullptr : keys, rowptr,
                                                  validate ? status : nullptr);
    GAOT_LAUNCH_CHECK();
    if (validate) {      // one 4-byte read-back per externally supplied edge list, before anything gathers by its indices
        int32_t h = 0;
        GAOT_CUDA(cudaMemcpyAsync(&h, status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GAOT_CUDA(cudaStreamSynchronize(st));
        if (h) {
            set_error("csr: edge_index is invalid for n_src=%lld, nq=%lld:%s%s%s", (long long)n_src, (long long)nq,
                      (h & 1) ? " query index (row 1) out of range;" : "", (h & 2) ? " source index (row 0) out of range;" : "",
                      (h & 4) ? " marked query-sorted but row 1 is not monotone;" : "");
            return GAOT_ERR_INVALID;
        }
    }
    int rc = exclusive_scan_i32(rowptr, rowptr, nq, true, scan_ws, cb, st);
    if (rc) return rc;
    if (!sorted) {
        rc = radix_sort_u64(keys, tmp, E, 32, 32 + bits_for(nq > 0 ? nq - 1 : 0), sort_ws, sb, st);
        if (rc) return rc;
    }
    csr_gather_kernel<<<nblk(E, 256), 256, 0, st>>>(src, qry, E, sorted ? nullptr : keys, csr_src, csr_qry, perm);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // extern "C"
