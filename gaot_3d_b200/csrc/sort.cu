// sort.cu -- stable LSD radix sort of uint64 keys (8-bit digits), used by the
// coalesce step (reference magno.py:220,293 -> torch_geometric.utils.coalesce) and by the
// query-major CSR side-band of arbitrarily ordered edge lists.
//
// Per pass: (1) per-block digit histogram -> hist[digit][block]; (2) exclusive scan of the
// digit-major table; (3) stable scatter -- ranks inside a block come from warp match-any
// ballots + a per-digit prefix over the block's warps, so equal digits keep input order.
#include "common.cuh"

namespace gaot {

constexpr int SORT_THREADS = 256;
constexpr int SORT_CHUNKS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_CHUNKS;   // keys per block
constexpr int SORT_WARPS = SORT_THREADS / 32;

__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist,
                 int nblocks) {
    __shared__ int32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int c = 0; c < SORT_CHUNKS; ++c) {
        const int64_t i = base + (int64_t)c * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t* __restrict__ out, int64_t n, int shift,
                    const int32_t* __restrict__ offs, int nblocks) {
    __shared__ int32_t warp_cnt[SORT_WARPS][256];
    __shared__ int32_t base_off[256];      // global offset of this block's next key per digit
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    base_off[tid] = offs[(int64_t)tid * nblocks + blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    for (int c = 0; c < SORT_CHUNKS; ++c) {
        const int64_t cbase = base + (int64_t)c * SORT_THREADS;
        if (cbase >= n) break;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) warp_cnt[w][tid] = 0;
        __syncthreads();
        const int64_t i = cbase + tid;
        const bool valid = i < n;
        uint64_t key = valid ? keys[i] : 0;
        const int digit = valid ? (int)((key >> shift) & 0xff) : 256 + lane;   // invalid lanes never match
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_cnt[warp][digit] = __popc(peers);
        __syncthreads();
        // thread `tid` owns digit `tid`: exclusive prefix over warps, then advance the base
        int32_t run = base_off[tid];
        const int32_t start = run;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const int32_t cnt = warp_cnt[w][tid];
            warp_cnt[w][tid] = run - start;       // offset of warp w inside this chunk for digit tid
            run += cnt;
        }
        __syncthreads();
        if (valid) out[(int64_t)base_off[digit] + warp_cnt[warp][digit] + rank_in_warp] = key;
        __syncthreads();
        base_off[tid] = run;
        // (next iteration's zeroing of warp_cnt happens after this barrier-protected read)
        __syncthreads();
    }
}

static inline int64_t sort_blocks(int64_t n) { int64_t b = (n + SORT_TILE - 1) / SORT_TILE; return b < 1 ? 1 : b; }

size_t sort_workspace_bytes(int64_t n) {
    const int64_t nb = sort_blocks(n);
    return align_up((size_t)(256 * nb + 1) * sizeof(int32_t)) + scan_workspace_bytes(256 * nb) + 256;
}

int radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int bit_lo, int bit_hi,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
    if (n <= 0) return GAOT_OK;
    if (n >= (int64_t)1 << 31) { set_error("radix_sort: n too large"); return GAOT_ERR_UNSUPPORTED; }
    const int64_t nb = sort_blocks(n);
    Arena ar(ws, ws_bytes);
    int32_t* hist = ar.take<int32_t>((size_t)256 * nb + 1);
    const size_t scan_bytes = scan_workspace_bytes(256 * nb);
    char* scan_ws = ar.take<char>(scan_bytes);
    if (!hist || !scan_ws) { set_error("radix_sort: workspace too small"); return GAOT_ERR_WORKSPACE; }
    uint64_t* src = keys; uint64_t* dst = tmp;
    for (int shift = bit_lo; shift < bit_hi; shift += 8) {
        sort_hist_kernel<<<(unsigned)nb, SORT_THREADS, 0, st>>>(src, n, shift, hist, (int)nb);
        GAOT_LAUNCH_CHECK();
        int rc = exclusive_scan_i32(hist, hist, 256 * nb, false, scan_ws, scan_bytes, st);
        if (rc) return rc;
        sort_scatter_kernel<<<(unsigned)nb, SORT_THREADS, 0, st>>>(src, dst, n, shift, hist, (int)nb);
        GAOT_LAUNCH_CHECK();
        uint64_t* t = src; src = dst; dst = t;
    }
    if (src != keys) GAOT_CUDA(cudaMemcpyAsync(keys, src, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return GAOT_OK;
}

}  // namespace gaot
