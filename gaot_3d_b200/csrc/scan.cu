// scan.cu -- multi-level exclusive prefix sum (int32), the "prefix-sum CSR edge emitter"
// building block used by the graph build (cell starts, rowptr, compaction offsets).
#include "common.cuh"

namespace gaot {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                  int32_t* __restrict__ sums) {
    __shared__ int32_t warp_tot[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        tsum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_tot[lane];
        int32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;           // exclusive warp offsets
        if (lane == 31) sums[blockIdx.x] = winc;
    }
    __syncthreads();
    int32_t run = warp_tot[warp] + inc - tsum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ sums) {
    const int32_t add = sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) out[base + i] += add;
}

__global__ void scan_total_kernel(int32_t* dst, const int32_t* src) { *dst = *src; }

size_t scan_workspace_bytes(int64_t n) {
    size_t tot = 0;
    while (true) {
        int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
        if (nb < 1) nb = 1;
        tot += align_up((size_t)(nb + 1) * sizeof(int32_t));
        if (nb <= 1) break;
        n = nb;
    }
    return tot + 256;
}

static int scan_rec(const int32_t* in, int32_t* out, int64_t n, bool write_total, Arena& ar, cudaStream_t st) {
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb < 1) nb = 1;
    int32_t* sums = ar.take<int32_t>((size_t)nb + 1);
    if (!sums) { set_error("scan: workspace too small"); return GAOT_ERR_WORKSPACE; }
    scan_block_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums);
    GAOT_LAUNCH_CHECK();
    if (nb > 1) {
        int rc = scan_rec(sums, sums, nb, true, ar, st);
        if (rc) return rc;
        scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(out, n, sums);
        GAOT_LAUNCH_CHECK();
    }
    if (write_total) {
        scan_total_kernel<<<1, 1, 0, st>>>(out + n, nb > 1 ? sums + nb : sums);
        GAOT_LAUNCH_CHECK();
    }
    return GAOT_OK;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, bool write_total,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
    Arena ar(ws, ws_bytes);
    return scan_rec(in, out, n, write_total, ar, st);
}

}  // namespace gaot
