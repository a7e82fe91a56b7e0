// tmem_probe.cu -- single-SM microbenchmarks behind the attention kernels' roofline model (B200, sm_100a):
//   A: tcgen05.ld 32x32b.x32 throughput with W warps          -> TMEM read bytes / clk / SM
//   B: MUFU.EX2 throughput with W warps                        -> ex2 / clk / SM
//   C: the softmax inner step (LDTM.x32 -> 32 ex2 -> 16 packs -> STTM.x16) with W warps -> clk per 128x128 tile
//   D: C without the exponentials (LDTM + pack + STTM)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gaot_3d_b200/csrc profiles/tools/tmem_probe.cu -o gpurun_out/tmem_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace gaot;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* out, float* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tc::tmem_alloc(&tbase, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 32u;
    float acc = 0.f;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(-(float)(i + lane) * 0.01f);
    if (MODE != 1) { tc::tmem_st32(tl, r); tc::tmem_wait_st(); }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {                       // LDTM only, two loads in flight
            uint32_t a[32], b[32];
            tc::tmem_ld32_nowait(tl, a);
            tc::tmem_ld32_nowait(tl, b);
            tc::tmem_wait_ld();
            acc += __uint_as_float(a[it & 31]) + __uint_as_float(b[(it + 7) & 31]);
        } else if (MODE == 1) {                // MUFU only
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(ex2a(__uint_as_float(r[i]) * 0.5f));
        } else {                               // softmax step
            float v[32];
            tc::tmem_ld32(tl, v);
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float p0 = v[c] + acc, p1 = v[c + 1] + acc;
                if (MODE == 2) { p0 = ex2a(p0); p1 = ex2a(p1); }
                pk[c >> 1] = tc::pack_bf16(p0, p1);
            }
            tc::tmem_st16(tl, pk);
            tc::tmem_wait_st();
            acc += 1e-9f;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (MODE == 1) { for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]); }
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

int main() {
    long long* d_out; float* d_sink; long long h;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4096);
    const int iters = 2000;
    const char* names[4] = {"A LDTM.x32 x2", "B MUFU.EX2 x32", "C LDTM+32 ex2+pack+STTM.x16", "D LDTM+pack+STTM.x16"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<1, warps * 32>>>(iters, d_out, d_sink);
                if (mode == 1) probe<1><<<1, warps * 32>>>(iters, d_out, d_sink);
                if (mode == 2) probe<2><<<1, warps * 32>>>(iters, d_out, d_sink);
                if (mode == 3) probe<3><<<1, warps * 32>>>(iters, d_out, d_sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            const double clk = (double)h / iters;
            if (mode == 0) printf("%-32s warps %2d: %8.1f clk/iter  -> %7.1f B/clk/SM TMEM read\n", names[mode], warps, clk, warps * 2 * 4096.0 / clk);
            else if (mode == 1) printf("%-32s warps %2d: %8.1f clk/iter  -> %7.2f ex2/clk/SM\n", names[mode], warps, clk, warps * 32 * 32.0 / clk);
            else printf("%-32s warps %2d: %8.1f clk/iter  -> %7.1f clk per 128x128 tile (16 warp-iterations)\n", names[mode], warps, clk, clk * 16.0 / warps);
        }
    return 0;
}
