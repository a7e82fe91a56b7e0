// tc_probe.cu -- TEST INFRASTRUCTURE (built into tests/probe/libgaot_tcprobe.so, NOT part of libgaot_b200.so):
// single-CTA tcgen05 GEMM used by tests/test_gpu_tc_probe.py to pin the
// shared-memory descriptor conventions of tc05.cuh on real hardware:
//   D[128,N] = A * B  (bf16 operands, fp32 accumulate in TMEM), for every combination of
//   K-major / MN-major A and B operands built from the library's chunk-major tile format.
#include <cstdarg>
#include "common.cuh"
#include "tc05.cuh"

// standalone: the two hooks of common.cuh that the product library defines in abi.cu
namespace gaot {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void count_launch(int) {}
}  // namespace gaot

namespace gaot {

__global__ void __launch_bounds__(128, 1)
tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                int variant) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool b_mn = variant & 1, a_mn = variant & 2, swap = variant & 4;
    uint8_t* sA = sm;                       // 128 x K bf16
    uint8_t* sB = sm + 128 * K * 2;         // N x K bf16
    const uint32_t rowsA = a_mn ? K : 128, colsA = a_mn ? 128 : K;
    const uint32_t rowsB = b_mn ? K : N, colsB = b_mn ? N : K;
    for (uint32_t idx = tid; idx < rowsA * colsA; idx += 128) {
        const uint32_t r = idx / colsA, c = idx % colsA;
        *reinterpret_cast<__nv_bfloat16*>(sA + tc::cm_off(rowsA, r, c)) = __float2bfloat16(A[idx]);
    }
    for (uint32_t idx = tid; idx < rowsB * colsB; idx += 128) {
        const uint32_t r = idx / colsB, c = idx % colsB;
        *reinterpret_cast<__nv_bfloat16*>(sB + tc::cm_off(rowsB, r, c)) = __float2bfloat16(B[idx]);
    }
    uint32_t ncols = 32; while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
        const uint32_t sa = tc::smem_u32(sA), sb = tc::smem_u32(sB);
        for (int s = 0; s < K / 16; ++s) {
            uint32_t a_start, a_lbo, a_sbo, b_start, b_lbo, b_sbo;
            if (a_mn) { a_start = sa + s * 256u; a_lbo = 128u; a_sbo = rowsA * 16u; }
            else      { a_start = sa + s * 2u * rowsA * 16u; a_lbo = rowsA * 16u; a_sbo = 128u; }
            if (b_mn) { b_start = sb + s * 256u; b_lbo = 128u; b_sbo = rowsB * 16u; }
            else      { b_start = sb + s * 2u * rowsB * 16u; b_lbo = rowsB * 16u; b_sbo = 128u; }
            if (swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
            tc::mma_bf16(tmem, tc::make_desc(a_start, a_lbo, a_sbo), tc::make_desc(b_start, b_lbo, b_sbo), idesc, s > 0);
        }
        tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace gaot

using namespace gaot;

extern "C" int gaot_tc_probe(const float* A, const float* B, float* D, int N, int K, int variant, void* stream) {
    GAOT_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "tc_probe: N must be a multiple of 16 in [16,256]");
    GAOT_CHECK_ARG(K >= 16 && K <= 256 && K % 16 == 0, "tc_probe: K must be a multiple of 16 in [16,256]");
    const size_t smem = (size_t)(128 + N) * K * 2;
    GAOT_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, variant);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}
