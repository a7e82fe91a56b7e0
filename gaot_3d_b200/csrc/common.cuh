// common.cuh -- shared helpers for libgaot_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/gaot_b200.h"

namespace gaot {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GAOT_CHECK_ARG(cond, ...)                                   \
    do { if (!(cond)) { gaot::set_error(__VA_ARGS__); return GAOT_ERR_INVALID; } } while (0)

#define GAOT_CUDA(call)                                                                   \
    do { cudaError_t _e = (call); if (_e != cudaSuccess) {                                \
        gaot::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,             \
                        cudaGetErrorString(_e)); return GAOT_ERR_CUDA; } } while (0)

#define GAOT_LAUNCH_CHECK()                                                               \
    do { gaot::count_launch(); cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
        gaot::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,             \
                        cudaGetErrorString(_e)); return GAOT_ERR_CUDA; } } while (0)

// Optional device-side timing of the main kernels (bench.py roofline): CUDA events recorded on
// the launching stream around the launch, read back by gaot_profile_summary().
struct KernelTimer {
    const char* name; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; double work;
    KernelTimer(const char* name, cudaStream_t st, double work = 0.0);
    ~KernelTimer();
};
#define GAOT_TIME_KERNEL(name, st, work) gaot::KernelTimer _gaot_kt(name, st, work)

constexpr int kNumSMs = 148;   // B200

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Arena {
    char* base; size_t cap; size_t off;
    Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
    template <typename T> T* take(size_t n) {
        size_t b = align_up(n * sizeof(T));
        if (off + b > cap) { off = cap + 1; return nullptr; }
        T* r = (T*)(base + off); off += b; return r;
    }
    bool ok() const { return off <= cap; }
};

// ---- device scans (scan.cu) ----
// exclusive scan of n int32 values; out[n] (if write_total) receives the total.  in may alias out.
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, bool write_total,
                       void* ws, size_t ws_bytes, cudaStream_t st);

// ---- stable LSD radix sort of uint64 keys on bits [bit_lo, bit_hi) (sort.cu) ----
size_t sort_workspace_bytes(int64_t n);
// result ends up in `keys` (ping-pongs with `tmp` internally; both size n)
int radix_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int bit_lo, int bit_hi,
                   void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace gaot
