// abi.cu -- error/diagnostic plumbing of the C ABI and the host-buffer convenience entry points.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <vector>
#include <map>
#include <string>

namespace gaot {
static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProfRec { const char* name; cudaEvent_t e0, e1; double work; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

KernelTimer::KernelTimer(const char* n, cudaStream_t s, double w) : name(n), st(s), work(w) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
    cudaEventRecord(e0, st);
}
KernelTimer::~KernelTimer() {
    if (!e0) return;
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({name, e0, e1, work});
}
}  // namespace gaot

using namespace gaot;

extern "C" {

const char* gaot_last_error(void) { return g_err; }
int gaot_abi_version(void) { return 1; }
int64_t gaot_launch_count(void) { return g_launches.load(); }
void gaot_launch_count_reset(void) { g_launches.store(0); }
void gaot_launch_count_add(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void gaot_profile_enable(int on) {
    g_prof_on.store(on ? 1 : 0);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
}

// "name calls total_ms total_work\n" per kernel; synchronises on the recorded events.
int gaot_profile_summary(char* buf, size_t buf_bytes) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    struct Acc { int64_t calls = 0; double ms = 0, work = 0; };
    std::map<std::string, Acc> acc;
    for (auto& r : g_prof) {
        if (cudaEventSynchronize(r.e1) != cudaSuccess) { set_error("profile: event sync failed"); return GAOT_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        Acc& a = acc[r.name]; a.calls++; a.ms += ms; a.work += r.work;
    }
    std::string out;
    char line[256];
    for (auto& kv : acc) {
        snprintf(line, sizeof(line), "%s %lld %.6f %.6e\n", kv.first.c_str(), (long long)kv.second.calls, kv.second.ms, kv.second.work);
        out += line;
    }
    if (out.size() + 1 > buf_bytes) { set_error("profile: buffer too small"); return GAOT_ERR_INVALID; }
    memcpy(buf, out.c_str(), out.size() + 1);
    return GAOT_OK;
}

// Host-buffer graph build: H2D copy of positions, kernels, D2H copy of the edge list.
int gaot_radius_host(const float* x_host, int64_t nx, const float* y_host, int64_t ny, double r, int cap,
                     int64_t* out_y_host, int64_t* out_x_host, int64_t* E_host) {
    GAOT_CHECK_ARG(E_host != nullptr, "radius_host: E_host is null");
    *E_host = 0;
    if (nx == 0 || ny == 0) return GAOT_OK;
    cudaStream_t st = nullptr;
    float *dx = nullptr, *dy = nullptr; void* ws = nullptr; int32_t* rowptr = nullptr;
    int64_t *oy = nullptr, *ox = nullptr;
    const size_t wsb = gaot_radius_workspace_bytes(nx, ny);
    int rc = GAOT_OK;
    GAOT_CUDA(cudaMalloc(&dx, (size_t)nx * 12)); GAOT_CUDA(cudaMalloc(&dy, (size_t)ny * 12));
    GAOT_CUDA(cudaMalloc(&ws, wsb)); GAOT_CUDA(cudaMalloc(&rowptr, (size_t)(ny + 1) * 4));
    GAOT_CUDA(cudaMemcpyAsync(dx, x_host, (size_t)nx * 12, cudaMemcpyHostToDevice, st));
    GAOT_CUDA(cudaMemcpyAsync(dy, y_host, (size_t)ny * 12, cudaMemcpyHostToDevice, st));
    rc = gaot_radius_count(dx, nx, dy, ny, r, cap, ws, wsb, rowptr, E_host, st);
    if (rc == GAOT_OK && *E_host > 0) {
        GAOT_CUDA(cudaMalloc(&oy, (size_t)*E_host * 8)); GAOT_CUDA(cudaMalloc(&ox, (size_t)*E_host * 8));
        rc = gaot_radius_emit(dx, nx, dy, ny, r, cap, ws, wsb, rowptr, oy, ox, st);
        if (rc == GAOT_OK) {
            GAOT_CUDA(cudaMemcpyAsync(out_y_host, oy, (size_t)*E_host * 8, cudaMemcpyDeviceToHost, st));
            GAOT_CUDA(cudaMemcpyAsync(out_x_host, ox, (size_t)*E_host * 8, cudaMemcpyDeviceToHost, st));
            GAOT_CUDA(cudaStreamSynchronize(st));
        }
    }
    cudaFree(dx); cudaFree(dy); cudaFree(ws); cudaFree(rowptr); cudaFree(oy); cudaFree(ox);
    return rc;
}

int gaot_knn_host(const float* x_host, int64_t nx, const float* y_host, int64_t ny, int k,
                  int64_t* out_y_host, int64_t* out_x_host, int64_t* E_host) {
    GAOT_CHECK_ARG(E_host != nullptr, "knn_host: E_host is null");
    *E_host = 0;
    if (nx == 0 || ny == 0) return GAOT_OK;
    if (k > nx) k = (int)nx;
    cudaStream_t st = nullptr;
    float *dx = nullptr, *dy = nullptr; void* ws = nullptr; int64_t *oy = nullptr, *ox = nullptr;
    const size_t wsb = gaot_knn_workspace_bytes(nx, ny);
    const size_t ob = (size_t)ny * k * 8;
    GAOT_CUDA(cudaMalloc(&dx, (size_t)nx * 12)); GAOT_CUDA(cudaMalloc(&dy, (size_t)ny * 12));
    GAOT_CUDA(cudaMalloc(&ws, wsb)); GAOT_CUDA(cudaMalloc(&oy, ob)); GAOT_CUDA(cudaMalloc(&ox, ob));
    GAOT_CUDA(cudaMemcpyAsync(dx, x_host, (size_t)nx * 12, cudaMemcpyHostToDevice, st));
    GAOT_CUDA(cudaMemcpyAsync(dy, y_host, (size_t)ny * 12, cudaMemcpyHostToDevice, st));
    int rc = gaot_knn(dx, nx, dy, ny, k, ws, wsb, oy, ox, st);
    if (rc == GAOT_OK) {
        GAOT_CUDA(cudaMemcpyAsync(out_y_host, oy, ob, cudaMemcpyDeviceToHost, st));
        GAOT_CUDA(cudaMemcpyAsync(out_x_host, ox, ob, cudaMemcpyDeviceToHost, st));
        GAOT_CUDA(cudaStreamSynchronize(st));
        *E_host = ny * k;
    }
    cudaFree(dx); cudaFree(dy); cudaFree(ws); cudaFree(oy); cudaFree(ox);
    return rc;
}

}  // extern "C"
