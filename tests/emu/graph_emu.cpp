// graph_emu.cpp -- DEVELOPMENT/TEST AID ONLY (never linked into libgaot_b200.so).
// Drives the exact __host__ __device__ per-query bodies of gaot_3d_b200/csrc/graph_core.cuh
// from a plain C++ loop, so the cell-list logic (grid sizing, clamping, ring expansion,
// index-ordered cap, tie rules) can be checked against the oracle on a box without a GPU.
// The in-cell order of sources is deliberately shuffled: on the GPU it is the (arbitrary)
// completion order of atomics and results must not depend on it.
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include "../../gaot_3d_b200/csrc/graph_core.cuh"

using namespace gaot;

struct Cells { GridParams g; std::vector<int> start; std::vector<SrcPoint> pts; };

static Cells build(const float* x, int64_t nx, float r, int mode, int max_cells, int max_dim, unsigned seed) {
    Cells c;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int64_t i = 0; i < nx; ++i)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], x[i * 3 + a]); hi[a] = std::max(hi[a], x[i * 3 + a]); }
    c.g = compute_grid_params(lo, hi, nx, r, mode, max_cells, max_dim);
    std::vector<int> cell(nx);
    c.start.assign((size_t)c.g.ncells + 2, 0);
    for (int64_t i = 0; i < nx; ++i) { cell[i] = cell_of(c.g, x[i * 3], x[i * 3 + 1], x[i * 3 + 2]); c.start[cell[i] + 1]++; }
    for (int i = 0; i <= c.g.ncells; ++i) c.start[i + 1] += c.start[i];
    std::vector<int64_t> order(nx);
    for (int64_t i = 0; i < nx; ++i) order[i] = i;
    std::mt19937 rng(seed);
    std::shuffle(order.begin(), order.end(), rng);
    std::vector<int> fill(c.g.ncells + 1, 0);
    c.pts.resize(nx);
    for (int64_t oi = 0; oi < nx; ++oi) {
        int64_t i = order[oi];
        SrcPoint s; s.x = x[i * 3]; s.y = x[i * 3 + 1]; s.z = x[i * 3 + 2]; s.idx = (int)i;
        c.pts[c.start[cell[i]] + fill[cell[i]]++] = s;
    }
    return c;
}

extern "C" {

// out_y/out_x sized ny*cap; returns E
int64_t emu_radius(const float* x, int64_t nx, const float* y, int64_t ny, double r, int cap,
                   int max_cells, int max_dim, int64_t* out_y, int64_t* out_x) {
    if (nx == 0 || ny == 0) return 0;
    Cells c = build(x, nx, (float)r, 0, max_cells, max_dim, 123);
    const float r2 = (float)(r * r);
    std::vector<int> list(cap > 128 ? cap : 128);
    int64_t E = 0;
    for (int64_t q = 0; q < ny; ++q) {
        int cnt = radius_query<false>(c.g, c.start.data(), c.pts.data(), y[q * 3], y[q * 3 + 1], y[q * 3 + 2], r2, cap, nullptr);
        int n = radius_query<true>(c.g, c.start.data(), c.pts.data(), y[q * 3], y[q * 3 + 1], y[q * 3 + 2], r2, cap, list.data());
        if (cnt != n) return -1;
        for (int j = 0; j < n; ++j) { out_y[E] = q; out_x[E] = list[j]; ++E; }
    }
    return E;
}

int64_t emu_knn(const float* x, int64_t nx, const float* y, int64_t ny, int k, int max_cells, int max_dim,
                int64_t* out_y, int64_t* out_x) {
    if (nx == 0 || ny == 0) return 0;
    Cells c = build(x, nx, 0.f, 1, max_cells, max_dim, 321);
    std::vector<float> bd(k); std::vector<int> bi(k);
    int64_t E = 0;
    for (int64_t q = 0; q < ny; ++q) {
        int n = knn_query<128>(c.g, c.start.data(), c.pts.data(), y[q * 3], y[q * 3 + 1], y[q * 3 + 2], k, bd.data(), bi.data());
        for (int j = 0; j < n; ++j) { out_y[E] = q; out_x[E] = bi[j]; ++E; }
    }
    return E;
}

void emu_grid(const float* x, int64_t nx, double r, int mode, int max_cells, int max_dim, int* dims, float* h) {
    Cells c = build(x, nx, (float)r, mode, max_cells, max_dim, 1);
    dims[0] = c.g.nx; dims[1] = c.g.ny; dims[2] = c.g.nz; dims[3] = c.g.reach; *h = c.g.h;
}
}
