// graph_core.cuh -- per-query bodies of the cell-list radius / kNN search.
//
// Written as __host__ __device__ functions over plain pointers so that the exact same
// code is (a) launched one-thread-per-query by graph.cu and (b) driven by a plain C++
// loop in tests/emu/graph_emu.cpp (development aid, never shipped) to check the integer
// / compare logic against the oracle without a GPU.
//
// Semantics follow torch_cluster's CUDA kernels as called by the reference at
// src/model/layers/magno.py:183-200, :242-260:
//   d2 = ((dx*dx + dy*dy) + dz*dz) in fp32 WITHOUT fma contraction;
//   radius: d2 < fl32(r*r), first `cap` by ascending source index;
//   knn: ascending (d2, index).
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define GAOT_HD __host__ __device__ __forceinline__
#else
#define GAOT_HD inline
#endif

namespace gaot {

struct GridParams {
    float ox, oy, oz;     // grid origin (min corner of the source bounding box)
    float h, inv_h;       // cubic cell edge
    int nx, ny, nz;       // cells per axis
    int reach;            // cells to visit on each side for a radius query
    int ncells;
};

struct alignas(16) SrcPoint { float x, y, z; int idx; };

GAOT_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
GAOT_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
GAOT_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
GAOT_HD float dist2(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = f_sub(ax, bx), dy = f_sub(ay, by), dz = f_sub(az, bz);
    return f_add(f_add(f_mul(dx, dx), f_mul(dy, dy)), f_mul(dz, dz));
}

GAOT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

GAOT_HD int cell_coord(float p, float o, float inv_h, int n) {
    float t = (p - o) * inv_h;
    // guard NaN / huge values before the int conversion
    if (!(t > 0.0f)) return 0;
    if (t >= (float)n) return n - 1;
    return (int)t;
}
GAOT_HD int cell_id(const GridParams& g, int cx, int cy, int cz) { return (cx * g.ny + cy) * g.nz + cz; }
GAOT_HD int cell_of(const GridParams& g, float x, float y, float z) {
    return cell_id(g, cell_coord(x, g.ox, g.inv_h, g.nx), cell_coord(y, g.oy, g.inv_h, g.ny),
                   cell_coord(z, g.oz, g.inv_h, g.nz));
}


// mode 0: radius grid (h just above r, reach 1); mode 1: kNN grid (h from source density).
// The cell count is bounded (max_cells, max_dim per axis) so every workspace size is
// data independent; when the bound bites, h grows and `reach` stays valid (h >= r).
GAOT_HD GridParams compute_grid_params(const float* lo, const float* hi, int64_t n_src, float r, int mode,
                                       int max_cells, int max_dim) {
    GridParams gp;
    float ext[3];
    float emax = 0.f;
    for (int a = 0; a < 3; ++a) {
        ext[a] = hi[a] - lo[a];
        if (!(ext[a] >= 0.f)) ext[a] = 0.f;
        emax = ext[a] > emax ? ext[a] : emax;
    }
    if (!(emax > 0.f)) emax = 1.0f;
    float h;
    if (mode == 0) {
        h = r * 1.0001f + 1e-30f;
        if (!(h > 1e-20f)) h = emax;            // r == 0: nothing matches anyway
    } else {
        float vol = 1.f;
        for (int a = 0; a < 3; ++a) { float e = ext[a] > 1e-3f * emax ? ext[a] : 1e-3f * emax; vol *= e; }
        float per = (float)n_src * 0.5f; if (per < 1.0f) per = 1.0f;
        h = cbrtf(vol / per);
    }
    { float hmin = emax / (float)(max_dim - 1); if (h < hmin) h = hmin; }
    int d[3] = {1, 1, 1};
    for (int it = 0; it < 64; ++it) {
        for (int a = 0; a < 3; ++a) {
            int v = (int)floorf(ext[a] / h) + 1;
            d[a] = v < 1 ? 1 : (v > max_dim ? max_dim : v);
        }
        double tot = (double)d[0] * d[1] * d[2];
        if (tot <= (double)max_cells) break;
        h *= 1.01f * (float)cbrt(tot / (double)max_cells);
    }
    gp.ox = lo[0]; gp.oy = lo[1]; gp.oz = lo[2];
    gp.h = h; gp.inv_h = 1.0f / h;
    gp.nx = d[0]; gp.ny = d[1]; gp.nz = d[2];
    gp.ncells = d[0] * d[1] * d[2];
    gp.reach = (mode == 0) ? ((int)floorf(r / h) + 1) : 0;
    return gp;
}

// ---------------------------------------------------------------- radius
// Returns min(#matches, cap).  When EMIT, `list` (capacity >= cap) receives the `cap`
// smallest matching source indices in ascending order.
template <bool EMIT>
GAOT_HD int radius_query(const GridParams& g, const int* __restrict__ cell_start,
                         const SrcPoint* __restrict__ pts, float qx, float qy, float qz,
                         float r2, int cap, int* list) {
    const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx);
    const int cy = cell_coord(qy, g.oy, g.inv_h, g.ny);
    const int cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
    const int R = g.reach;
    // a query far outside the grid cannot have neighbours: its clamped cell is only
    // meaningful if the true coordinate is within `reach` cells of the grid
    int n = 0;
    const int x0 = clampi(cx - R, 0, g.nx - 1), x1 = clampi(cx + R, 0, g.nx - 1);
    const int y0 = clampi(cy - R, 0, g.ny - 1), y1 = clampi(cy + R, 0, g.ny - 1);
    const int z0 = clampi(cz - R, 0, g.nz - 1), z1 = clampi(cz + R, 0, g.nz - 1);
    for (int ix = x0; ix <= x1; ++ix) {
        for (int iy = y0; iy <= y1; ++iy) {
            const int beg = cell_start[cell_id(g, ix, iy, z0)];
            const int end = cell_start[cell_id(g, ix, iy, z1) + 1];
            for (int p = beg; p < end; ++p) {
                const SrcPoint s = pts[p];
                const float d2 = dist2(s.x, s.y, s.z, qx, qy, qz);
                if (d2 < r2) {
                    if (!EMIT) {
                        ++n;
                    } else {
                        // keep the `cap` smallest indices, ascending
                        if (n < cap) {
                            int j = n++;
                            while (j > 0 && list[j - 1] > s.idx) { list[j] = list[j - 1]; --j; }
                            list[j] = s.idx;
                        } else if (s.idx < list[cap - 1]) {
                            int j = cap - 1;
                            while (j > 0 && list[j - 1] > s.idx) { list[j] = list[j - 1]; --j; }
                            list[j] = s.idx;
                        }
                    }
                }
            }
        }
    }
    if (!EMIT && n > cap) n = cap;
    return n;
}

// ---------------------------------------------------------------- kNN
// bd/bi (capacity >= k) receive the k best in ascending (d2, idx) order; returns the
// number found (min(k, #sources with d2 < 1e10)).
template <int KMAX>
GAOT_HD int knn_query(const GridParams& g, const int* __restrict__ cell_start,
                      const SrcPoint* __restrict__ pts, float qx, float qy, float qz,
                      int k, float* bd, int* bi) {
    const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx);
    const int cy = cell_coord(qy, g.oy, g.inv_h, g.ny);
    const int cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
    int n = 0;
    const int maxdim = g.nx > g.ny ? (g.nx > g.nz ? g.nx : g.nz) : (g.ny > g.nz ? g.ny : g.nz);
    for (int rho = 0; rho < maxdim; ++rho) {
        const int x0 = cx - rho, x1 = cx + rho, y0 = cy - rho, y1 = cy + rho, z0 = cz - rho, z1 = cz + rho;
        const int xa = x0 < 0 ? 0 : x0, xb = x1 >= g.nx ? g.nx - 1 : x1;
        const int ya = y0 < 0 ? 0 : y0, yb = y1 >= g.ny ? g.ny - 1 : y1;
        const int za = z0 < 0 ? 0 : z0, zb = z1 >= g.nz ? g.nz - 1 : z1;
        for (int ix = xa; ix <= xb; ++ix) {
            for (int iy = ya; iy <= yb; ++iy) {
                const bool inner = (rho > 0) && ix > x0 && ix < x1 && iy > y0 && iy < y1;
                // inner columns were fully visited up to rho-1: only the two new z end cells remain
                const int nseg = inner ? 2 : 1;
                for (int sgi = 0; sgi < nseg; ++sgi) {
                    int lo, hi;
                    if (!inner) { lo = za; hi = zb; }
                    else if (sgi == 0) { if (z0 < 0) continue; lo = hi = z0; }
                    else { if (z1 >= g.nz) continue; lo = hi = z1; }
                    const int beg = cell_start[cell_id(g, ix, iy, lo)];
                    const int end = cell_start[cell_id(g, ix, iy, hi) + 1];
                    for (int p = beg; p < end; ++p) {
                        const SrcPoint s = pts[p];
                        const float d2 = dist2(s.x, s.y, s.z, qx, qy, qz);
                        if (!(d2 < 1e10f)) continue;
                        bool take = n < k;
                        if (!take) take = (d2 < bd[k - 1]) || (d2 == bd[k - 1] && s.idx < bi[k - 1]);
                        if (take) {
                            int j = n < k ? n++ : k - 1;
                            while (j > 0 && (bd[j - 1] > d2 || (bd[j - 1] == d2 && bi[j - 1] > s.idx))) {
                                bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j;
                            }
                            bd[j] = d2; bi[j] = s.idx;
                        }
                    }
                }
            }
        }
        // everything outside the visited block is at least `lb` away (faces clipped by the
        // grid boundary have no sources beyond them)
        const bool covers = x0 <= 0 && y0 <= 0 && z0 <= 0 && x1 >= g.nx - 1 && y1 >= g.ny - 1 && z1 >= g.nz - 1;
        if (covers) break;
        if (n >= k) {
            float lb = 3.0e38f;
            if (x0 > 0)        { float t = qx - (g.ox + (float)x0 * g.h);       lb = t < lb ? t : lb; }
            if (x1 < g.nx - 1) { float t = (g.ox + (float)(x1 + 1) * g.h) - qx; lb = t < lb ? t : lb; }
            if (y0 > 0)        { float t = qy - (g.oy + (float)y0 * g.h);       lb = t < lb ? t : lb; }
            if (y1 < g.ny - 1) { float t = (g.oy + (float)(y1 + 1) * g.h) - qy; lb = t < lb ? t : lb; }
            if (z0 > 0)        { float t = qz - (g.oz + (float)z0 * g.h);       lb = t < lb ? t : lb; }
            if (z1 < g.nz - 1) { float t = (g.oz + (float)(z1 + 1) * g.h) - qz; lb = t < lb ? t : lb; }
            lb -= 1e-3f * g.h;                       // cell-assignment / rounding slack
            if (lb > 0.0f && bd[k - 1] < lb * lb * 0.999f) break;
        }
    }
    return n;
}

}  // namespace gaot
