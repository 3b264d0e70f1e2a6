// Uniform-grid spatial index shared by the kNN, outlier and branch-selection kernels:
// counting sort of points into cells of size h (x fastest, so a run of cells along x is one
// contiguous range of the sorted array).  Binning is monotone in each coordinate, so the cell
// range [cell(q-r), cell(q+r)] provably covers every point within r of q.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace st {

struct Grid {
    float ox, oy, oz, inv_h, h;
    int nx, ny, nz;
};

struct GridBuild {
    Grid g;
    int32_t *cell_start;  // [cells+1]
    float4 *sorted;       // [m] xyz + index bits
    int64_t cells;
};

__device__ __forceinline__ int ordered_int(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float from_ordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

static __global__ void k_bbox(const float *__restrict__ p, int m, int *__restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int v = ordered_int(p[3 * (size_t)i + a]);
            lo[a] = min(lo[a], v);
            hi[a] = max(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(mm + a, lo[a]); atomicMax(mm + 3 + a, hi[a]); }
    }
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int n) {
    int c = (int)floorf((v - o) * inv_h);
    return max(0, min(n - 1, c));
}

// Visits the grid rows (runs of cells along x) that can hold a point within sqrt(lim2) of q, NEAREST ROWS FIRST
// (cell offsets 0, +1, -1, +2, ... in z and y), and hands each row's range of the cell-sorted array to
// row_fn(beg, end).  row_fn may shrink lim2 (e.g. to the K-th best squared distance found so far): rows and
// cells that lie farther than that are skipped, so in a dense neighbourhood only a few cells are ever read.
// Bounds are conservative (binning is monotone; `slack` covers its rounding): no point with d2 <= lim2 is missed.
template <typename F>
__device__ __forceinline__ void for_rows_near_first(const Grid &g, const int32_t *__restrict__ cell_start, float qx, float qy, float qz,
                                                    float reach, float &lim2, F &&row_fn) {
    const float rr = reach * 1.0001f + 1e-7f;
    const int x0 = cell_coord(qx - rr, g.ox, g.inv_h, g.nx), x1 = cell_coord(qx + rr, g.ox, g.inv_h, g.nx);
    const int y0 = cell_coord(qy - rr, g.oy, g.inv_h, g.ny), y1 = cell_coord(qy + rr, g.oy, g.inv_h, g.ny);
    const int z0 = cell_coord(qz - rr, g.oz, g.inv_h, g.nz), z1 = cell_coord(qz + rr, g.oz, g.inv_h, g.nz);
    const int cyq = cell_coord(qy, g.oy, g.inv_h, g.ny), czq = cell_coord(qz, g.oz, g.inv_h, g.nz);
    const float slack = 1e-4f * g.h + 1e-6f;
    const int nz2 = 2 * max(czq - z0, z1 - czq), ny2 = 2 * max(cyq - y0, y1 - cyq);
    for (int az = 0; az <= nz2; ++az) {
        const int cz = czq + ((az & 1) ? ((az + 1) >> 1) : -((az + 1) >> 1));
        if (cz < z0 || cz > z1) continue;
        float dz = 0.f;
        if (cz > czq) dz = fmaxf(0.f, (g.oz + (float)cz * g.h - slack) - qz);
        else if (cz < czq) dz = fmaxf(0.f, qz - (g.oz + (float)(cz + 1) * g.h + slack));
        if (dz * dz > lim2 * 1.00001f) continue;
        for (int ay = 0; ay <= ny2; ++ay) {
            const int cy = cyq + ((ay & 1) ? ((ay + 1) >> 1) : -((ay + 1) >> 1));
            if (cy < y0 || cy > y1) continue;
            float dy = 0.f;
            if (cy > cyq) dy = fmaxf(0.f, (g.oy + (float)cy * g.h - slack) - qy);
            else if (cy < cyq) dy = fmaxf(0.f, qy - (g.oy + (float)(cy + 1) * g.h + slack));
            const float rem = lim2 * 1.00001f - dz * dz - dy * dy;
            if (rem < 0.f) continue;
            const float half = sqrtf(rem) * 1.0001f + slack;
            const int xa = max(x0, cell_coord(qx - half, g.ox, g.inv_h, g.nx)), xb = min(x1, cell_coord(qx + half, g.ox, g.inv_h, g.nx));
            const int rowc = (cz * g.ny + cy) * g.nx;
            row_fn(__ldg(cell_start + rowc + xa), __ldg(cell_start + rowc + xb + 1));      // cells contiguous in x
        }
    }
}

static __global__ void k_cell_count(const float *__restrict__ p, int m, Grid g, int32_t *__restrict__ cell_of, int32_t *count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int cx = cell_coord(p[3 * (size_t)i], g.ox, g.inv_h, g.nx);
    int cy = cell_coord(p[3 * (size_t)i + 1], g.oy, g.inv_h, g.ny);
    int cz = cell_coord(p[3 * (size_t)i + 2], g.oz, g.inv_h, g.nz);
    int c = (cz * g.ny + cy) * g.nx + cx;
    cell_of[i] = c;
    atomicAdd(count + c, 1);
}

static __global__ void k_cell_scatter(const float *__restrict__ p, int m, const int32_t *__restrict__ cell_of,
                               const int32_t *__restrict__ cell_start, int32_t *cursor, float4 *__restrict__ sorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int c = cell_of[i];
    int pos = cell_start[c] + atomicAdd(cursor + c, 1);
    sorted[pos] = make_float4(p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2], __int_as_float(i));
}

static inline int64_t cell_budget(int64_t m) {
    int64_t b = 8 * m;
    if (b < 4096) b = 4096;
    if (b > (1ll << 24)) b = 1ll << 24;
    return b;
}

static inline size_t grid_ws_bytes(int64_t m) {
    int64_t cells = cell_budget(m) + 1;
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (int *)nullptr, (int *)nullptr, (int)cells);
    return align_up(scan) + 3 * align_up(cells * 4) + align_up(m * 4) + align_up(m * 16) + 4096;
}

// Builds the grid over dst with target cell size h_target (enlarged if the cell budget would be exceeded).
static inline int build_grid(const float *dst, int64_t m, float h_target, Carver &cv, GridBuild &gb, cudaStream_t s) {
    int *mm = cv.take<int>(8);
    int h_mm[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    ST_CHECK_CUDA(cudaMemcpyAsync(mm, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, s));
    k_bbox<<<148, 256, 0, s>>>(dst, (int)m, mm);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        int l = h_mm[a], h = h_mm[3 + a];
        l = l >= 0 ? l : l ^ 0x7FFFFFFF;
        h = h >= 0 ? h : h ^ 0x7FFFFFFF;
        memcpy(&lo[a], &l, 4);
        memcpy(&hi[a], &h, 4);
    }
    int64_t budget = cell_budget(m);
    double h = h_target > 1e-6f ? h_target : 1e-6;
    int64_t nx, ny, nz;
    while (true) {
        nx = (int64_t)((hi[0] - lo[0]) / h) + 1;
        ny = (int64_t)((hi[1] - lo[1]) / h) + 1;
        nz = (int64_t)((hi[2] - lo[2]) / h) + 1;
        if (nx * ny * nz <= budget) break;
        h *= 1.26;
    }
    gb.g = Grid{lo[0], lo[1], lo[2], (float)(1.0 / h), (float)h, (int)nx, (int)ny, (int)nz};
    gb.cells = nx * ny * nz;
    int32_t *count = cv.take<int32_t>(budget + 1);
    gb.cell_start = cv.take<int32_t>(budget + 1);
    int32_t *cursor = cv.take<int32_t>(budget + 1);
    int32_t *cell_of = cv.take<int32_t>(m);
    gb.sorted = cv.take<float4>(m);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, count, gb.cell_start, (int)(gb.cells + 1));
    void *scan_ws = cv.take<char>(scan_bytes);
    if (!cv.ok()) { set_error("kNN grid: workspace too small"); return ST_ERR_WORKSPACE; }
    ST_CHECK_CUDA(cudaMemsetAsync(count, 0, (gb.cells + 1) * 4, s));
    ST_CHECK_CUDA(cudaMemsetAsync(cursor, 0, (gb.cells + 1) * 4, s));
    unsigned g = (unsigned)cdiv(m, 256);
    k_cell_count<<<g, 256, 0, s>>>(dst, (int)m, gb.g, cell_of, count);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, count, gb.cell_start, (int)(gb.cells + 1), s));
    k_cell_scatter<<<g, 256, 0, s>>>(dst, (int)m, cell_of, gb.cell_start, cursor, gb.sorted);
    ST_CHECK_LAUNCH();
    return ST_OK;
}


}  // namespace st
