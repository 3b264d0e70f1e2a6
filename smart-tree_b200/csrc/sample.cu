// Greedy branch extraction (reference sample_tree, smart_tree/skeleton/path.py:49-140), one
// resident CTA per connected component, no host round trips:
//   * "farthest unallocated vertex" = next live entry of a list sorted once by (distance desc,
//     index asc): distances only ever drop to -1, so a forward cursor replaces the per-branch argmax
//   * the route to the first allocated ancestor is traced by one thread (pointer chase)
//   * points near the path are claimed through the uniform grid: one warp per path vertex scans the
//     cell rows within r = max path radius and races a 64-bit atomicMin of (d2 bits, path position)
//     per point -> nearest path vertex, ties to the lowest position, exactly the FRNN K=1 result
#include "grid.cuh"

using namespace st;

constexpr unsigned long long BEST_NONE = 0xFFFFFFFFFFFFFFFFull;

// debug counters of the last launch's block 0 (cycles per phase, iterations, path vertices)
__device__ unsigned long long g_st_stats[8];

__global__ void k_st_init(const int32_t *__restrict__ pred, const float *__restrict__ tree_dist, int n, float *distw,
                          uint8_t *alloc, int32_t *branch_id, unsigned long long *best, const int32_t *__restrict__ comp_off,
                          int n_comp, unsigned long long *keys, int32_t *vals) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    float d = pred[v] > 0 ? tree_dist[v] : -1.f;   // path.py:71-72  (`preds > 0`, sic)
    distw[v] = d;
    alloc[v] = 0;
    branch_id[v] = -1;
    best[v] = BEST_NONE;
    int lo = 0, hi = n_comp;                        // component of v: last c with comp_off[c] <= v
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (comp_off[mid] <= v) lo = mid; else hi = mid; }
    unsigned bits = d > 0.f ? __float_as_uint(d) : 0u;
    keys[v] = ((unsigned long long)lo << 32) | (unsigned long long)(0xFFFFFFFFu - bits);
    vals[v] = v;
}

struct SampleArgs {
    const float *pts;
    const float *radii;
    const int32_t *pred;
    const int32_t *comp_off;
    Grid g;
    const int32_t *cell_start;
    const float4 *sorted;
    const int32_t *order;
    float *distw;
    uint8_t *alloc;
    int32_t *branch_id;
    unsigned long long *best;
    int32_t *path_out, *branch_len, *branch_parent, *comp_nb, *comp_np;
};

template <bool CLAIM>
__device__ __forceinline__ void scan_path_vertex(const SampleArgs &a, int base, int nc, int v, unsigned pos, float r, float r2,
                                                 int lane, bool emit, int bid) {
    const float px = a.pts[3 * (size_t)(base + v)], py = a.pts[3 * (size_t)(base + v) + 1], pz = a.pts[3 * (size_t)(base + v) + 2];
    const float vr = a.radii[base + v];
    const Grid &g = a.g;
    float rr = r * 1.0001f + 1e-7f;
    int x0 = cell_coord(px - rr, g.ox, g.inv_h, g.nx), x1 = cell_coord(px + rr, g.ox, g.inv_h, g.nx);
    int y0 = cell_coord(py - rr, g.oy, g.inv_h, g.ny), y1 = cell_coord(py + rr, g.oy, g.inv_h, g.ny);
    int z0 = cell_coord(pz - rr, g.oz, g.inv_h, g.nz), z1 = cell_coord(pz + rr, g.oz, g.inv_h, g.nz);
    for (int cz = z0; cz <= z1; ++cz)
        for (int cy = y0; cy <= y1; ++cy) {
            int rowc = (cz * g.ny + cy) * g.nx;
            int beg = __ldg(a.cell_start + rowc + x0), end = __ldg(a.cell_start + rowc + x1 + 1);
            for (int t = beg + lane; t < end; t += 32) {
                float4 q = __ldg(a.sorted + t);
                int gi = __float_as_int(q.w);
                if (gi < base || gi >= base + nc) continue;
                float d2 = dist2_exact(q.x, q.y, q.z, px, py, pz);
                if (!(d2 < r2)) continue;
                if (CLAIM) {
                    unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | pos;
                    atomicMin(a.best + gi, key);
                } else {
                    unsigned long long b = __ldcg(a.best + gi);
                    if ((unsigned)(b & 0xFFFFFFFFull) != pos || (unsigned)(b >> 32) != __float_as_uint(d2)) continue;
                    if (sqrtf(d2) < vr) {           // path.py:37-39: inside the radius of its nearest path vertex
                        a.distw[gi] = -1.f;
                        a.alloc[gi] = 1;
                        if (emit) a.branch_id[gi] = bid;
                    }
                    __stcg(a.best + gi, BEST_NONE);
                }
            }
        }
}

__global__ void __launch_bounds__(1024) k_sample_tree(SampleArgs a) {
    const int c = blockIdx.x;
    const int base = a.comp_off[c];
    const int nc = a.comp_off[c + 1] - base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_minpos, s_len, s_term, s_parent, s_rbits;
    int cursor = 0, bid = 0, pcur = 0;
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc = clock64();
#define ST_PHASE(i) do { if (tid == 0) { long long _t = clock64(); st[i] += (unsigned long long)(_t - tc); tc = _t; } } while (0)
    while (true) {
        // ---- 1. farthest live vertex
        int f = -1;
        while (cursor < nc) {
            if (tid == 0) s_minpos = INT_MAX;
            __syncthreads();
            int pos = cursor + tid;
            if (pos < nc) {
                int v = a.order[base + pos];
                if (a.distw[v] > 0.f) atomicMin(&s_minpos, pos);
            }
            __syncthreads();
            int m = s_minpos;
            __syncthreads();
            if (m != INT_MAX) { f = a.order[base + m] - base; cursor = m + 1; break; }
            cursor += blockDim.x;
        }
        if (f < 0) break;
        ST_PHASE(0);
        // ---- 2. trace the route to the first allocated ancestor (path.py:9-16)
        if (tid == 0) {
            int len = 0, i = f;
            int *out = a.path_out + base + pcur;
            while (i >= 0 && !a.alloc[base + i] && pcur + len < nc) {
                out[len++] = i;
                i = a.pred[base + i];
            }
            s_len = len;
            s_term = i;
            s_parent = a.branch_id[base + (i >= 0 ? i : nc - 1)];   // -1 wraps to the last vertex (path.py:132)
            s_rbits = 0;
        }
        __syncthreads();
        ST_PHASE(1);
        const int len = s_len;
        if (tid == 0) { st[5] += 1; st[6] += len; }
        const int *path = a.path_out + base + pcur;   // farthest first; position pos = len-1-jj is root side first
        // ---- 3. search radius = max radius over the path
        float rl = 0.f;
        for (int jj = tid; jj < len; jj += blockDim.x) rl = fmaxf(rl, a.radii[base + path[jj]]);
        for (int o = 16; o; o >>= 1) rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));
        if (lane == 0 && rl > 0.f) atomicMax(&s_rbits, __float_as_int(rl));
        __syncthreads();
        const float r = __int_as_float(s_rbits);
        const float r2 = __fmul_rn(r, r);
        const bool emit = len >= 2;
        // ---- 4./5. claim the nearest path vertex per point, then resolve the winners
        if (r > 0.f) {
            for (int jj = warp; jj < len; jj += (blockDim.x >> 5))
                scan_path_vertex<true>(a, base, nc, path[jj], (unsigned)(len - 1 - jj), r, r2, lane, emit, bid);
            __syncthreads();
            ST_PHASE(2);
            for (int jj = warp; jj < len; jj += (blockDim.x >> 5))
                scan_path_vertex<false>(a, base, nc, path[jj], (unsigned)(len - 1 - jj), r, r2, lane, emit, bid);
        }
        __syncthreads();
        ST_PHASE(3);
        // ---- 6. the path itself
        for (int jj = tid; jj < len; jj += blockDim.x) {
            int v = base + path[jj];
            a.distw[v] = -1.f;
            a.alloc[v] = 1;
            if (emit) a.branch_id[v] = bid;
        }
        __syncthreads();
        // ---- 7. emit the branch (root side first)
        if (emit) {
            int *pp = a.path_out + base + pcur;
            for (int jj = tid; jj < len / 2; jj += blockDim.x) {
                int t = pp[jj];
                pp[jj] = pp[len - 1 - jj];
                pp[len - 1 - jj] = t;
            }
            if (tid == 0) {
                a.branch_len[base + bid] = len;
                a.branch_parent[base + bid] = s_parent;
            }
            ++bid;
            pcur += len;
        }
        __syncthreads();
        ST_PHASE(4);
    }
    if (tid == 0) { a.comp_nb[c] = bid; a.comp_np[c] = pcur; }
    if (tid == 0 && c == 0) for (int i = 0; i < 8; ++i) g_st_stats[i] = st[i];
}

static size_t sort_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)n);
    return b;
}

extern "C" size_t st_sample_tree_workspace_bytes(int64_t n, int32_t n_comp) {
    return grid_ws_bytes(n) + align_up(sort_bytes(n)) + 2 * align_up(n * 8) + 2 * align_up(n * 4) + align_up(n * 4) + align_up(n) +
           align_up(n * 4) + align_up(n * 8) + 4096;
}

extern "C" int st_sample_tree(const float *medial_pts, const float *radii, const int32_t *pred, const float *tree_dist,
                              const int32_t *comp_off, int32_t n_comp, int64_t n, float cell_size, int32_t *path_vertices,
                              int32_t *branch_len, int32_t *branch_parent, int32_t *comp_n_branches, int32_t *comp_n_path,
                              void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || n_comp == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31), "n");
    Carver cv(workspace, workspace_bytes);
    float *distw = cv.take<float>(n);
    uint8_t *alloc = cv.take<uint8_t>(n);
    int32_t *branch_id = cv.take<int32_t>(n);
    unsigned long long *best = cv.take<unsigned long long>(n);
    unsigned long long *keys = cv.take<unsigned long long>(n);
    unsigned long long *keys2 = cv.take<unsigned long long>(n);
    int32_t *vals = cv.take<int32_t>(n);
    int32_t *order = cv.take<int32_t>(n);
    size_t sb = sort_bytes(n);
    void *sort_ws = cv.take<char>(sb);
    if (!cv.ok()) { set_error("st_sample_tree: workspace too small"); return ST_ERR_WORKSPACE; }
    k_st_init<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(pred, tree_dist, (int)n, distw, alloc, branch_id, best, comp_off, n_comp, keys, vals);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_ws, sb, keys, keys2, vals, order, (int)n, 0, 64, s));
    GridBuild gb;
    int rc = build_grid(medial_pts, n, cell_size, cv, gb, s);
    if (rc) return rc;
    SampleArgs a{medial_pts, radii, pred, comp_off, gb.g, gb.cell_start, gb.sorted, order, distw, alloc, branch_id, best,
                 path_vertices, branch_len, branch_parent, comp_n_branches, comp_n_path};
    k_sample_tree<<<n_comp, 1024, 0, s>>>(a);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// debug only (not part of include/st_b200.h): cycles spent by component 0 in
// [find, trace, claim, resolve, finish], iterations, traced path vertices
extern "C" int st_debug_sample_stats(unsigned long long *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_st_stats, sizeof(unsigned long long) * 8));
    return ST_OK;
}
