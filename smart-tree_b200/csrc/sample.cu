// Greedy branch extraction (reference sample_tree, smart_tree/skeleton/path.py:49-140), one
// resident CTA per connected component, no host round trips:
//   * "farthest unallocated vertex" = next live entry of a list sorted once by (distance desc,
//     index asc): distances only ever drop to -1, so a forward cursor replaces the per-branch argmax
//   * the route to the first allocated ancestor is traced by one thread (pointer chase)
//   * points near the path are claimed through the uniform grid: one warp per path vertex scans the
//     cell rows within r = max path radius and races a 64-bit atomicMin of (d2 bits, path position)
//     per point -> nearest path vertex, ties to the lowest position, exactly the FRNN K=1 result
#include <string.h>

#include "grid.cuh"

using namespace st;

constexpr unsigned long long BEST_NONE = 0xFFFFFFFFFFFFFFFFull;

// debug counters of the last launch's block 0 (cycles per phase, iterations, path vertices)
__device__ unsigned long long g_st_stats[8];
__device__ int g_st_iter[1024][8];    // per iteration of block 0: route length, R, cycles of the claim phase, cycles of the iteration

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

__global__ void k_vertex_base(const int32_t *__restrict__ comp_off, int n_comp, int n, int32_t *__restrict__ vbase) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int lo = 0, hi = n_comp;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (comp_off[mid] <= v) lo = mid; else hi = mid; }
    vbase[v] = comp_off[lo];
}

struct SampleArgs {
    const float *pts;
    const float *radii;
    const int32_t *pred;
    const int32_t *comp_off;
    Grid g;
    const int32_t *cell_start;
    const float4 *sorted;
    int32_t *order;         // (distance desc, index asc) list; k_sample_tree_c compacts it in place
    float *distw;
    uint8_t *alloc;
    int32_t *branch_id;
    unsigned long long *best;
    int32_t *path_out, *branch_len, *branch_parent, *comp_nb, *comp_np;
    int32_t *touched;      // [n] segment-local list of points claimed in the current iteration
    int32_t *touch_cnt;    // [2 * n_comp] append counters, alternating by iteration parity
    int32_t *pos;          // k_sample_tree_c only (else null): position of every vertex in `order`, for the tombstones below
};

// A vertex leaves the candidate list.  k_sample_tree_c scans the list for live entries every round; without help it has to
// look up distw of every entry it passes (random L2 loads: 8192 per step and CTA, the whole cost of its phase A), and most of
// them are dead.  So a kill also overwrites the vertex' list entry with -1 -- but only entries at or beyond `stable_from`:
// the entries of the current window are being compacted in place by another CTA at the same time (their liveness is
// re-checked through distw by the next scan, as before).
__device__ __forceinline__ void kill_vertex(const SampleArgs &a, int v, int stable_from) {
    a.distw[v] = -1.f;
    a.alloc[v] = 1;
    if (a.pos) {
        const int p = __ldcg(a.pos + v);
        if (p >= stable_from) a.order[p] = -1;
    }
}

// One (path vertex, grid row) task per warp: the rows within r of the vertex are numbered
// 0 .. RW*RW-1 around the vertex's own cell, so a short path still spreads over every warp of the cluster.
constexpr int PATH_SMEM = 1024;    // path vertices whose position is staged in shared memory per iteration

// One (path vertex, grid row) task per warp: the rows within r of the vertex are numbered 0 .. RW*RW-1
// around the vertex's own cell, so a short path still spreads over every warp of the cluster.  Every
// point within r races a 64-bit atomicMin of (d2 bits, path position); the first claimer of a point
// also appends it to the iteration's touched list, which is all the resolve phase has to walk.
__device__ __forceinline__ void claim_task(const SampleArgs &a, int base, int nc, float px, float py, float pz, unsigned pos, int row,
                                           int R, float r, float r2, int lane, int32_t *touch_cnt) {
    const Grid &g = a.g;
    const float rr = r * 1.0001f + 1e-7f;
    const int RW = 2 * R + 1;
    const int cz = cell_coord(pz, g.oz, g.inv_h, g.nz) + row / RW - R;
    const int cy = cell_coord(py, g.oy, g.inv_h, g.ny) + row % RW - R;
    if (cz < cell_coord(pz - rr, g.oz, g.inv_h, g.nz) || cz > cell_coord(pz + rr, g.oz, g.inv_h, g.nz)) return;
    if (cy < cell_coord(py - rr, g.oy, g.inv_h, g.ny) || cy > cell_coord(py + rr, g.oy, g.inv_h, g.ny)) return;
    const int x0 = cell_coord(px - rr, g.ox, g.inv_h, g.nx), x1 = cell_coord(px + rr, g.ox, g.inv_h, g.nx);
    const int rowc = (cz * g.ny + cy) * g.nx;
    const int beg = __ldg(a.cell_start + rowc + x0), end = __ldg(a.cell_start + rowc + x1 + 1);
    for (int t = beg + lane; t < end; t += 32) {
        float4 q = __ldg(a.sorted + t);
        int gi = __float_as_int(q.w);
        if (gi < base || gi >= base + nc) continue;
        float d2 = dist2_exact(q.x, q.y, q.z, px, py, pz);
        if (!(d2 < r2)) continue;
        unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | pos;
        unsigned long long old = atomicMin(a.best + gi, key);
        if (old == BEST_NONE) a.touched[base + atomicAdd(touch_cnt, 1)] = gi;
    }
}

// ---- claim with a neighbour pre-filter -----------------------------------------------------------------
// A point usually lies within r of dozens of consecutive path vertices but only its NEAREST one (smallest
// (d2 bits, path position) key) decides it.  A (point, vertex jj) pair whose key is beaten by one of the
// path neighbours jj-2 .. jj+2 can never be that minimum, so it is dropped after four extra distance
// evaluations; ~1 pair in 20 survives.  Survivors of a SHORT route (whole path in shared memory) are
// compacted into a per-warp queue and settled on the spot by a full scan of the path -- no atomics, no
// second pass; survivors of a long route race the 64-bit atomicMin and are settled by the resolve pass.
constexpr int SCAN_PATH = 512;
constexpr int QUEUE_LEN = 64;

__device__ __forceinline__ unsigned long long claim_key(float4 q, float4 o, unsigned pos) {
    return ((unsigned long long)__float_as_uint(dist2_exact(q.x, q.y, q.z, o.x, o.y, o.z)) << 32) | pos;
}

// lanes < count each settle one queued (sorted position, path index) pair
// `lohi` (optional): per path vertex j the index interval [lo, hi] (lo | hi << 16) that contains every path vertex within
// 2 r of it.  A vertex that beats j for a point within r of j is itself within r of the point, hence within 2 r of j: the
// scan over [lo, hi] decides exactly what the scan over the whole path decides, ~20x cheaper on a 300-vertex route.
// Window of path positions that can hold a vertex nearer to a point than vertex j does: a candidate point of j lies within
// j's own radius r_j (claim_flat), a nearer vertex therefore within 2 r_j of j.
__device__ __forceinline__ void path_windows(const float4 *s_path, int len, float r, int *lohi, int t0, int nt) {
    for (int j = t0; j < len; j += nt) {
        const float4 p = s_path[j];
        const float lim = 2.f * p.w * 1.001f + 1e-6f, lim2 = lim * lim;
        int lo = j, hi = j;
        for (int j2 = 0; j2 < len; ++j2) {
            const float4 q = s_path[j2];
            if (dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) <= lim2) { lo = min(lo, j2); hi = max(hi, j2); }
        }
        lohi[j] = lo | (hi << 16);
    }
}

__device__ __forceinline__ void claim_drain(const SampleArgs &a, const float4 *s_path, int len, const int2 *queue, int first, int count,
                                            int lane, int bid, const int *lohi = nullptr) {
    if (lane >= count) return;
    const int2 e = queue[first + lane];
    const float4 q = __ldg(a.sorted + e.x);
    const float4 p = s_path[e.y];
    const float d2 = dist2_exact(q.x, q.y, q.z, p.x, p.y, p.z);
    const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(len - 1 - e.y);
    bool own = true;
    int jlo = 0, jhi = len - 1;
    if (lohi) { const int w = lohi[e.y]; jlo = w & 0xFFFF; jhi = w >> 16; }
    for (int j2 = jlo; j2 <= jhi; ++j2) own = own && !(claim_key(q, s_path[j2], (unsigned)(len - 1 - j2)) < key);
    if (own && sqrtf(d2) < p.w) {
        const int gi = __float_as_int(q.w);
        kill_vertex(a, gi, 0);          // (no compaction is going on in the iterations that claim through this function)
        if (bid >= 0) a.branch_id[gi] = bid;
    }
}

// Collecting variant for the speculative batches (k_sample_tree_b): nothing is committed; a settled point is appended to
// the member's touched list and stamped with (epoch, member) so that later members of the batch can tell they were
// touched by an earlier one.  Called by whole warps (the append is warp-aggregated).
struct Collect {
    int32_t *tlist;       // this member's touched list
    int *s_tcnt;          // its length (shared memory)
    int32_t *stamp;       // per vertex: max over touchers of epoch * 16 + (15 - member)
    int sval;
};

__device__ __forceinline__ void claim_drain_collect(const SampleArgs &a, const float4 *s_path, int len, const int2 *queue, int first,
                                                    int count, int lane, const Collect &co) {
    bool hit = false;
    int gi = -1;
    if (lane < count) {
        const int2 e = queue[first + lane];
        const float4 q = __ldg(a.sorted + e.x);
        const float4 p = s_path[e.y];
        const float d2 = dist2_exact(q.x, q.y, q.z, p.x, p.y, p.z);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(len - 1 - e.y);
        bool own = true;
        for (int j2 = 0; j2 < len; ++j2) own = own && !(claim_key(q, s_path[j2], (unsigned)(len - 1 - j2)) < key);
        hit = own && sqrtf(d2) < p.w;
        gi = __float_as_int(q.w);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (!m) return;
    int pos = 0;
    if (lane == 0) pos = atomicAdd(co.s_tcnt, __popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (hit) {
        co.tlist[pos + __popc(m & ((1u << lane) - 1u))] = gi;
        atomicMax(co.stamp + gi, co.sval);
    }
}

// Range [beg, end) of the cell-sorted point array covered by grid row `row` (numbered 0 .. RW*RW-1 around the
// cell of path vertex p) within the search sphere of radius rr: the chord of the sphere along that row.
__device__ __forceinline__ void row_range(const SampleArgs &a, float4 p, int row, int R, float rr, int &beg, int &end) {
    const Grid &g = a.g;
    beg = end = 0;
    const int RW = 2 * R + 1;
    const int pcz = cell_coord(p.z, g.oz, g.inv_h, g.nz), pcy = cell_coord(p.y, g.oy, g.inv_h, g.ny);
    const int cz = pcz + row / RW - R;
    const int cy = pcy + row % RW - R;
    if (cz < cell_coord(p.z - rr, g.oz, g.inv_h, g.nz) || cz > cell_coord(p.z + rr, g.oz, g.inv_h, g.nz)) return;
    if (cy < cell_coord(p.y - rr, g.oy, g.inv_h, g.ny) || cy > cell_coord(p.y + rr, g.oy, g.inv_h, g.ny)) return;
    // binning is monotone, so every point of row (cy, cz) is at least dy / dz away from p in y / z (near-side cell
    // boundary, with slack for the rounding of the binning)
    const float slack = 1e-4f * g.h + 1e-6f;
    float dy = 0.f, dz = 0.f;
    if (cy > pcy) dy = fmaxf(0.f, (g.oy + (float)cy * g.h - slack) - p.y);
    else if (cy < pcy) dy = fmaxf(0.f, p.y - (g.oy + (float)(cy + 1) * g.h + slack));
    if (cz > pcz) dz = fmaxf(0.f, (g.oz + (float)cz * g.h - slack) - p.z);
    else if (cz < pcz) dz = fmaxf(0.f, p.z - (g.oz + (float)(cz + 1) * g.h + slack));
    const float rem = rr * rr - dy * dy - dz * dz;
    if (rem < 0.f) return;
    const float half = sqrtf(rem) * 1.0001f + slack;
    const int x0 = cell_coord(p.x - half, g.ox, g.inv_h, g.nx), x1 = cell_coord(p.x + half, g.ox, g.inv_h, g.nx);
    const int rowc = (cz * g.ny + cy) * g.nx;
    beg = __ldg(a.cell_start + rowc + x0);
    end = __ldg(a.cell_start + rowc + x1 + 1);
}

__device__ __forceinline__ int block_excl_scan_1024(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[31];
    return x - v + (warp ? s_warp[warp - 1] : 0);
}

// Claim pass of one CTA over its share of the (path vertex, grid row) tasks, FLATTENED: medial points pile up
// along the branch axes, so a few rows hold most of the candidates and a warp per row leaves the cluster
// waiting for one warp.  Per round every thread looks up the range of ONE task (a single round trip for
// 1024 tasks), a block scan lays the ranges end to end, and the 1024 threads walk the concatenated points.
template <int MODE>
__device__ __forceinline__ void claim_flat(const SampleArgs &a, int base, int nc, const float4 *s_path, int len, int nflat, int R2, int R,
                                           float r, float r2, int bid, int32_t *touch_cnt, int2 *queue, int *s_off, int *s_beg,
                                           int *s_jj, int *s_warp, unsigned cr, unsigned CL, const Collect *co = nullptr,
                                           const int *lohi = nullptr) {
    constexpr bool SHORT = MODE != 0;
    const int tid = threadIdx.x, lane = tid & 31;
    const float rr = r * 1.0001f + 1e-7f;
    const int lim = len < PATH_SMEM ? len : PATH_SMEM;
    const float4 none = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    int qn = 0;
    for (int tb = 0; (long long)tb * CL < nflat; tb += 1024) {
        const long long t = (long long)(tb + tid) * CL + cr;      // tasks interleaved over the CTAs of the cluster
        int beg = 0, end = 0, jj = 0;
        if (t < nflat) {
            jj = (int)(t / R2);
            // SHORT modes decide ownership by an explicit nearest-vertex test (claim_drain), so a vertex only has to offer the
            // points it could claim itself, i.e. those within ITS radius (a claimed point lies within the radius of its nearest
            // vertex); the atomicMin variant needs every vertex within the route's largest radius to register
            row_range(a, s_path[jj], (int)(t % R2), R, SHORT ? s_path[jj].w * 1.0001f + 1e-7f : rr, beg, end);
        }
        int total;
        const int ex = block_excl_scan_1024(end > beg ? end - beg : 0, s_warp, total);
        s_off[tid] = ex; s_beg[tid] = beg; s_jj[tid] = jj;
        __syncthreads();
        auto locate = [&](int e, int &tp, int &j) -> float4 {
            if (e >= total) { tp = 0; j = 0; return none; }
            int lo = 0, hi = 1023;
            while (lo < hi) {                       // last task whose offset is <= e (empty tasks share the offset of their successor)
                const int mid = (lo + hi + 1) >> 1;
                if (s_off[mid] <= e) lo = mid; else hi = mid - 1;
            }
            tp = s_beg[lo] + (e - s_off[lo]);
            j = s_jj[lo];
            return __ldg(a.sorted + tp);
        };
        int tpn, jn;
        float4 qnext = locate(tid, tpn, jn);
        for (int e0 = 0; e0 < total; e0 += 1024) {
            const float4 q = qnext;
            const int tp = tpn, j = jn;
            qnext = locate(e0 + 1024 + tid, tpn, jn);       // next batch in flight while this one is filtered
            const int gi = __float_as_int(q.w);
            bool cand = false;
            unsigned long long key = 0;
            if (gi >= base && gi < base + nc) {
                const float4 p = s_path[j];
                const float d2 = dist2_exact(q.x, q.y, q.z, p.x, p.y, p.z);
                const float rj = p.w * 1.0001f + 1e-7f;
                if (SHORT ? d2 <= rj * rj : d2 < r2) {
                    key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(len - 1 - j);
                    cand = true;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j2 = j + (k < 2 ? k - 2 : k - 1);
                        if (j2 >= 0 && j2 < lim && claim_key(q, s_path[j2], (unsigned)(len - 1 - j2)) < key) cand = false;
                    }
                }
            }
            if (SHORT) {
                const unsigned m = __ballot_sync(0xffffffffu, cand);
                if (cand) queue[qn + __popc(m & ((1u << lane) - 1u))] = make_int2(tp, j);
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    if (MODE == 2) claim_drain_collect(a, s_path, len, queue, qn, 32, lane, *co);
                    else claim_drain(a, s_path, len, queue, qn, 32, lane, bid, lohi);
                    __syncwarp();
                }
            } else if (cand) {
                const unsigned long long old = atomicMin(a.best + gi, key);
                if (old == BEST_NONE) a.touched[base + atomicAdd(touch_cnt, 1)] = gi;
            }
        }
        __syncthreads();
    }
    if (SHORT) {
        __syncwarp();
        if (MODE == 2) claim_drain_collect(a, s_path, len, queue, 0, qn, lane, *co);
        else claim_drain(a, s_path, len, queue, 0, qn, lane, bid, lohi);
    }
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

// Ancestor table in base 8: jump[(7*L + d-1)][v] = (d * 8^L)-th ancestor of v in the predecessor tree
// (component-local ids, -1 past the root), L = 0..3, d = 1..7.  Hop count h < 4096 = up to four dependent
// loads (one per non-zero octal digit); the 64 nearest ancestors cost at most two.
constexpr int JUMP_L = 4;
constexpr int JUMP_LEVELS = 7 * JUMP_L;

// digit 1 of level L: 8^L = 7 * 8^(L-1) + 8^(L-1)
__global__ void k_jump_first(const int32_t *__restrict__ pred, int32_t *__restrict__ jump, const int32_t *__restrict__ vbase, int n, int L) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int32_t *dst = jump + (size_t)(7 * L) * n;
    if (L == 0) { dst[v] = pred[v]; return; }
    const int32_t *p7 = jump + (size_t)(7 * (L - 1) + 6) * n, *p1 = jump + (size_t)(7 * (L - 1)) * n;
    const int x = p7[v];
    dst[v] = x < 0 ? -1 : p1[vbase[v] + x];
}
// digits 2..7 of level L by repeated application of digit 1
__global__ void k_jump_rest(int32_t *__restrict__ jump, const int32_t *__restrict__ vbase, int n, int L) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int32_t *p1 = jump + (size_t)(7 * L) * n;
    const int b = vbase[v];
    int x = p1[v];
    for (int d = 2; d <= 7; ++d) {
        x = x < 0 ? -1 : p1[b + x];
        jump[(size_t)(7 * L + d - 1) * n + v] = x;
    }
}

// One thread-block CLUSTER per connected component.  Every CTA of the cluster redundantly (and
// deterministically) finds the next farthest vertex and traces its route by binary lifting -- thread j
// resolves the j-th ancestor in <= 10 dependent loads, so 1024 hops cost one round instead of 1024
// pointer-chase latencies -- then the CTAs split the path vertices between them for the claim /
// resolve scans.  Two cluster barriers per branch.
__global__ void __launch_bounds__(1024, 1) k_sample_tree(SampleArgs a, const int32_t *__restrict__ jump, int n_total) {
    const unsigned CL = cluster_size(), cr = cluster_rank();
    const int c = blockIdx.x / CL;
    const int base = a.comp_off[c];
    const int nc = a.comp_off[c + 1] - base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = cr * 32 + warp, nwarp = CL * 32;
    const int gtid = cr * 1024 + tid, nthr = CL * 1024;
    __shared__ int s_minpos, s_first, s_term, s_rbits;
    __shared__ float4 s_path[PATH_SMEM];      // xyz + radius of the path vertices (first PATH_SMEM of them)
    __shared__ int2 s_queue[32][QUEUE_LEN];   // per-warp survivors of the neighbour pre-filter (short routes)
    __shared__ int s_off[1024], s_beg[1024], s_jj[1024], s_scan[32];     // flattened claim: per-task offset / range start / path index
    int cursor = 0, bid = 0, pcur = 0, iter = 0;
    __shared__ unsigned long long st[8];
    __shared__ long long s_tc, s_t0;
    if (tid == 0) { for (int i = 0; i < 8; ++i) st[i] = 0; s_tc = clock64(); s_t0 = s_tc; }
#define ST_PHASE(i) do { if (tid == 0) { long long _t = clock64(); st[i] += (unsigned long long)(_t - s_tc); s_tc = _t; } } while (0)
    while (true) {
        // ---- 1. farthest live vertex: next entry of the (distance desc, index asc) list that is still unallocated
        int f = -1;
        while (cursor < nc) {
            if (tid == 0) s_minpos = INT_MAX;
            __syncthreads();
            int pos = cursor + tid;
            if (pos < nc) {
                int v = __ldg(a.order + base + pos);
                if (__ldcg(a.distw + v) > 0.f) atomicMin(&s_minpos, pos);
            }
            __syncthreads();
            int m = s_minpos;
            __syncthreads();
            if (m != INT_MAX) { f = __ldg(a.order + base + m) - base; cursor = m + 1; break; }
            cursor += blockDim.x;
        }
        if (f < 0) break;
        ST_PHASE(0);
        // ---- 2. route to the first allocated ancestor (path.py:9-16), 1024 hops per round
        int len = 0, cur = f, term = -1;
        float rl = 0.f;                           // largest radius among the path vertices this thread staged
        int *out = a.path_out + base + pcur;      // farthest first here; reversed when the branch is emitted
        int span = 64;                            // most routes are a few dozen hops: try 64 ancestors (<= 2 dependent loads) first
        while (true) {
            if (tid == 0) s_first = 1024;
            __syncthreads();
            int x = cur;
            if (tid < span) {
#pragma unroll
                for (int L = 0; L < JUMP_L; ++L) {
                    const int d = (tid >> (3 * L)) & 7;
                    if (d && x >= 0) x = __ldg(jump + (size_t)(7 * L + d - 1) * n_total + base + x);
                }
                bool stop = x < 0 || __ldcg(a.alloc + base + x) != 0;
                if (stop) atomicMin(&s_first, tid);
            }
            __syncthreads();
            const int first = s_first;
            if (first == 1024 && span < 1024) { span = 1024; __syncthreads(); continue; }   // not within 64 hops: full width
            if (tid == first) s_term = x;
            const int cnt = min(first, max(nc - pcur - len, 0));
            if (tid < cnt) {
                out[len + tid] = x;
                // stage position + radius of this path vertex straight from the register that holds its id
                const int v = base + x;
                const float rv = a.radii[v];
                if (len + tid < PATH_SMEM) s_path[len + tid] = make_float4(a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], rv);
                rl = fmaxf(rl, rv);
            }
            __syncthreads();
            len += cnt;
            if (first < 1024) { term = s_term; break; }
            if (cnt < 1024) { term = -1; break; }           // defensive: path longer than the component
            cur = __ldg(jump + (size_t)(7 * 3 + 1) * n_total + base + cur);      // 2 * 8^3 = 1024 hops further
            __syncthreads();
        }
        const int parent = __ldcg(a.branch_id + base + (term >= 0 ? term : nc - 1));   // -1 wraps to the last vertex (path.py:132)
        if (tid == 0) s_rbits = 0;
        __syncthreads();
        ST_PHASE(1);
        if (tid == 0) { st[5] += 1; st[6] += len; }
        const int *path = out;
        // ---- 3. search radius = max radius over the path (staged in shared memory by the trace above)
        for (int o = 16; o; o >>= 1) rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));
        if (lane == 0 && rl > 0.f) atomicMax(&s_rbits, __float_as_int(rl));
        __syncthreads();
        const float r = __int_as_float(s_rbits);
        const float r2 = __fmul_rn(r, r);
        long long t_a = clock64(), t_b = t_a, t_c = t_a;
        const bool emit = len >= 2;
        int32_t *const cnt_cur = a.touch_cnt + 2 * c + (iter & 1);
        if (cr == 0 && tid == 0) a.touch_cnt[2 * c + ((iter + 1) & 1)] = 0;      // next iteration's counter (idle since iter-1's last barrier)
        const int R = (int)ceilf(r * 1.0001f * a.g.inv_h) + 1;       // cells reached on either side of the vertex's cell
        const int R2 = (2 * R + 1) * (2 * R + 1);
        const long long ntask = (long long)len * R2;
        const int nflat = (int)min((long long)min(len, PATH_SMEM) * R2, (long long)INT_MAX);
        if (len <= SCAN_PATH) {
            // ---- 4s. short route (the common case): claim and resolve in one pass, no atomics (path.py:37-39)
            if (r > 0.f)
                claim_flat<1>(a, base, nc, s_path, len, nflat, R2, R, r, r2, emit ? bid : -1, nullptr, s_queue[warp], s_off, s_beg, s_jj, s_scan, cr, CL);
        } else {
            // ---- 4. claim: every point within r of the path records its nearest path vertex
            if (r > 0.f) {
                claim_flat<0>(a, base, nc, s_path, len, nflat, R2, R, r, r2, -1, cnt_cur, nullptr, s_off, s_beg, s_jj, s_scan, cr, CL);
                for (long long t = (long long)nflat + gwarp; t < ntask; t += nwarp) {      // vertices beyond the staged part of the path
                    const int jj = (int)(t / R2);
                    const int v = base + path[jj];
                    claim_task(a, base, nc, a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], (unsigned)(len - 1 - jj),
                               (int)(t % R2), R, r, r2, lane, cnt_cur);
                }
            }
            cluster_sync_all();
            ST_PHASE(2);
            // ---- 5. resolve: walk the touched list once; a point is on the branch iff it lies inside the radius of
            //         its nearest path vertex (path.py:37-39).
            const int ntouch = __ldcg(cnt_cur);
            for (int k = gtid; k < ntouch; k += nthr) {
                const int gi = __ldcg(a.touched + base + k);
                const unsigned long long bkey = __ldcg(a.best + gi);
                const int jj = len - 1 - (int)(unsigned)(bkey & 0xFFFFFFFFull);
                const float d2 = __uint_as_float((unsigned)(bkey >> 32));
                const float vr = jj < PATH_SMEM ? s_path[jj].w : a.radii[base + path[jj]];
                if (sqrtf(d2) < vr) {
                    a.distw[gi] = -1.f;
                    a.alloc[gi] = 1;
                    if (emit) a.branch_id[gi] = bid;
                }
                __stcg(a.best + gi, BEST_NONE);
            }
        }
        t_b = clock64();
        // ---- 6. allocate the path itself
        for (int jj = gtid; jj < len; jj += nthr) {
            int v = base + path[jj];
            a.distw[v] = -1.f;
            a.alloc[v] = 1;
            if (emit) a.branch_id[v] = bid;
        }
        t_c = clock64();
        cluster_sync_all();
        if (tid == 0 && c == 0 && cr == 0 && iter < 1024) {
            g_st_iter[iter][0] = len; g_st_iter[iter][1] = R; g_st_iter[iter][2] = (int)(clock64() - s_tc); g_st_iter[iter][3] = (int)(clock64() - s_t0);
            g_st_iter[iter][4] = (int)(t_a - s_tc); g_st_iter[iter][5] = (int)(t_b - t_a); g_st_iter[iter][6] = (int)(t_c - t_b); g_st_iter[iter][7] = (int)(clock64() - t_c);
            s_t0 = clock64();
        }
        ST_PHASE(3);
        // ---- 7. emit the branch (root side first); only rank 0 touches the output
        if (emit) {
            if (cr == 0) {
                for (int jj = tid; jj < len / 2; jj += blockDim.x) {
                    int t = out[jj];
                    out[jj] = out[len - 1 - jj];
                    out[len - 1 - jj] = t;
                }
                if (tid == 0) {
                    a.branch_len[base + bid] = len;
                    a.branch_parent[base + bid] = parent;
                }
            }
            ++bid;
            pcur += len;
        }
        ++iter;
        ST_PHASE(4);
    }
    if (tid == 0 && cr == 0) { a.comp_nb[c] = bid; a.comp_np[c] = pcur; }
    if (tid == 0 && c == 0 && cr == 0) { st[7] = CL; for (int i = 0; i < 8; ++i) g_st_stats[i] = st[i]; }
    cluster_sync_all();   // no CTA of the cluster may exit while others still expect it at a barrier
}


// ------------------------------------------------------------------------------------ speculative batches
// The greedy loop above is strictly sequential: 634 iterations of ~9 us on the bench tree, most of them twigs of a few
// vertices whose cost is a handful of dependent L2 round trips and two cluster barriers.  Successive farthest vertices
// are usually far apart (tips of different twigs), so this kernel runs up to CL of them AT ONCE, one per CTA of the
// cluster, and then keeps the longest prefix that provably equals the sequential result:
//   * the next CL live entries of the sorted list are the candidates f_0 .. f_{B-1}; CTA j traces f_j to its first
//     allocated ancestor and collects (without committing) the points its path would claim -- claims do not depend on
//     the allocation state (path.py:30-46 searches all points), the route only through where it stops
//   * every vertex a member touches (path + claimed points) is stamped with the smallest member index that touched it
//   * member j is exactly what the sequential loop would do iff no earlier member touched its start vertex, any vertex
//     of its route, or the vertex whose label becomes its parent id.  A member whose start vertex was claimed by an
//     accepted earlier member is skipped (the sequential loop would never pick it); the first member that fails the
//     test ends the batch -- it and everything after it is retried in the next round.  Member 0 is always valid.
//   * accepted members commit together; labels go through atomicMax (branch ids grow with the member index, so "the
//     latest writer wins" of path.py:135-138 is the maximum)
// Routes of 64 or more vertices are not batched: if candidate 0's is that long, the whole cluster runs the iteration
// of k_sample_tree on it; a long route further back ends the batch before it.
constexpr int BATCH_PATH = 64;
// debug counters of component 0 of the last launch: [rounds, whole-cluster (long route) iterations, batches, members offered,
// accepted, skipped, batches cut by a long member, cut because the start vertex was touched, cut because the route / parent
// vertex was touched, claimed points of accepted members, cycles in batches, cycles in long iterations, ...]
__device__ unsigned long long g_stb_stats[16];
__device__ unsigned long long g_stc_phase[8];     // k_sample_tree_c, CTA 0 of component 0, cycles: B select, C routes, C fill + windows, barrier 1, E verdict inputs, verdict, F commit, barrier 2
constexpr int ST_ACCEPT = 1, ST_SKIP = 2, ST_CUT = 0;

__global__ void __launch_bounds__(1024, 1) k_sample_tree_b(SampleArgs a, const int32_t *__restrict__ jump, int n_total, int32_t *stamp,
                                                           int32_t *tlist_all, int32_t *binfo) {
    const unsigned CL = cluster_size(), cr = cluster_rank();
    const int c = blockIdx.x / CL;
    const int base = a.comp_off[c];
    const int nc = a.comp_off[c + 1] - base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = cr * 32 + warp, nwarp = CL * 32;
    const int gtid = cr * 1024 + tid, nthr = CL * 1024;
    const int B = (int)min(CL, 16u);
    __shared__ int s_first, s_first2, s_term, s_rbits, s_tcnt, s_mf, s_mp, s_stop, s_newbid, s_newpcur;
    __shared__ int s_cand[16], s_status[16], s_bid[16], s_pcur[16];
    __shared__ int s_pv[BATCH_PATH];
    __shared__ float4 s_path[PATH_SMEM];
    __shared__ int2 s_queue[32][QUEUE_LEN];
    __shared__ int s_off[1024], s_beg[1024], s_jj[1024], s_scan[32];
    int32_t *const tlist = tlist_all + (size_t)cr * n_total + base;
    int32_t *const info = binfo + (size_t)c * 64;
    int cursor = 0, bid = 0, pcur = 0, iter = 0, epoch = 1;
    unsigned long long dbg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dbg[i] = 0;
    long long t_mark = clock64();
    while (true) {
        ++dbg[0];
        // ---- A. the next B live entries of the (distance desc, index asc) list, same in every CTA
        int ncand = 0;
        for (int sp = cursor; ncand < B && sp < nc; sp += 1024) {
            const int pos = sp + tid;
            int live = 0;
            if (pos < nc) live = __ldcg(a.distw + __ldg(a.order + base + pos)) > 0.f;
            int total;
            const int rank = block_excl_scan_1024(live, s_scan, total);
            if (live && ncand + rank < B) s_cand[ncand + rank] = pos;
            ncand = min(B, ncand + total);
            __syncthreads();
        }
        if (ncand == 0) break;
        const bool member = (int)cr < ncand;
        const int f0 = __ldg(a.order + base + s_cand[0]) - base;
        const int fm = member ? __ldg(a.order + base + s_cand[cr]) - base : -1;
        // ---- B. probe: ancestors 0..63 of candidate 0 (threads 0..63: is its route long?) and of this CTA's candidate
        //         (threads 64..127); ancestor h < 64 = two octal digits = at most two dependent loads
        if (tid == 0) { s_first = 1024; s_first2 = 1024; s_rbits = 0; s_tcnt = 0; s_mf = 16; s_mp = 16; s_term = -1; }
        __syncthreads();
        int x = -1;
        const int h = tid & 63;
        if (tid < 128) {
            const int src = tid < 64 ? f0 : fm;
            if (src >= 0) {
                x = src;
#pragma unroll
                for (int L = 0; L < 2; ++L) {
                    const int d = (h >> (3 * L)) & 7;
                    if (d && x >= 0) x = __ldg(jump + (size_t)(7 * L + d - 1) * n_total + base + x);
                }
                const bool stop = x < 0 || __ldcg(a.alloc + base + x) != 0;
                if (stop) atomicMin(tid < 64 ? &s_first : &s_first2, h);
            }
        }
        __syncthreads();
        if (s_first == 1024) {
            // =========================== long route: one iteration of k_sample_tree on candidate 0, whole cluster
            const int f = f0;
            int len = 0, cur = f, term = -1;
            float rl = 0.f;
            int *out = a.path_out + base + pcur;
            __syncthreads();
            while (true) {
                if (tid == 0) s_first = 1024;
                __syncthreads();
                int y = cur;
#pragma unroll
                for (int L = 0; L < JUMP_L; ++L) {
                    const int d = (tid >> (3 * L)) & 7;
                    if (d && y >= 0) y = __ldg(jump + (size_t)(7 * L + d - 1) * n_total + base + y);
                }
                const bool stop = y < 0 || __ldcg(a.alloc + base + y) != 0;
                if (stop) atomicMin(&s_first, tid);
                __syncthreads();
                const int first = s_first;
                if (tid == first) s_term = y;
                const int cnt = min(first, max(nc - pcur - len, 0));
                if (tid < cnt) {
                    out[len + tid] = y;
                    const int v = base + y;
                    const float rv = a.radii[v];
                    if (len + tid < PATH_SMEM) s_path[len + tid] = make_float4(a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], rv);
                    rl = fmaxf(rl, rv);
                }
                __syncthreads();
                len += cnt;
                if (first < 1024) { term = s_term; break; }
                if (cnt < 1024) { term = -1; break; }
                cur = __ldg(jump + (size_t)(7 * 3 + 1) * n_total + base + cur);
                __syncthreads();
            }
            const int parent = __ldcg(a.branch_id + base + (term >= 0 ? term : nc - 1));
            const int *path = out;
            for (int o = 16; o; o >>= 1) rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));
            if (lane == 0 && rl > 0.f) atomicMax(&s_rbits, __float_as_int(rl));
            __syncthreads();
            const float r = __int_as_float(s_rbits);
            const float r2 = __fmul_rn(r, r);
            const bool emit = len >= 2;
            int32_t *const cnt_cur = a.touch_cnt + 2 * c + (iter & 1);
            if (cr == 0 && tid == 0) a.touch_cnt[2 * c + ((iter + 1) & 1)] = 0;
            const int R = (int)ceilf(r * 1.0001f * a.g.inv_h) + 1;
            const int R2 = (2 * R + 1) * (2 * R + 1);
            const long long ntask = (long long)len * R2;
            const int nflat = (int)min((long long)min(len, PATH_SMEM) * R2, (long long)INT_MAX);
            if (len <= SCAN_PATH) {
                if (r > 0.f)
                    claim_flat<1>(a, base, nc, s_path, len, nflat, R2, R, r, r2, emit ? bid : -1, nullptr, s_queue[warp], s_off, s_beg, s_jj, s_scan, cr, CL);
            } else {
                if (r > 0.f) {
                    claim_flat<0>(a, base, nc, s_path, len, nflat, R2, R, r, r2, -1, cnt_cur, nullptr, s_off, s_beg, s_jj, s_scan, cr, CL);
                    for (long long t = (long long)nflat + gwarp; t < ntask; t += nwarp) {
                        const int jj = (int)(t / R2);
                        const int v = base + path[jj];
                        claim_task(a, base, nc, a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], (unsigned)(len - 1 - jj),
                                   (int)(t % R2), R, r, r2, lane, cnt_cur);
                    }
                }
                cluster_sync_all();
                const int ntouch = __ldcg(cnt_cur);
                for (int k = gtid; k < ntouch; k += nthr) {
                    const int gi = __ldcg(a.touched + base + k);
                    const unsigned long long bkey = __ldcg(a.best + gi);
                    const int jj = len - 1 - (int)(unsigned)(bkey & 0xFFFFFFFFull);
                    const float d2 = __uint_as_float((unsigned)(bkey >> 32));
                    const float vr = jj < PATH_SMEM ? s_path[jj].w : a.radii[base + path[jj]];
                    if (sqrtf(d2) < vr) {
                        a.distw[gi] = -1.f;
                        a.alloc[gi] = 1;
                        if (emit) a.branch_id[gi] = bid;
                    }
                    __stcg(a.best + gi, BEST_NONE);
                }
            }
            for (int jj = gtid; jj < len; jj += nthr) {
                const int v = base + path[jj];
                a.distw[v] = -1.f;
                a.alloc[v] = 1;
                if (emit) a.branch_id[v] = bid;
            }
            cluster_sync_all();
            if (emit) {
                if (cr == 0) {
                    for (int jj = tid; jj < len / 2; jj += blockDim.x) {
                        const int t = out[jj];
                        out[jj] = out[len - 1 - jj];
                        out[len - 1 - jj] = t;
                    }
                    if (tid == 0) {
                        a.branch_len[base + bid] = len;
                        a.branch_parent[base + bid] = parent;
                    }
                }
                ++bid;
                pcur += len;
            }
            ++iter;
            cursor = s_cand[0] + 1;
            ++dbg[1];
            { const long long t = clock64(); dbg[11] += (unsigned long long)(t - t_mark); t_mark = t; }
            __syncthreads();
            continue;
        }
        // =========================== batch: CTA j works on candidate j
        const int first2 = s_first2;
        const bool active = member && first2 < 1024;      // (a member whose route is long ends the batch: see the resolution)
        const int len = active ? first2 : 0;
        float rl = 0.f;
        if (active && tid >= 64 && tid < 128) {
            if (h < len) {
                s_pv[h] = x;
                const int v = base + x;
                const float rv = a.radii[v];
                s_path[h] = make_float4(a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], rv);
                rl = rv;
            } else if (h == len) {
                s_term = x;
            }
        }
        if (tid >= 64 && tid < 128) {
            for (int o = 16; o; o >>= 1) rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));
            if (lane == 0 && rl > 0.f) atomicMax(&s_rbits, __float_as_int(rl));
        }
        __syncthreads();
        const int term = s_term;
        int parent = -1;
        if (active) parent = __ldcg(a.branch_id + base + (term >= 0 ? term : nc - 1));      // state before this batch (nobody has committed yet)
        const float r = __int_as_float(s_rbits);
        const float r2 = __fmul_rn(r, r);
        const int sval = epoch * 16 + (15 - (int)cr);
        if (active) {
            if (r > 0.f) {
                const int R = (int)ceilf(r * 1.0001f * a.g.inv_h) + 1;
                const int R2 = (2 * R + 1) * (2 * R + 1);
                const Collect co{tlist, &s_tcnt, stamp, sval};
                claim_flat<2>(a, base, nc, s_path, len, len * R2, R2, R, r, r2, -1, nullptr, s_queue[warp], s_off, s_beg, s_jj, s_scan, 0u, 1u, &co);
            }
            if (tid < len) atomicMax(stamp + base + s_pv[tid], sval);
        }
        cluster_sync_all();                                  // 1: every member's stamps are in place
        if (active) {
            if (tid <= len) {
                const int v = tid < len ? s_pv[tid] : (term >= 0 ? term : nc - 1);
                const int sv = __ldcg(stamp + base + v);
                const int who = (sv >> 4) == epoch ? 15 - (sv & 15) : 16;      // smallest member index that touched v in this batch
                if (tid == 0) s_mf = who; else atomicMin(&s_mp, who);
            }
        }
        __syncthreads();
        if (member && tid == 0) {
            __stcg(info + 4 * cr + 0, s_mf);
            __stcg(info + 4 * cr + 1, s_mp);
            __stcg(info + 4 * cr + 2, len);
            __stcg(info + 4 * cr + 3, active ? 1 : 0);
        }
        cluster_sync_all();                                  // 2: every member's verdict inputs are published
        if (tid == 0) {
            int b = bid, p = pcur, stop = ncand;
            for (int j = 0; j < ncand; ++j) {
                const int mf = __ldcg(info + 4 * j), mp = __ldcg(info + 4 * j + 1), lj = __ldcg(info + 4 * j + 2), act = __ldcg(info + 4 * j + 3);
                if (!act) { stop = j; break; }
                if (mf < j) {
                    if (s_status[mf] == ST_ACCEPT) { s_status[j] = ST_SKIP; continue; }      // claimed by an accepted earlier member: never picked
                    stop = j;
                    break;
                }
                if (mp < j) { stop = j; break; }
                s_status[j] = ST_ACCEPT;
                s_bid[j] = b;
                s_pcur[j] = p;
                if (lj >= 2) { ++b; p += lj; }
            }
            for (int j = stop; j < 16; ++j) s_status[j] = ST_CUT;
            s_stop = stop; s_newbid = b; s_newpcur = p;
            ++dbg[2];
            dbg[3] += ncand;
            for (int j = 0; j < stop; ++j) { if (s_status[j] == ST_ACCEPT) ++dbg[4]; else ++dbg[5]; }
            if (stop < ncand) {
                const int mf = __ldcg(info + 4 * stop), mp = __ldcg(info + 4 * stop + 1), act = __ldcg(info + 4 * stop + 3);
                if (!act) ++dbg[6]; else if (mf < stop) ++dbg[7]; else if (mp < stop) ++dbg[8];
            }
        }
        __syncthreads();
        if (active && s_status[cr] == ST_ACCEPT) {
            const bool emit = len >= 2;
            const int mybid = s_bid[cr], mypcur = s_pcur[cr];
            const int ntouch = s_tcnt;
            for (int k = tid; k < ntouch; k += 1024) {
                const int gi = tlist[k];
                a.distw[gi] = -1.f;
                a.alloc[gi] = 1;
                if (emit) atomicMax(a.branch_id + gi, mybid);
            }
            if (tid < len) {
                const int v = base + s_pv[tid];
                a.distw[v] = -1.f;
                a.alloc[v] = 1;
                if (emit) {
                    atomicMax(a.branch_id + v, mybid);
                    a.path_out[base + mypcur + (len - 1 - tid)] = s_pv[tid];      // root side first
                }
            }
            if (emit && tid == 0) {
                a.branch_len[base + mybid] = len;
                a.branch_parent[base + mybid] = parent;
            }
        }
        if (tid == 0) { const long long t = clock64(); dbg[10] += (unsigned long long)(t - t_mark); t_mark = t; }
        const int stop = s_stop;
        bid = s_newbid;
        pcur = s_newpcur;
        cursor = stop < ncand ? s_cand[stop] : s_cand[ncand - 1] + 1;
        ++epoch;
        cluster_sync_all();                                  // 3: commits visible before the next round reads the state
    }
    if (tid == 0 && cr == 0) { a.comp_nb[c] = bid; a.comp_np[c] = pcur; }
    if (tid == 0 && cr == 0 && c == 0) { dbg[15] = CL; for (int i = 0; i < 16; ++i) g_stb_stats[i] = dbg[i]; }
    cluster_sync_all();
}


// ------------------------------------------------------------------------------------ speculative rounds, cluster-cooperative
// Second design of the speculative scheme (k_sample_tree_b above ran one member per CTA: counters on the bench tree showed
// (i) three of four candidates taken in sorted order are claimed by the member just before them -- the vertices next to
// a branch tip all have nearly the tip's distance -- and (ii) the claim of a 40-vertex route on ONE CTA costs more than
// the whole sequential iteration on sixteen).  Here
//   * a round looks at a WINDOW of up to 256 live entries of the sorted list and picks up to 16 members that are
//     mutually far apart (a heuristic: it only decides how much of the round survives, never the result);
//   * every CTA knows every member's route; the claim tasks of all members are laid end to end and spread over all
//     CTAs of the cluster, survivors go to ONE shared list tagged with their member;
//   * every CTA then derives the verdicts itself from the stamps (no exchange): member g stands iff no earlier member
//     touched its start vertex, route or parent vertex; a window entry that was passed over must have been claimed by an
//     ACCEPTED member before it in the list -- otherwise the sequential loop would have picked it, and the round ends
//     right before it;
//   * accepted members commit together (labels through atomicMax: branch ids grow with the member index).
// Two cluster barriers per round.  Routes of 64 or more vertices run as one whole-cluster iteration of k_sample_tree.
constexpr int MB = 16;         // members per round
constexpr int MP = 128;        // route slots per member (two per thread of the member's 64)
constexpr int MPS = 7;         // log2(MP)
constexpr int WIN = 512;       // live entries examined per round
constexpr int SCAN_PPT = 8;    // list positions per thread and scan step (the loads of a step are in flight together: a step costs two L2 round trips + a block scan whatever its width)
constexpr int SCAN_STEPS = 4;  // scan steps of 1024 * SCAN_PPT list positions per round at most (the list is compacted as the loop goes)

struct RoundSmem {
    float4 path[MB * MP];      // xyz + radius of member g's route vertex h at [g * MP + h]  (the long-route path reuses it as [1024])
    int2 queue[32][QUEUE_LEN];
    int off[1024], beg[1024], jj[1024];
    int pv[MB * MP];           // vertex ids (component-local) of the routes
    int lohi[MB * MP];         // path_windows of every route vertex (member-local indices)
    int win_pos[WIN + 1];      // positions (in the sorted list) of the window's live entries (+ the first one beyond it)
    int win_gap[WIN];          // >= 0: member index of a selected entry;  < 0: -(g + 1), passed over after member g
    float4 win_pt[WIN];        // xyz + radius of the window entries (member selection)
    int win_v[WIN];            // vertex id (global) of the window entries: `order` is compacted in place at the end of the round
    float4 selpt[MB];          // start vertex of each member: xyz + radius
    int sel[MB], len[MB], term[MB], rbits[MB], parent[MB], mf[MB], mp[MB], status[MB], bid[MB], pcur[MB], toff[MB + 1], R[MB];
    int gapneed[MB], gapbad[MB];
    int scan[32];
    int first[MB];
    int nsel, nwin, stop_entry, newbid, newpcur, minpos_dummy, first_long, term_long, rbits_long;
};

__device__ __forceinline__ int stamp_who(int sv, int epoch) { return (sv >> 4) == epoch ? 15 - (sv & 15) : 16; }

__global__ void __launch_bounds__(1024, 1) k_sample_tree_c(SampleArgs a, const int32_t *__restrict__ jump, int n_total, int32_t *stamp,
                                                           int32_t *clist_all, int32_t *ccnt_all, int win, int scan_steps, float sep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RoundSmem &S = *reinterpret_cast<RoundSmem *>(smem_raw);
    const unsigned CL = cluster_size(), cr = cluster_rank();
    const int c = blockIdx.x / CL;
    const int base = a.comp_off[c];
    const int nc = a.comp_off[c + 1] - base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = cr * 32 + warp, nwarp = CL * 32;
    const int gtid = cr * 1024 + tid, nthr = CL * 1024;
    int32_t *const clist = clist_all + (size_t)16 * base;      // claimed (member << 28 | local vertex), capacity 16 * nc
    int32_t *const ccnt = ccnt_all + 2 * c;                     // two counters, alternating by round parity
    int cursor = 0, bid = 0, pcur = 0, iter = 0, epoch = 1;
    unsigned long long dbg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dbg[i] = 0;
    unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_f = 0;
#ifdef ST_SAMPLE_TRACE      // per-phase cycle counters of the rounds (tools/sample_sweep.py; build with ST_NVCC_EXTRA=-DST_SAMPLE_TRACE)
#define ST_PH(i) do { if (tid == 0) { const long long t_ = clock64(); ph[i] += (unsigned long long)(t_ - t_f); t_f = t_; } } while (0)
#else
#define ST_PH(i) do { (void)t_f; } while (0)
#endif
    long long t_mark = clock64();
    while (true) {
        ++dbg[0];
        // ---- A. window: the next live entries of the (distance desc, index asc) list, scanned 1024 * SCAN_PPT positions at a
        //         time
        long long t_ph = clock64();
        int nlive = 0, scan_end = cursor;
        for (int step = 0; step < scan_steps && nlive <= win && scan_end < nc; ++step) {
            // a warp takes 32 * SCAN_PPT consecutive positions, lane-contiguous (one line per load instruction); the rank of a
            // live entry = live entries of the warps before + of the warp's earlier load rounds + of the lower lanes
            const int wbase = scan_end + warp * (32 * SCAN_PPT);
            int vtx[SCAN_PPT];
            unsigned masks[SCAN_PPT];
#pragma unroll
            for (int k = 0; k < SCAN_PPT; ++k) {
                const int p = wbase + k * 32 + lane;
                vtx[k] = p < nc ? __ldcg(a.order + base + p) : -1;
            }
            int wtotal = 0;
#pragma unroll
            for (int k = 0; k < SCAN_PPT; ++k) {
                const bool live = vtx[k] >= 0 && __ldcg(a.distw + vtx[k]) > 0.f;      // (-1 = tombstone: no look-up)
                masks[k] = __ballot_sync(0xffffffffu, live);
                wtotal += __popc(masks[k]);
            }
            if (lane == 0) S.scan[warp] = wtotal;
            __syncthreads();
            const int x = S.scan[lane];
            int incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int rank = nlive + __shfl_sync(0xffffffffu, incl - x, warp);
#pragma unroll
            for (int k = 0; k < SCAN_PPT; ++k) {
                if ((masks[k] >> lane) & 1u) {
                    const int r = rank + __popc(masks[k] & ((1u << lane) - 1u));
                    if (r <= win) { S.win_pos[r] = wbase + k * 32 + lane; if (r < win) S.win_v[r] = vtx[k]; }
                }
                rank += __popc(masks[k]);
            }
            nlive += total;
            __syncthreads();
            scan_end = min(scan_end + 1024 * SCAN_PPT, nc);
        }
        if (nlive == 0) {
            if (scan_end >= nc) break;
            cursor = scan_end;
            continue;
        }
        int nwin = min(nlive, win);
        // position where the window ends: the first live entry beyond it, or the end of the scanned range
        int win_end = nlive > win ? S.win_pos[win] : scan_end;
        if (tid < nwin) {
            const int v = S.win_v[tid];
            S.win_pt[tid] = make_float4(a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], a.radii[v]);
        }
        __syncthreads();
        if (tid == 0) { const long long t = clock64(); dbg[12] += (unsigned long long)(t - t_ph); t_ph = t; t_f = t; }
        // ---- B. members: greedy in list order, an entry is taken iff it is far from every member taken so far (warp 0)
        if (warp == 0) {
            int nsel = 0;
            for (int e0 = 0; e0 < nwin && nsel < MB; e0 += 32) {
                const int e = e0 + lane;
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                bool far = false;
                if (e < nwin) {
                    p = S.win_pt[e];
                    far = true;
                    for (int m = 0; m < nsel; ++m) {
                        const float4 q = S.selpt[m];
                        const float rho = sep * fmaxf(p.w, q.w);
                        if (dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < rho * rho) far = false;
                    }
                }
                int gap = nsel - 1;                       // member index the passed-over entries of this chunk come after
                unsigned todo = __ballot_sync(0xffffffffu, far);
                int mine = -1;                            // member index if this lane's entry gets selected
                int my_gap = gap;
                while (todo && nsel < MB) {
                    const int l = __ffs(todo) - 1;
                    todo &= ~(1u << l);
                    const float4 q = make_float4(__shfl_sync(0xffffffffu, p.x, l), __shfl_sync(0xffffffffu, p.y, l),
                                                 __shfl_sync(0xffffffffu, p.z, l), __shfl_sync(0xffffffffu, p.w, l));
                    if (lane == l) { mine = nsel; S.selpt[nsel] = p; S.sel[nsel] = e; }
                    // lanes after l that were still candidates re-test against the new member; lanes after l come after it in the list
                    if (lane > l) {
                        my_gap = nsel;
                        const float rho = sep * fmaxf(p.w, q.w);
                        if (far && dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < rho * rho) far = false;
                    }
                    todo &= __ballot_sync(0xffffffffu, far);
                    ++nsel;
                }
                // lanes whose entry lies beyond the MB-th member are outside the round (handled by truncating nwin below)
                if (e < nwin) S.win_gap[e] = mine >= 0 ? mine : -(my_gap + 1);
                __syncwarp();
            }
            if (lane == 0) S.nsel = nsel;
        }
        __syncthreads();
        ST_PH(0);
        int nsel = S.nsel;
        if (nsel == MB && S.sel[MB - 1] + 1 < nwin) {           // the round ends right after its last member
            nwin = S.sel[MB - 1] + 1;
            win_end = S.win_pos[nwin];
        }
        // ---- C. routes: thread (g, h) resolves ancestor h of member g (two octal digits = at most two dependent loads)
        const int g = tid >> 6, h = tid & 63;
        if (tid < MB) { S.first[tid] = 1024; S.rbits[tid] = 0; S.mf[tid] = 16; S.mp[tid] = 16; S.term[tid] = -1; S.gapneed[tid] = 0; S.gapbad[tid] = INT_MAX; }
        __syncthreads();
        int xs[2] = {-1, -1};                               // ancestors h and h + 64 of member g
        if (g < nsel) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int hh = h + 64 * t;
                int x = S.win_v[S.sel[g]] - base;
#pragma unroll
                for (int L = 0; L < 3; ++L) {
                    const int d = (hh >> (3 * L)) & 7;
                    if (d && x >= 0) x = __ldg(jump + (size_t)(7 * L + d - 1) * n_total + base + x);
                }
                xs[t] = x;
                const bool stop = x < 0 || __ldcg(a.alloc + base + x) != 0;
                if (stop) atomicMin(&S.first[g], hh);
            }
        }
        __syncthreads();
        ST_PH(1);
        if (S.first[0] == 1024) {
            // =========================== long route: one iteration of k_sample_tree on candidate 0, whole cluster
            const int f = S.win_v[0] - base;
            int len = 0, cur = f, term = -1;
            float rl = 0.f;
            int *out = a.path_out + base + pcur;
            if (tid == 0) { S.rbits_long = 0; S.term_long = -1; }
            __syncthreads();
            while (true) {
                if (tid == 0) S.first_long = 1024;
                __syncthreads();
                int y = cur;
#pragma unroll
                for (int L = 0; L < JUMP_L; ++L) {
                    const int d = (tid >> (3 * L)) & 7;
                    if (d && y >= 0) y = __ldg(jump + (size_t)(7 * L + d - 1) * n_total + base + y);
                }
                const bool stop = y < 0 || __ldcg(a.alloc + base + y) != 0;
                if (stop) atomicMin(&S.first_long, tid);
                __syncthreads();
                const int first = S.first_long;
                if (tid == first) S.term_long = y;
                const int cnt = min(first, max(nc - pcur - len, 0));
                if (tid < cnt) {
                    out[len + tid] = y;
                    const int v = base + y;
                    const float rv = a.radii[v];
                    if (len + tid < PATH_SMEM) S.path[len + tid] = make_float4(a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], rv);
                    rl = fmaxf(rl, rv);
                }
                __syncthreads();
                len += cnt;
                if (first < 1024) { term = S.term_long; break; }
                if (cnt < 1024) { term = -1; break; }
                cur = __ldg(jump + (size_t)(7 * 3 + 1) * n_total + base + cur);
                __syncthreads();
            }
            const int parent = __ldcg(a.branch_id + base + (term >= 0 ? term : nc - 1));
            const int *path = out;
            for (int o = 16; o; o >>= 1) rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));
            if (lane == 0 && rl > 0.f) atomicMax(&S.rbits_long, __float_as_int(rl));
            __syncthreads();
            const float r = __int_as_float(S.rbits_long);
            const float r2 = __fmul_rn(r, r);
            const bool emit = len >= 2;
            int32_t *const cnt_cur = a.touch_cnt + 2 * c + (iter & 1);
            if (cr == 0 && tid == 0) a.touch_cnt[2 * c + ((iter + 1) & 1)] = 0;
            const int R = (int)ceilf(r * 1.0001f * a.g.inv_h) + 1;
            const int R2 = (2 * R + 1) * (2 * R + 1);
            const long long ntask = (long long)len * R2;
            const int nflat = (int)min((long long)min(len, PATH_SMEM) * R2, (long long)INT_MAX);
            if (len <= SCAN_PATH) {
                if (r > 0.f) {
                    path_windows(S.path, len, r, S.lohi, tid, 1024);
                    __syncthreads();
                    claim_flat<1>(a, base, nc, S.path, len, nflat, R2, R, r, r2, emit ? bid : -1, nullptr, S.queue[warp], S.off, S.beg, S.jj, S.scan, cr, CL,
                                  nullptr, S.lohi);
                }
            } else {
                if (r > 0.f) {
                    claim_flat<0>(a, base, nc, S.path, len, nflat, R2, R, r, r2, -1, cnt_cur, nullptr, S.off, S.beg, S.jj, S.scan, cr, CL);
                    for (long long t = (long long)nflat + gwarp; t < ntask; t += nwarp) {
                        const int jj = (int)(t / R2);
                        const int v = base + path[jj];
                        claim_task(a, base, nc, a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], (unsigned)(len - 1 - jj),
                                   (int)(t % R2), R, r, r2, lane, cnt_cur);
                    }
                }
                cluster_sync_all();
                const int ntouch = __ldcg(cnt_cur);
                for (int k = gtid; k < ntouch; k += nthr) {
                    const int gi = __ldcg(a.touched + base + k);
                    const unsigned long long bkey = __ldcg(a.best + gi);
                    const int jj = len - 1 - (int)(unsigned)(bkey & 0xFFFFFFFFull);
                    const float d2 = __uint_as_float((unsigned)(bkey >> 32));
                    const float vr = jj < PATH_SMEM ? S.path[jj].w : a.radii[base + path[jj]];
                    if (sqrtf(d2) < vr) {
                        kill_vertex(a, gi, 0);
                        if (emit) a.branch_id[gi] = bid;
                    }
                    __stcg(a.best + gi, BEST_NONE);
                }
            }
            for (int jj = gtid; jj < len; jj += nthr) {
                const int v = base + path[jj];
                kill_vertex(a, v, 0);
                if (emit) a.branch_id[v] = bid;
            }
            cluster_sync_all();
            if (emit) {
                if (cr == 0) {
                    for (int jj = tid; jj < len / 2; jj += blockDim.x) {
                        const int t = out[jj];
                        out[jj] = out[len - 1 - jj];
                        out[len - 1 - jj] = t;
                    }
                    if (tid == 0) {
                        a.branch_len[base + bid] = len;
                        a.branch_parent[base + bid] = parent;
                    }
                }
                ++bid;
                pcur += len;
            }
            ++iter;
            cursor = S.win_pos[0] + 1;
            ++dbg[1];
            { const long long t = clock64(); dbg[11] += (unsigned long long)(t - t_mark); t_mark = t; }
            __syncthreads();
            continue;
        }
        // =========================== round: a long member ends the round before it
        for (int m = 1; m < nsel; ++m)
            if (S.first[m] == 1024) { nsel = m; nwin = S.sel[m]; win_end = S.win_pos[nwin]; ++dbg[6]; break; }
        const int epoch_tag = epoch * 16;
        float rl = 0.f;
        if (g < nsel) {
            const int len = S.first[g];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int hh = h + 64 * t, x = xs[t];
                if (hh < len) {
                    S.pv[g * MP + hh] = x;
                    const int v = base + x;
                    const float rv = a.radii[v];
                    S.path[g * MP + hh] = make_float4(a.pts[3 * (size_t)v], a.pts[3 * (size_t)v + 1], a.pts[3 * (size_t)v + 2], rv);
                    rl = fmaxf(rl, rv);
                    if ((unsigned)(g % (int)CL) == cr) atomicMax(stamp + v, epoch_tag + 15 - g);      // the route stamps its own vertices
                } else if (hh == len) {
                    S.term[g] = x;
                }
            }
            if (h == 0) S.len[g] = len;
        }
        for (int o = 16; o; o >>= 1) rl = fmaxf(rl, __shfl_xor_sync(0xffffffffu, rl, o));       // a warp lies inside one member's 64 threads
        if (lane == 0 && rl > 0.f && g < nsel) atomicMax(&S.rbits[g], __float_as_int(rl));
        __syncthreads();
        if (tid < nsel) {
            const int term = S.term[tid];
            S.parent[tid] = __ldcg(a.branch_id + base + (term >= 0 ? term : nc - 1));      // state before this round (nobody has committed yet)
        }
        if (g < nsel) path_windows(S.path + g * MP, S.len[g], __int_as_float(S.rbits[g]), S.lohi + g * MP, h, 64);
        if (tid == 0) {
            int t = 0;
            for (int m = 0; m < nsel; ++m) {
                const float r = __int_as_float(S.rbits[m]);
                const int R = r > 0.f ? (int)ceilf(r * 1.0001f * a.g.inv_h) + 1 : 0;
                S.R[m] = R;
                S.toff[m] = t;
                t += r > 0.f ? S.len[m] * (2 * R + 1) * (2 * R + 1) : 0;
            }
            S.toff[nsel] = t;
            if (cr == 0) ccnt[(epoch + 1) & 1] = 0;      // next round's list counter (idle since the previous round's last barrier)
        }
        __syncthreads();
        if (tid == 0) { const long long t = clock64(); dbg[13] += (unsigned long long)(t - t_ph); t_ph = t; }
        ST_PH(2);
        // ---- D. claim, all members at once, tasks interleaved over the CTAs of the cluster
        int32_t *const cnt_cur = ccnt + (epoch & 1);
        {
            const int T = S.toff[nsel];
            int qn = 0;
            int2 *queue = S.queue[warp];
            const float4 none = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            auto drain = [&](int first, int count) {
                bool hit = false;
                int word = 0, gi = -1, m = 0;
                if (lane < count) {
                    const int2 e = queue[first + lane];
                    m = e.y >> MPS;
                    const int j = e.y & (MP - 1), len = S.len[m];
                    const float4 q = __ldg(a.sorted + e.x);
                    const float4 p = S.path[e.y];
                    const float d2 = dist2_exact(q.x, q.y, q.z, p.x, p.y, p.z);
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(len - 1 - j);
                    bool own = true;
                    const int w = S.lohi[e.y];
                    for (int j2 = w & 0xFFFF; j2 <= (w >> 16); ++j2) own = own && !(claim_key(q, S.path[m * MP + j2], (unsigned)(len - 1 - j2)) < key);
                    hit = own && sqrtf(d2) < p.w;
                    gi = __float_as_int(q.w);
                    word = (m << 28) | (gi - base);
                }
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (!mask) return;
                int pos = 0;
                if (lane == 0) pos = atomicAdd(cnt_cur, __popc(mask));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (hit) {
                    clist[pos + __popc(mask & ((1u << lane) - 1u))] = word;
                    atomicMax(stamp + gi, epoch_tag + 15 - m);
                }
            };
            for (int tb = 0; (long long)tb * CL < T; tb += 1024) {
                const long long t = (long long)(tb + tid) * CL + cr;
                int beg = 0, end = 0, slot = 0;
                if (t < T) {
                    int m = 0;
                    while (m + 1 < nsel && S.toff[m + 1] <= t) ++m;
                    const int R = S.R[m], R2 = (2 * R + 1) * (2 * R + 1);
                    const int local = (int)(t - S.toff[m]);
                    const int jj = local / R2;
                    slot = m * MP + jj;
                    // rows within the vertex' OWN radius (see claim_flat): the grid of rows is laid out for the route's largest
                    row_range(a, S.path[slot], local % R2, R, S.path[slot].w * 1.0001f + 1e-7f, beg, end);
                }
                int total;
                const int ex = block_excl_scan_1024(end > beg ? end - beg : 0, S.scan, total);
                S.off[tid] = ex; S.beg[tid] = beg; S.jj[tid] = slot;
                __syncthreads();
                auto locate = [&](int e, int &tp, int &sl) -> float4 {
                    if (e >= total) { tp = 0; sl = 0; return none; }
                    int lo = 0, hi = 1023;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (S.off[mid] <= e) lo = mid; else hi = mid - 1;
                    }
                    tp = S.beg[lo] + (e - S.off[lo]);
                    sl = S.jj[lo];
                    return __ldg(a.sorted + tp);
                };
                int tpn, sln;
                float4 qnext = locate(tid, tpn, sln);
                for (int e0 = 0; e0 < total; e0 += 1024) {
                    const float4 q = qnext;
                    const int tp = tpn, sl = sln;
                    qnext = locate(e0 + 1024 + tid, tpn, sln);
                    const int gi = __float_as_int(q.w);
                    bool cand = false;
                    if (gi >= base && gi < base + nc) {
                        const int m = sl >> MPS, j = sl & (MP - 1), len = S.len[m];
                        const float4 p = S.path[sl];
                        const float d2 = dist2_exact(q.x, q.y, q.z, p.x, p.y, p.z);
                        const float rj = p.w * 1.0001f + 1e-7f;
                        if (d2 <= rj * rj) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(len - 1 - j);
                            cand = true;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int j2 = j + (k < 2 ? k - 2 : k - 1);
                                if (j2 >= 0 && j2 < len && claim_key(q, S.path[m * MP + j2], (unsigned)(len - 1 - j2)) < key) cand = false;
                            }
                        }
                    }
                    const unsigned mk = __ballot_sync(0xffffffffu, cand);
                    if (cand) queue[qn + __popc(mk & ((1u << lane) - 1u))] = make_int2(tp, sl);
                    qn += __popc(mk);
                    __syncwarp();
                    if (qn >= 32) {
                        qn -= 32;
                        drain(qn, 32);
                        __syncwarp();
                    }
                }
                __syncthreads();
            }
            __syncwarp();
            drain(0, qn);
        }
        if (tid == 0) { const long long t = clock64(); dbg[14] += (unsigned long long)(t - t_ph); t_ph = t; t_f = t; }
        cluster_sync_all();                                  // 1: every claim and stamp of the round is in place
        ST_PH(3);
        // ---- E. verdict inputs, every CTA for itself: who touched the members' vertices and the passed-over entries
        if (g < nsel) {
            const int len = S.len[g];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int hh = h + 64 * t;
                if (hh <= len) {
                    const int term = S.term[g];
                    const int v = hh < len ? S.pv[g * MP + hh] : (term >= 0 ? term : nc - 1);
                    const int who = stamp_who(__ldcg(stamp + base + v), epoch);
                    if (hh == 0) S.mf[g] = who; else atomicMin(&S.mp[g], who);
                }
            }
        }
        if (tid < nwin && S.win_gap[tid] < 0) {
            const int after = -S.win_gap[tid] - 1;          // passed over after member `after`
            const int v = S.win_v[tid];
            const int who = stamp_who(__ldcg(stamp + v), epoch);
            if (who <= after) atomicOr(&S.gapneed[after], 1 << who);
            else atomicMin(&S.gapbad[after], tid);           // nobody before it claimed it: the sequential loop would pick it
        }
        __syncthreads();
        ST_PH(4);
        if (tid == 0) {
            int b = bid, p = pcur, stop_entry = nwin;
            unsigned acc = 0;
            int m = 0;
            for (; m < nsel; ++m) {
                const int mf = S.mf[m], mp = S.mp[m];
                if (mf < m) {
                    if (acc & (1u << mf)) S.status[m] = ST_SKIP;      // claimed by an accepted earlier member: never picked
                    else { stop_entry = S.sel[m]; ++dbg[7]; break; }
                } else if (mp < m) {
                    stop_entry = S.sel[m]; ++dbg[8];
                    break;
                } else {
                    S.status[m] = ST_ACCEPT;
                    acc |= 1u << m;
                    S.bid[m] = b;
                    S.pcur[m] = p;
                    if (S.len[m] >= 2) { ++b; p += S.len[m]; }
                    ++dbg[4];
                }
                // the entries passed over after member m must all have been claimed by accepted members
                if (S.gapbad[m] != INT_MAX || (S.gapneed[m] & ~acc)) {
                    stop_entry = S.sel[m] + 1; ++dbg[9];
                    ++m;
                    break;
                }
            }
            for (int k = m; k < MB; ++k) S.status[k] = ST_CUT;
            S.stop_entry = stop_entry; S.newbid = b; S.newpcur = p;
            ++dbg[2];
            dbg[3] += nsel;
            dbg[5] += nwin;
        }
        __syncthreads();
        ST_PH(5);
        // ---- F. commit, whole cluster
        {
            const int ntouch = __ldcg(cnt_cur);
            for (int k = gtid; k < ntouch; k += nthr) {
                const int word = __ldcg(clist + k);
                const int m = (word >> 28) & 15, v = base + (word & 0x0FFFFFFF);
                if (S.status[m] == ST_ACCEPT) {
                    kill_vertex(a, v, base + win_end);
                    if (S.len[m] >= 2) atomicMax(a.branch_id + v, S.bid[m]);
                }
            }
            if (g < nsel && S.status[g] == ST_ACCEPT && (unsigned)(g % (int)CL) == cr) {
                const int len = S.len[g];
                const bool emit = len >= 2;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int hh = h + 64 * t;
                    if (hh < len) {
                        const int lv = S.pv[g * MP + hh];
                        const int v = base + lv;
                        kill_vertex(a, v, base + win_end);
                        if (emit) {
                            atomicMax(a.branch_id + v, S.bid[g]);
                            a.path_out[base + S.pcur[g] + (len - 1 - hh)] = lv;      // root side first
                        }
                    }
                }
                if (emit && h == 0) {
                    a.branch_len[base + S.bid[g]] = len;
                    a.branch_parent[base + S.bid[g]] = S.parent[g];
                }
            }
        }
        if (tid == 0) { const long long t = clock64(); dbg[10] += (unsigned long long)(t - t_mark); t_mark = t; }
        const int stop_entry = S.stop_entry;
        bid = S.newbid;
        pcur = S.newpcur;
        // The window entries from the cut onwards are the only live entries of [their position, win_end): pack them against
        // win_end (they only move right, in order), so that the next round's scan starts on a dense run of live entries
        // instead of re-reading the dead ones.  Nobody reads the list before the barrier below (win_v is the round's copy).
        const int keep = nwin - stop_entry;
        if (cr == 0 && tid < keep) {
            a.order[base + win_end - keep + tid] = S.win_v[stop_entry + tid];
            if (a.pos) a.pos[S.win_v[stop_entry + tid]] = base + win_end - keep + tid;
        }
        cursor = win_end - keep;
        ++epoch;
        ST_PH(6);
        cluster_sync_all();                                  // 2: commits visible before the next round reads the state
        ST_PH(7);
    }
    if (tid == 0 && cr == 0) { a.comp_nb[c] = bid; a.comp_np[c] = pcur; }
    if (tid == 0 && cr == 0 && c == 0) { dbg[15] = CL; for (int i = 0; i < 16; ++i) g_stb_stats[i] = dbg[i]; for (int i = 0; i < 8; ++i) g_stc_phase[i] = ph[i]; }
#undef ST_PH
    cluster_sync_all();
}

// positions of the vertices in the sorted list; entries that are dead from the start (path.py:71-72) become tombstones
__global__ void k_st_positions(int32_t *__restrict__ order, const float *__restrict__ distw, int n, int32_t *__restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = order[i];
    pos[v] = i;
    if (!(distw[v] > 0.f)) order[i] = -1;
}

static size_t sort_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)n);
    return b;
}

extern "C" size_t st_sample_tree_workspace_bytes(int64_t n, int32_t n_comp) {
    return grid_ws_bytes(n) + align_up(sort_bytes(n)) + 2 * align_up(n * 8) + 2 * align_up(n * 4) + align_up(n * 4) + align_up(n) +
           align_up(n * 4) + align_up(n * 8) + align_up((size_t)JUMP_LEVELS * n * 4) + 2 * align_up(n * 4) + align_up((2 * (size_t)n_comp + 2) * 4) +
           align_up(n * 4) + align_up((size_t)16 * n * 4) + align_up((size_t)n_comp * 64 * 4) + align_up(n * 4) + 8192;      // batches: stamps, touched lists, verdicts, list positions
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
    int32_t *touched = cv.take<int32_t>(n);
    int32_t *touch_cnt = cv.take<int32_t>(2 * (size_t)n_comp + 2);
    int32_t *jump = cv.take<int32_t>((size_t)JUMP_LEVELS * n);
    int32_t *vbase = cv.take<int32_t>(n);
    int32_t *stamp = cv.take<int32_t>(n);
    int32_t *tlist_all = cv.take<int32_t>((size_t)16 * n);
    int32_t *binfo = cv.take<int32_t>((size_t)n_comp * 64);
    int32_t *pos = cv.take<int32_t>(n);
    size_t sb = sort_bytes(n);
    void *sort_ws = cv.take<char>(sb);
    if (!cv.ok()) { set_error("st_sample_tree: workspace too small"); return ST_ERR_WORKSPACE; }
    // schedule of the greedy loop (same result either way): speculative rounds spread over the cluster (default), one
    // speculative member per CTA (ST_SAMPLE_MODE=members), or strictly sequential (ST_SAMPLE_SEQUENTIAL=1)
    int mode = 2;
    if (const char *e = getenv("ST_SAMPLE_MODE")) mode = !strcmp(e, "members") ? 1 : !strcmp(e, "sequential") ? 0 : 2;
    if (getenv("ST_SAMPLE_SEQUENTIAL")) mode = 0;
    if (mode) {
        ST_CHECK_CUDA(cudaMemsetAsync(stamp, 0, n * sizeof(int32_t), s));
        ST_CHECK_CUDA(cudaMemsetAsync(binfo, 0, (size_t)n_comp * 64 * sizeof(int32_t), s));
    }
    k_st_init<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(pred, tree_dist, (int)n, distw, alloc, branch_id, best, comp_off, n_comp, keys, vals);
    ST_CHECK_LAUNCH();
    int key_bits = 33;                                   // 32 distance bits + the component index above them
    while (key_bits < 64 && (1ll << (key_bits - 32)) < (long long)n_comp) ++key_bits;
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_ws, sb, keys, keys2, vals, order, (int)n, 0, key_bits, s));
    GridBuild gb;
    int rc = build_grid(medial_pts, n, cell_size, cv, gb, s);
    if (rc) return rc;
    // binary-lifting table over the (static) predecessor tree
    k_vertex_base<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(comp_off, n_comp, (int)n, vbase);
    ST_CHECK_LAUNCH();
    for (int L = 0; L < JUMP_L; ++L) {
        k_jump_first<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(pred, jump, vbase, (int)n, L);
        ST_CHECK_LAUNCH();
        k_jump_rest<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(jump, vbase, (int)n, L);
        ST_CHECK_LAUNCH();
    }
    SampleArgs a{medial_pts, radii, pred, comp_off, gb.g, gb.cell_start, gb.sorted, order, distw, alloc, branch_id, best,
                 path_vertices, branch_len, branch_parent, comp_n_branches, comp_n_path, touched, touch_cnt, nullptr};
    ST_CHECK_CUDA(cudaMemsetAsync(touch_cnt, 0, (2 * (size_t)n_comp + 2) * sizeof(int32_t), s));
    // largest cluster the device can co-schedule: more CTAs per component = more lanes on the scans
    static int cluster_cached = 0;
    int CL = cluster_cached;
    if (CL == 0) {
        cudaFuncSetAttribute((const void *)k_sample_tree, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute((const void *)k_sample_tree_b, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute((const void *)k_sample_tree_c, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute((const void *)k_sample_tree_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RoundSmem));
        for (int cand : {16, 8, 4, 2, 1}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cand);
            cfg.blockDim = dim3(1024);
            cudaLaunchAttribute at;
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = cand; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            int nclusters = 0;
            int nclusters_b = 0, nclusters_c = 0;
            cudaLaunchConfig_t cfg_c = cfg;
            cfg_c.dynamicSmemBytes = sizeof(RoundSmem);
            if (cudaOccupancyMaxActiveClusters(&nclusters, (const void *)k_sample_tree, &cfg) == cudaSuccess && nclusters >= 1 &&
                cudaOccupancyMaxActiveClusters(&nclusters_b, (const void *)k_sample_tree_b, &cfg) == cudaSuccess && nclusters_b >= 1 &&
                cudaOccupancyMaxActiveClusters(&nclusters_c, (const void *)k_sample_tree_c, &cfg_c) == cudaSuccess && nclusters_c >= 1) { CL = cand; break; }
        }
        cudaGetLastError();
        if (CL == 0) CL = 1;
        cluster_cached = CL;
    }
    if (const char *e = getenv("ST_SAMPLE_CLUSTER")) {      // smaller cluster (= smaller speculative batch): tests
        int v = atoi(e);
        if (v >= 1 && v <= CL && (v & (v - 1)) == 0) CL = v;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_comp * CL);
    cfg.blockDim = dim3(1024);
    cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nt = (int)n;
    const int32_t *jump_c = jump;
    if (mode == 2) {
        ST_REQUIRE(n < (1ll << 28), "components of at most 2^28 vertices");
        if (!getenv("ST_SAMPLE_NO_TOMBSTONES")) {
            k_st_positions<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(order, distw, (int)n, pos);
            ST_CHECK_LAUNCH();
            a.pos = pos;
        }
        cfg.dynamicSmemBytes = sizeof(RoundSmem);
        // live entries examined per round / scan steps of 8192 list positions per round: any value gives the same result
        int win = 128, scan_steps = SCAN_STEPS;      // tools/sample_sweep.py: 128 live entries per round were best on the bench tree (4.7 ms; 512: 6.0 ms)
        if (const char *e = getenv("ST_SAMPLE_WIN")) { int v = atoi(e); if (v >= 1 && v <= WIN) win = v; }
        if (const char *e = getenv("ST_SAMPLE_SCAN_STEPS")) { int v = atoi(e); if (v >= 1 && v <= 64) scan_steps = v; }
        // members of a round are at least sep * max(radius) apart (a heuristic: it only decides how much of the round survives)
        float sep = 1.f;      // tools/sample_sweep.py: 2.0 -> 158 rounds / 99 cut by a passed-over entry, 1.0 -> 112 / 19
        if (const char *e = getenv("ST_SAMPLE_SEP")) { float v = (float)atof(e); if (v >= 0.f) sep = v; }
        ST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_sample_tree_c, a, jump_c, nt, stamp, tlist_all, binfo, win, scan_steps, sep));
    } else if (mode == 1) {
        ST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_sample_tree_b, a, jump_c, nt, stamp, tlist_all, binfo));
    } else {
        ST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_sample_tree, a, jump_c, nt));
    }
    return ST_OK;
}

// debug only (not part of include/st_b200.h): cycles spent by component 0 in
// [find, trace, claim, resolve, finish], iterations, traced path vertices
extern "C" int st_debug_sample_iters(int *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_st_iter, sizeof(int) * 8192));
    return ST_OK;
}
extern "C" int st_debug_sample_batch_stats(unsigned long long *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_stb_stats, sizeof(unsigned long long) * 16));
    return ST_OK;
}
extern "C" int st_debug_sample_round_phases(unsigned long long *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_stc_phase, sizeof(unsigned long long) * 8));
    return ST_OK;
}
extern "C" int st_debug_sample_stats(unsigned long long *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_st_stats, sizeof(unsigned long long) * 8));
    return ST_OK;
}
