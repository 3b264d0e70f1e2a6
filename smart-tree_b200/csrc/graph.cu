// Connected components (lock-free union-find, root = smallest vertex id), CSR build,
// single/multi-source shortest paths (pull-based fp32 relaxation to the least fixed point in a
// resident persistent grid), deterministic predecessors, tree path lengths, tube projection.
#include <cub/cub.cuh>

#include "common.cuh"

using namespace st;

// ------------------------------------------------------------------------------------ connected components
__device__ __forceinline__ int uf_find(int32_t *parent, int v) {
    int p = parent[v];
    while (p != v) {
        int gp = parent[p];
        if (gp != p) parent[v] = gp;  // path halving (benign race: only ever moves towards the root)
        v = p;
        p = gp;
    }
    return v;
}

__global__ void k_cc_init(int32_t *parent, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = i;
}

__global__ void k_cc_hook(const int32_t *__restrict__ edges, int64_t ne, int32_t *parent) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int u = edges[2 * e], v = edges[2 * e + 1];
    if (u == v) return;
    int ru = uf_find(parent, u), rv = uf_find(parent, v);
    while (ru != rv) {
        if (ru < rv) { int t = ru; ru = rv; rv = t; }          // ru = larger root, rv = smaller
        int old = atomicCAS(parent + ru, ru, rv);               // hang the larger root under the smaller
        if (old == ru) break;
        ru = uf_find(parent, old);
        rv = uf_find(parent, rv);
    }
}

__global__ void k_cc_flatten(int32_t *parent, int n, int32_t *size) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = i;
    while (true) { int p = parent[r]; if (p == r) break; r = p; }
    parent[i] = r;      // racing writers all write a value on the path to the same root
    atomicAdd(size + r, 1);
}

__global__ void k_cc_sizes(const int32_t *__restrict__ label, int n, int32_t *size) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = label[i];
    if (r != i) size[i] = size[r];
}

__global__ void k_cc_relabel(int32_t *label, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = label[i];
    while (true) { int p = label[r]; if (p == r) break; r = p; }
    label[i] = r;
}

extern "C" int st_connected_components(const int32_t *edges, int64_t n_edges, int64_t n, int32_t *label, int32_t *size, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    unsigned g = (unsigned)cdiv(n, 256);
    k_cc_init<<<g, 256, 0, s>>>(label, (int)n);
    ST_CHECK_LAUNCH();
    if (n_edges) {
        k_cc_hook<<<(unsigned)cdiv(n_edges, 256), 256, 0, s>>>(edges, n_edges, label);
        ST_CHECK_LAUNCH();
    }
    k_cc_relabel<<<g, 256, 0, s>>>(label, (int)n);   // two passes: after the first every entry points at a root
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cudaMemsetAsync(size, 0, n * 4, s));
    k_cc_flatten<<<g, 256, 0, s>>>(label, (int)n, size);
    ST_CHECK_LAUNCH();
    k_cc_sizes<<<g, 256, 0, s>>>(label, (int)n, size);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ CSR
__global__ void k_csr_degree(const int32_t *__restrict__ edges, int64_t ne, int32_t *deg, const int32_t *__restrict__ vmap) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int u = edges[2 * e], v = edges[2 * e + 1];
    if (vmap) { u = __ldg(vmap + u); v = __ldg(vmap + v); if (u < 0 || v < 0) return; }
    if (u == v) return;
    atomicAdd(deg + u, 1);
    atomicAdd(deg + v, 1);
}

__global__ void k_csr_fill(const int32_t *__restrict__ edges, const float *__restrict__ weights, int64_t ne,
                           const int32_t *__restrict__ row_ptr, int32_t *cursor, int32_t *__restrict__ col, float *__restrict__ w,
                           const int32_t *__restrict__ vmap) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int u = edges[2 * e], v = edges[2 * e + 1];
    if (vmap) { u = __ldg(vmap + u); v = __ldg(vmap + v); if (u < 0 || v < 0) return; }
    if (u == v) return;
    float ww = weights[e];
    int pu = row_ptr[u] + atomicAdd(cursor + u, 1);
    col[pu] = v; w[pu] = ww;
    int pv = row_ptr[v] + atomicAdd(cursor + v, 1);
    col[pv] = u; w[pv] = ww;
}

extern "C" size_t st_csr_workspace_bytes(int64_t n, int64_t n_edges) {
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (int *)nullptr, (int *)nullptr, (int)(n + 1));
    return align_up(scan) + 2 * align_up((n + 1) * 4) + 1024;
}

extern "C" int st_csr_build(const int32_t *edges, const float *weights, int64_t n_edges, const int32_t *vertex_map, int64_t n,
                            int32_t *row_ptr, int32_t *col, float *w, int64_t *n_arcs_host, void *workspace, size_t workspace_bytes,
                            void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    *n_arcs_host = 0;
    if (n == 0) return ST_OK;
    Carver cv(workspace, workspace_bytes);
    int32_t *deg = cv.take<int32_t>(n + 1);
    int32_t *cursor = cv.take<int32_t>(n + 1);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, deg, row_ptr, (int)(n + 1));
    void *scan_ws = cv.take<char>(scan_bytes);
    if (!cv.ok()) { set_error("st_csr_build: workspace too small"); return ST_ERR_WORKSPACE; }
    ST_CHECK_CUDA(cudaMemsetAsync(deg, 0, (n + 1) * 4, s));
    ST_CHECK_CUDA(cudaMemsetAsync(cursor, 0, (n + 1) * 4, s));
    if (n_edges) {
        k_csr_degree<<<(unsigned)cdiv(n_edges, 256), 256, 0, s>>>(edges, n_edges, deg, vertex_map);
        ST_CHECK_LAUNCH();
    }
    ST_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, deg, row_ptr, (int)(n + 1), s));
    if (n_edges) {
        k_csr_fill<<<(unsigned)cdiv(n_edges, 256), 256, 0, s>>>(edges, weights, n_edges, row_ptr, cursor, col, w, vertex_map);
        ST_CHECK_LAUNCH();
    }
    int32_t arcs = 0;
    ST_CHECK_CUDA(cudaMemcpyAsync(&arcs, row_ptr + n, 4, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    *n_arcs_host = arcs;
    return ST_OK;
}

// ------------------------------------------------------------------------------------ resident-grid barrier
// All CTAs are co-resident (cooperative launch).  `ctr` counts arrivals monotonically.
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned &phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ++phase;
        unsigned target = phase * gridDim.x;
        atomicAdd(ctr, 1u);
        while (*(volatile unsigned *)ctr < target) { __nanosleep(32); }
        __threadfence();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------ SSSP
// Pull-based asynchronous relaxation in a resident grid.  Every warp owns up to SSSP_G groups of 32
// consecutive vertices and polls one change COUNTER per vertex (bumped by whoever lowers a neighbour); a
// vertex whose counter moved re-evaluates
//     d[v] = min(d[v], min_u fl32(d[u] + w(u,v)))
// cooperatively (lanes stride over its arcs, L2-coherent loads), publishes an improvement (store, fence) and
// bumps the counters of exactly those neighbours the new value can still improve, judged from the distances
// it has just read (stale values only cause harmless extra wake-ups).  No flag clearing, no second read: a
// shortest-path chain advances one hop per poll + one round trip.  fl32(+) is monotone, so ANY schedule
// converges to the same least fixed point (== fp32 Dijkstra).  Grid barriers only every SSSP_PASSES polls;
// the kernel stops after a whole chunk in which no counter moved.
constexpr int SSSP_PASSES = 64;
constexpr int SSSP_G = 4;
constexpr int SSSP_NF_PASSES = 64;    // polls between barriers in the near-far kernel (tools/sssp_sweep.py: 64 polls, step 0.125 m best on the bench tree)
constexpr float ST_INF = __builtin_huge_valf();

struct SsspCtl {
    unsigned barrier;
    unsigned changed[3];
    unsigned chunks;
    unsigned evals, improvements, wakeups, lane_mode_groups;    // debug counters
    unsigned min_pend[3];                                       // smallest parked candidate per chunk (float bits)
};

__global__ void __launch_bounds__(256) k_sssp(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                              const float *__restrict__ w, int n, float *dist, int *dirty, SsspCtl *ctl, float delta, int npass, int adv) {
    // Distance-ordered ("near-far") discipline on top of the asynchronous relaxation: an improvement is only
    // accepted -- written and propagated -- while it lies below the current threshold T; larger candidates
    // stay parked in a register of the owning lane until T reaches them.  Vertices are therefore settled in
    // roughly increasing distance and each one is improved a few times instead of dozens (plain chaotic
    // relaxation improved every vertex ~35 times on the bench graph).  T only moves at the grid barriers:
    // to (smallest parked candidate) + delta -- as soon as anything is parked (adv != 0: the fastest part of the
    // wave has reached T; slower regions simply keep relaxing below it), or only once a whole chunk accepted
    // nothing (adv == 0, the round-1 schedule: one idle chunk per threshold step).
    unsigned phase = 0;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int ngroups = (n + 31) >> 5;
    int seen[SSSP_G], rb[SSSP_G], re[SSSP_G];
    float pend[SSSP_G];
#pragma unroll
    for (int k = 0; k < SSSP_G; ++k) {
        const int v = ((w0 + k * nwarps) << 5) + lane;
        seen[k] = 0;
        pend[k] = ST_INF;
        rb[k] = v < n ? __ldg(row_ptr + v) : 0;
        re[k] = v < n ? __ldg(row_ptr + v + 1) : 0;
    }
    float T = delta;
    for (unsigned chunk = 0;; ++chunk) {
        bool consumed = false;
        for (int pass = 0; pass < npass; ++pass) {
#pragma unroll
            for (int k = 0; k < SSSP_G; ++k) {
                const int g = w0 + k * nwarps;
                if (g >= ngroups) continue;
                const int v = (g << 5) + lane;
                const int cnt = v < n ? __ldcg(dirty + v) : seen[k];
                const bool woke = cnt != seen[k] || pend[k] <= T;
                seen[k] = cnt;
                unsigned mask = __ballot_sync(0xffffffffu, woke);
                if (!mask) continue;
                __threadfence();          // counter observed -> the distance that caused it is visible
                if (woke) pend[k] = ST_INF;
                while (mask) {            // woken vertices one after the other, each relaxed by the whole warp
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int vv = (g << 5) + l;
                    const int b = __shfl_sync(0xffffffffu, rb[k], l), e = __shfl_sync(0xffffffffu, re[k], l);
                    const float cur = __ldcg(dist + vv);
                    float best = cur;
                    // first 64 arcs stay in registers for the wake-up test
                    int u0 = -1, u1 = -1;
                    float c0 = ST_INF, c1 = ST_INF, d0 = 0.f, d1 = 0.f, ww0 = 0.f, ww1 = 0.f;
                    if (b + lane < e) { u0 = __ldg(col + b + lane); ww0 = __ldg(w + b + lane); d0 = __ldcg(dist + u0); c0 = __fadd_rn(d0, ww0); }
                    if (b + 32 + lane < e) { u1 = __ldg(col + b + 32 + lane); ww1 = __ldg(w + b + 32 + lane); d1 = __ldcg(dist + u1); c1 = __fadd_rn(d1, ww1); }
                    best = fminf(best, fminf(c0, c1));
                    for (int a = b + 64 + lane; a < e; a += 32)
                        best = fminf(best, __fadd_rn(__ldcg(dist + __ldg(col + a)), __ldg(w + a)));
                    for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
                    if (best < cur) {
                        if (best <= T) {
                            consumed = true;
                            if (lane == 0) { __stcg(dist + vv, best); __threadfence(); }
                            __syncwarp();
                            // u can only improve through vv if d[vv] + w < d[u]
                            if (u0 >= 0 && __fadd_rn(best, ww0) < d0) atomicAdd(dirty + u0, 1);
                            if (u1 >= 0 && __fadd_rn(best, ww1) < d1) atomicAdd(dirty + u1, 1);
                            for (int a = b + 64 + lane; a < e; a += 32) {
                                const int u = __ldg(col + a);
                                if (__fadd_rn(best, __ldg(w + a)) < __ldcg(dist + u)) atomicAdd(dirty + u, 1);
                            }
                        } else if (lane == l) {
                            pend[k] = best;       // parked until the threshold reaches it (or a neighbour wakes it again)
                        }
                    }
                }
            }
        }
        // smallest parked candidate of this thread -> grid minimum (non-negative floats order like their bits)
        float pmin = ST_INF;
#pragma unroll
        for (int k = 0; k < SSSP_G; ++k) pmin = fminf(pmin, pend[k]);
        for (int o = 16; o; o >>= 1) pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        if (lane == 0 && pmin < ST_INF) atomicMin(&ctl->min_pend[chunk % 3], __float_as_uint(pmin));
        if (__syncthreads_or(consumed) && threadIdx.x == 0) atomicOr(&ctl->changed[chunk % 3], 1u);
        if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->changed[(chunk + 1) % 3] = 0; ctl->min_pend[(chunk + 1) % 3] = 0x7F800000u; }
        grid_barrier(&ctl->barrier, phase);
        const unsigned any = *(volatile unsigned *)&ctl->changed[chunk % 3];
        const unsigned mp = *(volatile unsigned *)&ctl->min_pend[chunk % 3];
        if (!any) {
            if (mp == 0x7F800000u) {          // nothing accepted, nothing parked: fixed point reached
                if (blockIdx.x == 0 && threadIdx.x == 0) ctl->chunks = chunk + 1;
                break;
            }
            T = fmaxf(T, __uint_as_float(mp)) + delta;      // every thread derives the same new threshold
        } else if (adv && mp != 0x7F800000u) {
            T = fmaxf(T, __uint_as_float(mp) + delta);
        }
    }
}

// fallback for graphs with more than SSSP_G * 32 vertices per resident warp: same relaxation, flag-clearing
// variant in which a warp strides over arbitrarily many groups
__global__ void __launch_bounds__(256) k_sssp_big(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                  const float *__restrict__ w, int n, float *dist, int *dirty, SsspCtl *ctl) {
    unsigned phase = 0;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int ngroups = (n + 31) >> 5;
    for (unsigned chunk = 0;; ++chunk) {
        bool consumed = false;
        for (int pass = 0; pass < SSSP_PASSES; ++pass) {
            for (int g = w0; g < ngroups; g += nwarps) {
                const int v = (g << 5) + lane;
                int flag = v < n ? __ldcg(dirty + v) : 0;
                unsigned mask = __ballot_sync(0xffffffffu, flag != 0);
                if (!mask) continue;
                if (flag) atomicExch(dirty + v, 0);
                __threadfence();
                consumed = true;
                while (mask) {
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int vv = (g << 5) + l;
                    const int b = __ldg(row_ptr + vv), e = __ldg(row_ptr + vv + 1);
                    const float cur = __ldcg(dist + vv);
                    float best = cur;
                    for (int a = b + lane; a < e; a += 32)
                        best = fminf(best, __fadd_rn(__ldcg(dist + __ldg(col + a)), __ldg(w + a)));
                    for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
                    if (best < cur) {
                        if (lane == 0) { __stcg(dist + vv, best); __threadfence(); }
                        __syncwarp();
                        for (int a = b + lane; a < e; a += 32) {
                            const int u = __ldg(col + a);
                            if (__fadd_rn(best, __ldg(w + a)) < __ldcg(dist + u)) atomicExch(dirty + u, 1);
                        }
                    }
                }
            }
        }
        if (__syncthreads_or(consumed) && threadIdx.x == 0) atomicOr(&ctl->changed[chunk % 3], 1u);
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->changed[(chunk + 1) % 3] = 0;
        grid_barrier(&ctl->barrier, phase);
        unsigned any = *(volatile unsigned *)&ctl->changed[chunk % 3];
        if (!any) {
            if (blockIdx.x == 0 && threadIdx.x == 0) ctl->chunks = chunk + 1;
            break;
        }
    }
}

// Near-far relaxation (as k_sssp) for graphs too large for register-resident ownership: a warp strides over
// arbitrarily many groups of 32 vertices and keeps their poll state (last seen change counter, parked candidate)
// in global memory -- only the owning warp ever touches it.  The flag variant above improved every vertex dozens
// of times on a 1.3 M-vertex graph (77 ms); the distance-ordered schedule is what keeps the work near-linear.
__global__ void __launch_bounds__(256) k_sssp_nf_big(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                     const float *__restrict__ w, int n, float *dist, int *dirty, int *seen_g, float *pend_g,
                                                     SsspCtl *ctl, float delta, int npass, int adv) {
    unsigned phase = 0;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int ngroups = (n + 31) >> 5;
    float T = delta;
    for (unsigned chunk = 0;; ++chunk) {
        bool consumed = false;
        float pmin = ST_INF;
        for (int pass = 0; pass < npass; ++pass) {
            const bool last = pass == npass - 1;
            for (int g = w0; g < ngroups; g += nwarps) {
                const int v = (g << 5) + lane;
                const bool in = v < n;
                const int sn = in ? seen_g[v] : 0;
                const int cnt = in ? __ldcg(dirty + v) : sn;
                float pd = in ? pend_g[v] : ST_INF;
                const bool woke = cnt != sn || pd <= T;
                unsigned mask = __ballot_sync(0xffffffffu, woke);
                if (!mask) { if (last) pmin = fminf(pmin, pd); continue; }
                if (in && cnt != sn) seen_g[v] = cnt;
                __threadfence();          // counter observed -> the distance that caused it is visible
                if (woke) pd = ST_INF;
                while (mask) {            // woken vertices one after the other, each relaxed by the whole warp
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int vv = (g << 5) + l;
                    const int b = __ldg(row_ptr + vv), e = __ldg(row_ptr + vv + 1);
                    const float cur = __ldcg(dist + vv);
                    float best = cur;
                    int u0 = -1, u1 = -1;
                    float c0 = ST_INF, c1 = ST_INF, d0 = 0.f, d1 = 0.f, ww0 = 0.f, ww1 = 0.f;
                    if (b + lane < e) { u0 = __ldg(col + b + lane); ww0 = __ldg(w + b + lane); d0 = __ldcg(dist + u0); c0 = __fadd_rn(d0, ww0); }
                    if (b + 32 + lane < e) { u1 = __ldg(col + b + 32 + lane); ww1 = __ldg(w + b + 32 + lane); d1 = __ldcg(dist + u1); c1 = __fadd_rn(d1, ww1); }
                    best = fminf(best, fminf(c0, c1));
                    for (int a = b + 64 + lane; a < e; a += 32)
                        best = fminf(best, __fadd_rn(__ldcg(dist + __ldg(col + a)), __ldg(w + a)));
                    for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
                    if (best < cur) {
                        if (best <= T) {
                            consumed = true;
                            if (lane == 0) { __stcg(dist + vv, best); __threadfence(); }
                            __syncwarp();
                            if (u0 >= 0 && __fadd_rn(best, ww0) < d0) atomicAdd(dirty + u0, 1);
                            if (u1 >= 0 && __fadd_rn(best, ww1) < d1) atomicAdd(dirty + u1, 1);
                            for (int a = b + 64 + lane; a < e; a += 32) {
                                const int u = __ldg(col + a);
                                if (__fadd_rn(best, __ldg(w + a)) < __ldcg(dist + u)) atomicAdd(dirty + u, 1);
                            }
                        } else if (lane == l) {
                            pd = best;        // parked until the threshold reaches it (or a neighbour wakes it again)
                        }
                    }
                }
                if (in && woke) pend_g[v] = pd;
                if (last) pmin = fminf(pmin, pd);
            }
        }
        for (int o = 16; o; o >>= 1) pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        if (lane == 0 && pmin < ST_INF) atomicMin(&ctl->min_pend[chunk % 3], __float_as_uint(pmin));
        if (__syncthreads_or(consumed) && threadIdx.x == 0) atomicOr(&ctl->changed[chunk % 3], 1u);
        if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->changed[(chunk + 1) % 3] = 0; ctl->min_pend[(chunk + 1) % 3] = 0x7F800000u; }
        grid_barrier(&ctl->barrier, phase);
        const unsigned any = *(volatile unsigned *)&ctl->changed[chunk % 3];
        const unsigned mp = *(volatile unsigned *)&ctl->min_pend[chunk % 3];
        if (!any) {
            if (mp == 0x7F800000u) {
                if (blockIdx.x == 0 && threadIdx.x == 0) ctl->chunks = chunk + 1;
                break;
            }
            T = fmaxf(T, __uint_as_float(mp)) + delta;
        } else if (adv && mp != 0x7F800000u) {
            T = fmaxf(T, __uint_as_float(mp) + delta);
        }
    }
}


// ------------------------------------------------------------------------------------ SSSP with CTA-local propagation
// Same relaxation and the same distance-ordered acceptance as k_sssp, but every CTA keeps the distances and the wake-up
// counters of its own contiguous vertex range in SHARED memory.  With the vertices numbered in spatial (Z) order -- the
// caller builds the CSR that way and passes the map back to its own numbering as `orig_id` -- a CTA's range is a blob a few
// hops across: most hops of a shortest-path chain then stay inside one CTA and cost a shared-memory round trip instead of
// two L2 round trips (poll + relax), only the hops that cross a range boundary go through global memory as before.
//   * a vertex' distance lives in sdist (owner CTA, authoritative for its warps) AND in dist[] (written through on every
//     accepted improvement: what remote CTAs read);
//   * notifications to a neighbour in the same range bump its shared counter, to a neighbour elsewhere its global one;
//     the gpu-scope fence an accepted improvement needs before a REMOTE neighbour may be woken is only paid by vertices
//     that have remote neighbours to wake;
//   * per outer pass one poll of the global counters and up to `nlocal` polls of the shared ones.
// Any schedule converges to the same fp32 fixed point (see k_sssp).
template <int G>
__global__ void __launch_bounds__(1024, 1) k_sssp_local(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                        const float *__restrict__ w, int n, float *dist, int *dirty, SsspCtl *ctl, float delta,
                                                        int npass, int adv, int nlocal) {
    extern __shared__ __align__(16) unsigned char sssp_smem[];
    constexpr int VB = 1024 * G;                         // vertices per CTA
    volatile float *sdist = reinterpret_cast<volatile float *>(sssp_smem);
    int *sflag = reinterpret_cast<int *>(sssp_smem + (size_t)VB * sizeof(float));
    unsigned phase = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v0 = blockIdx.x * VB;
    for (int i = threadIdx.x; i < VB; i += 1024) {
        sdist[i] = v0 + i < n ? __ldcg(dist + v0 + i) : ST_INF;
        sflag[i] = 0;
    }
    __syncthreads();
    int seenL[G], seenG[G], rb[G], re[G];
    float pend[G];
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const int v = v0 + ((k * 32 + warp) << 5) + lane;
        seenL[k] = 0; seenG[k] = 0;
        pend[k] = ST_INF;
        rb[k] = v < n ? __ldg(row_ptr + v) : 0;
        re[k] = v < n ? __ldg(row_ptr + v + 1) : 0;
    }
    auto rd = [&](int u) -> float {                     // current distance of u: shared if it is ours, L2 otherwise
        const unsigned lu = (unsigned)(u - v0);
        return lu < (unsigned)VB ? sdist[lu] : __ldcg(dist + u);
    };
    float T = delta;
    for (unsigned chunk = 0;; ++chunk) {
        bool consumed = false;
        for (int pass = 0; pass < npass; ++pass) {
            unsigned wokeG = 0;                              // bit k: group k has a lane whose GLOBAL counter moved (per lane)
#pragma unroll
            for (int k = 0; k < G; ++k) {
                const int v = v0 + ((k * 32 + warp) << 5) + lane;
                const int cnt = v < n ? __ldcg(dirty + v) : seenG[k];
                if (cnt != seenG[k]) wokeG |= 1u << k;
                seenG[k] = cnt;
            }
            if (__any_sync(0xffffffffu, wokeG != 0)) __threadfence();      // counter observed -> the remote distance behind it is visible
            for (int q = 0; q < nlocal; ++q) {
                bool any = false;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    const int lv = ((k * 32 + warp) << 5) + lane;
                    const int v = v0 + lv;
                    const int cnt = *(volatile int *)(sflag + lv);
                    const bool woke = v < n && (cnt != seenL[k] || ((wokeG >> k) & 1u) || pend[k] <= T);
                    seenL[k] = cnt;
                    unsigned mask = __ballot_sync(0xffffffffu, woke);
                    if (!mask) continue;
                    any = true;
                    if (woke) pend[k] = ST_INF;
                    while (mask) {            // woken vertices one after the other, each relaxed by the whole warp
                        const int l = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int lvv = ((k * 32 + warp) << 5) + l;
                        const int vv = v0 + lvv;
                        const int b = __shfl_sync(0xffffffffu, rb[k], l), e = __shfl_sync(0xffffffffu, re[k], l);
                        const float cur = sdist[lvv];
                        float best = cur;
                        int u0 = -1, u1 = -1;
                        float c0 = ST_INF, c1 = ST_INF, d0 = 0.f, d1 = 0.f, ww0 = 0.f, ww1 = 0.f;
                        if (b + lane < e) { u0 = __ldg(col + b + lane); ww0 = __ldg(w + b + lane); d0 = rd(u0); c0 = __fadd_rn(d0, ww0); }
                        if (b + 32 + lane < e) { u1 = __ldg(col + b + 32 + lane); ww1 = __ldg(w + b + 32 + lane); d1 = rd(u1); c1 = __fadd_rn(d1, ww1); }
                        best = fminf(best, fminf(c0, c1));
                        for (int a = b + 64 + lane; a < e; a += 32) best = fminf(best, __fadd_rn(rd(__ldg(col + a)), __ldg(w + a)));
                        for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
                        if (best < cur) {
                            if (best <= T) {
                                consumed = true;
                                if (lane == 0) { sdist[lvv] = best; __stcg(dist + vv, best); }
                                __threadfence_block();
                                __syncwarp();
                                __threadfence_block();
                                // u can only improve through vv if d[vv] + w < d[u]
                                bool remote = false;
                                auto notify = [&](int u, float ww, float du) {
                                    if (u < 0 || !(__fadd_rn(best, ww) < du)) return;
                                    const unsigned lu = (unsigned)(u - v0);
                                    if (lu < (unsigned)VB) atomicAdd(sflag + lu, 1);
                                    else remote = true;
                                };
                                notify(u0, ww0, d0);
                                notify(u1, ww1, d1);
                                for (int a = b + 64 + lane; a < e; a += 32) { const int u = __ldg(col + a); notify(u, __ldg(w + a), rd(u)); }
                                if (__any_sync(0xffffffffu, remote)) {
                                    __threadfence();          // the improvement is visible device-wide before a remote owner is woken
                                    auto notify_remote = [&](int u, float ww, float du) {
                                        if (u < 0 || !(__fadd_rn(best, ww) < du)) return;
                                        if ((unsigned)(u - v0) >= (unsigned)VB) atomicAdd(dirty + u, 1);
                                    };
                                    notify_remote(u0, ww0, d0);
                                    notify_remote(u1, ww1, d1);
                                    for (int a = b + 64 + lane; a < e; a += 32) { const int u = __ldg(col + a); notify_remote(u, __ldg(w + a), __ldcg(dist + u)); }
                                }
                            } else if (lane == l) {
                                pend[k] = best;       // parked until the threshold reaches it (or a neighbour wakes it again)
                            }
                        }
                    }
                }
                wokeG = 0;
                if (!any) break;                    // nothing moved in this warp's groups: back to the global poll
            }
        }
        float pmin = ST_INF;
#pragma unroll
        for (int k = 0; k < G; ++k) pmin = fminf(pmin, pend[k]);
        for (int o = 16; o; o >>= 1) pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        if (lane == 0 && pmin < ST_INF) atomicMin(&ctl->min_pend[chunk % 3], __float_as_uint(pmin));
        if (__syncthreads_or(consumed) && threadIdx.x == 0) atomicOr(&ctl->changed[chunk % 3], 1u);
        if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->changed[(chunk + 1) % 3] = 0; ctl->min_pend[(chunk + 1) % 3] = 0x7F800000u; }
        grid_barrier(&ctl->barrier, phase);
        const unsigned anyc = *(volatile unsigned *)&ctl->changed[chunk % 3];
        const unsigned mp = *(volatile unsigned *)&ctl->min_pend[chunk % 3];
        if (!anyc) {
            if (mp == 0x7F800000u) {
                if (blockIdx.x == 0 && threadIdx.x == 0) ctl->chunks = chunk + 1;
                break;
            }
            T = fmaxf(T, __uint_as_float(mp)) + delta;
        } else if (adv && mp != 0x7F800000u) {
            T = fmaxf(T, __uint_as_float(mp) + delta);
        }
    }
}

// dist / pred in the caller's numbering when the graph is numbered in spatial order: out[orig[v]] = value of v, predecessor =
// the neighbour with the lowest ORIGINAL id among those that attain the distance (the tie rule is stated in the caller's ids)
__global__ void k_sssp_pred_orig(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const float *__restrict__ w,
                                 int n, const float *__restrict__ dist, const int32_t *__restrict__ orig, float *__restrict__ dist_out,
                                 int32_t *__restrict__ pred_out) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    float dv = dist[v];
    int best = INT_MAX;
    if (dv != ST_INF) {
        for (int a = row_ptr[v]; a < row_ptr[v + 1]; ++a) {
            int u = col[a];
            if (__fadd_rn(dist[u], w[a]) == dv) best = min(best, __ldg(orig + u));
        }
    }
    const int o = orig[v];
    pred_out[o] = best == INT_MAX ? -1 : best;
    dist_out[o] = dv == ST_INF ? FLT_MAX : dv;
}

__global__ void k_sssp_seed(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, int *dirty,
                            const int32_t *__restrict__ sources, int ns) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int s = sources[i];
    for (int a = row_ptr[s]; a < row_ptr[s + 1]; ++a) atomicAdd(dirty + col[a], 1);   // wake the sources' neighbours
}

__global__ void k_sssp_ctl_init(SsspCtl *ctl) { ctl->min_pend[0] = ctl->min_pend[1] = ctl->min_pend[2] = 0x7F800000u; }

__global__ void k_sssp_init(float *dist, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dist[i] = ST_INF;
}

__global__ void k_sssp_sources(float *dist, const int32_t *__restrict__ sources, int ns) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ns) dist[sources[i]] = 0.f;
}

__global__ void k_sssp_pred(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const float *__restrict__ w,
                            int n, const float *__restrict__ dist, int32_t *__restrict__ pred) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    float dv = dist[v];
    int best = INT_MAX;
    if (dv != ST_INF) {
        for (int a = row_ptr[v]; a < row_ptr[v + 1]; ++a) {
            int u = col[a];
            if (__fadd_rn(dist[u], w[a]) == dv) best = min(best, u);
        }
    }
    pred[v] = best == INT_MAX ? -1 : best;
}

__global__ void k_sssp_finish(float *dist, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && dist[i] == ST_INF) dist[i] = FLT_MAX;
}

__global__ void k_set_pred_sources(int32_t *pred, const int32_t *__restrict__ sources, int ns) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ns) pred[sources[i]] = -1;
}

static int coop_grid(const void *kernel, int threads, int device, int &blocks) {
    int per_sm = 0, sms = 0;
    ST_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    ST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    blocks = per_sm * sms;
    return ST_OK;
}

__global__ void k_set_pred_sources_orig(int32_t *pred, const int32_t *__restrict__ sources, const int32_t *__restrict__ orig, int ns) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ns) pred[orig[sources[i]]] = -1;
}

template <int G>
static int launch_sssp_local(int blocks_needed, int device, void **args, cudaStream_t s, bool &launched) {
    const size_t smem = (size_t)8 * 1024 * G;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute((const void *)k_sssp_local<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    int per_sm = 0, sms = 0;
    ST_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k_sssp_local<G>, 1024, smem));
    ST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    launched = false;
    if (per_sm * sms < blocks_needed) return ST_OK;
    ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_sssp_local<G>, dim3(blocks_needed), dim3(1024), args, smem, s));
    launched = true;
    return ST_OK;
}

extern "C" int st_sssp(const int32_t *row_ptr, const int32_t *col, const float *w, int64_t n, const int32_t *sources,
                       int32_t n_sources, float delta, float *dist_out, int32_t *pred, int32_t *sweeps_host, void *ctl_workspace,
                       const int32_t *orig_id, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (sweeps_host) *sweeps_host = 0;
    if (n == 0) return ST_OK;
    ST_REQUIRE(ctl_workspace != nullptr, "ctl_workspace (256 + 16n bytes) required");
    int device = 0;
    ST_CHECK_CUDA(cudaGetDevice(&device));
    SsspCtl *ctl = (SsspCtl *)ctl_workspace;
    // with orig_id the relaxation runs on an internal distance array in the graph's numbering; the results are scattered
    // to the caller's numbering at the end
    float *dist = orig_id ? (float *)((char *)ctl_workspace + 256 + 12 * (size_t)n) : dist_out;
    ST_CHECK_CUDA(cudaMemsetAsync(ctl, 0, sizeof(SsspCtl), s));
    k_sssp_ctl_init<<<1, 1, 0, s>>>(ctl);
    unsigned g = (unsigned)cdiv(n, 256);
    k_sssp_init<<<g, 256, 0, s>>>(dist, (int)n);
    ST_CHECK_LAUNCH();
    if (n_sources) {
        k_sssp_sources<<<(unsigned)cdiv(n_sources, 256), 256, 0, s>>>(dist, sources, n_sources);
        ST_CHECK_LAUNCH();
    }
    int *dirty = (int *)((char *)ctl_workspace + 256);
    ST_CHECK_CUDA(cudaMemsetAsync(dirty, 0, n * sizeof(int), s));
    if (n_sources) {
        k_sssp_seed<<<(unsigned)cdiv(n_sources, 256), 256, 0, s>>>(row_ptr, col, dirty, sources, n_sources);
        ST_CHECK_LAUNCH();
    }
    int blocks = 0;
    int rc = coop_grid((const void *)k_sssp, 256, device, blocks);
    if (rc) return rc;
    if (blocks > (int)g) blocks = (int)g;
    int nn = (int)n;
    if (!(delta > 0.f)) delta = 0.125f;
    int npass = SSSP_NF_PASSES;
    if (const char *e = getenv("ST_SSSP_PASSES")) { int v = atoi(e); if (v >= 1 && v <= 1024) npass = v; }
    int adv = 1;                           // threshold schedule (see k_sssp); any value gives the same result
    if (const char *e = getenv("ST_SSSP_ADVANCE")) adv = atoi(e) != 0;
    void *args[] = {(void *)&row_ptr, (void *)&col, (void *)&w, (void *)&nn, (void *)&dist, (void *)&dirty, (void *)&ctl, (void *)&delta, (void *)&npass, (void *)&adv};
    // near-far counter variant: poll state in registers when every resident warp can own its vertices there, in global
    // memory otherwise (ctl_workspace holds dirty[n], seen[n], pend[n]); ST_SSSP_FLAGS=1 forces the old flag variant
    const bool small = (int64_t)blocks * 8 * SSSP_G * 32 >= n && !getenv("ST_SSSP_FORCE_BIG");      // (env: tests exercise the large-graph kernel)
    // CTA-local propagation (k_sssp_local): the smallest range per CTA for which the whole graph is resident
    bool local_done = false;
    if (!getenv("ST_SSSP_NO_LOCAL") && !getenv("ST_SSSP_FORCE_BIG") && !getenv("ST_SSSP_FLAGS")) {
        int nlocal = 8;
        if (const char *e = getenv("ST_SSSP_NLOCAL")) { int v = atoi(e); if (v >= 1 && v <= 256) nlocal = v; }
        void *largs[] = {(void *)&row_ptr, (void *)&col, (void *)&w, (void *)&nn, (void *)&dist, (void *)&dirty, (void *)&ctl, (void *)&delta,
                         (void *)&npass, (void *)&adv, (void *)&nlocal};
        int gsel = 0;
        if (const char *e = getenv("ST_SSSP_LOCAL_G")) gsel = atoi(e);
        for (int G : {1, 2, 4, 8}) {
            if (local_done || (gsel && G != gsel)) continue;
            const int need = (int)cdiv(n, 1024 * (int64_t)G);
            int rc2 = ST_OK;
            if (G == 1) rc2 = launch_sssp_local<1>(need, device, largs, s, local_done);
            else if (G == 2) rc2 = launch_sssp_local<2>(need, device, largs, s, local_done);
            else if (G == 4) rc2 = launch_sssp_local<4>(need, device, largs, s, local_done);
            else rc2 = launch_sssp_local<8>(need, device, largs, s, local_done);
            if (rc2) return rc2;
        }
    }
    if (local_done) {
    } else if (small) {
        ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_sssp, dim3(blocks), dim3(256), args, 0, s));
    } else if (getenv("ST_SSSP_FLAGS")) {
        ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_sssp_big, dim3(blocks), dim3(256), args, 0, s));
    } else {
        int *seen_g = dirty + n;
        float *pend_g = (float *)(dirty + 2 * n);
        ST_CHECK_CUDA(cudaMemsetAsync(seen_g, 0, n * sizeof(int), s));
        k_sssp_init<<<g, 256, 0, s>>>(pend_g, (int)n);
        ST_CHECK_LAUNCH();
        int blocks2 = 0;
        rc = coop_grid((const void *)k_sssp_nf_big, 256, device, blocks2);
        if (rc) return rc;
        if (blocks2 > (int)g) blocks2 = (int)g;
        void *args2[] = {(void *)&row_ptr, (void *)&col, (void *)&w, (void *)&nn, (void *)&dist, (void *)&dirty, (void *)&seen_g, (void *)&pend_g,
                         (void *)&ctl, (void *)&delta, (void *)&npass, (void *)&adv};
        ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_sssp_nf_big, dim3(blocks2), dim3(256), args2, 0, s));
    }
    if (orig_id) {
        k_sssp_pred_orig<<<g, 256, 0, s>>>(row_ptr, col, w, (int)n, dist, orig_id, dist_out, pred);
        ST_CHECK_LAUNCH();
        if (n_sources) {
            k_set_pred_sources_orig<<<(unsigned)cdiv(n_sources, 256), 256, 0, s>>>(pred, sources, orig_id, n_sources);
            ST_CHECK_LAUNCH();
        }
    } else {
        k_sssp_pred<<<g, 256, 0, s>>>(row_ptr, col, w, (int)n, dist, pred);
        ST_CHECK_LAUNCH();
        k_sssp_finish<<<g, 256, 0, s>>>(dist, (int)n);
        ST_CHECK_LAUNCH();
        if (n_sources) {
            k_set_pred_sources<<<(unsigned)cdiv(n_sources, 256), 256, 0, s>>>(pred, sources, n_sources);
            ST_CHECK_LAUNCH();
        }
    }
    if (sweeps_host) {
        unsigned chunks = 0;
        ST_CHECK_CUDA(cudaMemcpyAsync(&chunks, &ctl->chunks, 4, cudaMemcpyDeviceToHost, s));
        ST_CHECK_CUDA(cudaStreamSynchronize(s));
        *sweeps_host = (int32_t)chunks;
    }
    return ST_OK;
}

// ------------------------------------------------------------------------------------ tree path lengths
// td[v] = td[pred[v]] + ||p_v - p_pred||, accumulated root -> leaf in fp32 (same order as an SSSP over
// the predecessor tree).  -1 marks "not ready"; resident threads poll their parent with L2-coherent loads.
__global__ void __launch_bounds__(256) k_tree_dist(const float *__restrict__ pts, const int32_t *__restrict__ pred, int n,
                                                   float *td, SsspCtl *ctl) {
    unsigned phase = 0;
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned chunk = 0;; ++chunk) {
        bool changed = false;
        for (int pass = 0; pass < SSSP_PASSES; ++pass) {
            for (int v = t0; v < n; v += stride) {
                if (__ldcg(td + v) >= 0.f) continue;
                int p = __ldg(pred + v);
                if (p < 0) continue;
                float tp = __ldcg(td + p);
                if (tp < 0.f) continue;
                float d = sqrtf(dist2_exact(pts[3 * (size_t)v], pts[3 * (size_t)v + 1], pts[3 * (size_t)v + 2],
                                            pts[3 * (size_t)p], pts[3 * (size_t)p + 1], pts[3 * (size_t)p + 2]));
                __stcg(td + v, __fadd_rn(tp, d));
                changed = true;
            }
        }
        if (__syncthreads_or(changed) && threadIdx.x == 0) atomicOr(&ctl->changed[chunk % 3], 1u);
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->changed[(chunk + 1) % 3] = 0;
        grid_barrier(&ctl->barrier, phase);
        unsigned any = *(volatile unsigned *)&ctl->changed[chunk % 3];
        if (!any) break;
    }
}

// Chain-walking variant: edge lengths are computed once (k_tree_edges), then every unfinished vertex walks up to
// TD_H predecessors looking for an ancestor whose length is final and adds the collected edge lengths back down in
// root -> leaf order -- bit for bit the sum the hop-by-hop kernel above forms, but the wave of finished vertices
// advances TD_H hops per L2 round trip instead of one (tree depth 368 on the bench graph: 0.7 ms -> ~0.05 ms).
// Finished values are final, so reading one "early" is harmless; passes are non-blocking (a thread that owns several
// vertices never waits inside one of them), so the resident grid always makes progress.
constexpr int TD_H = 32;

__global__ void k_tree_edges(const float *__restrict__ pts, const int32_t *__restrict__ pred, const uint8_t *__restrict__ is_root, int n,
                             int2 *__restrict__ pe, float *__restrict__ td) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int p = pred[v];
    float e = 0.f;
    if (p >= 0)
        e = sqrtf(dist2_exact(pts[3 * (size_t)v], pts[3 * (size_t)v + 1], pts[3 * (size_t)v + 2],
                              pts[3 * (size_t)p], pts[3 * (size_t)p + 1], pts[3 * (size_t)p + 2]));
    pe[v] = make_int2(p, __float_as_int(e));
    td[v] = is_root[v] ? 0.f : -1.f;
}

__global__ void __launch_bounds__(256) k_tree_dist_walk(const int2 *__restrict__ pe, int n, float *td) {
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    bool pending = true;
    while (pending) {
        pending = false;
        for (int v = t0; v < n; v += stride) {
            if (__ldcg(td + v) >= 0.f) continue;
            float e[TD_H];
            int cnt = 0, cur = v;
            float t = -1.f;
            bool dead = false;
#pragma unroll
            for (int h = 0; h < TD_H; ++h) {
                if (t < 0.f && !dead) {
                    const int2 x = __ldg(pe + cur);
                    if (x.x < 0) {
                        dead = true;                       // a non-root without predecessor: nothing below it is reachable
                    } else {
                        e[h] = __int_as_float(x.y);
                        cnt = h + 1;
                        cur = x.x;
                        t = __ldcg(td + cur);
                    }
                }
            }
            if (dead) { __stcg(td + v, FLT_MAX); continue; }
            if (t >= 0.f) {
                float s = t;
                if (t != FLT_MAX) {
#pragma unroll
                    for (int h = TD_H - 1; h >= 0; --h)
                        if (h < cnt) s = __fadd_rn(s, e[h]);
                }
                __stcg(td + v, s);
            } else {
                pending = true;
            }
        }
    }
}

__global__ void k_tree_dist_init(const int32_t *__restrict__ pred, const uint8_t *__restrict__ is_root, int n, float *td) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    td[i] = is_root[i] ? 0.f : -1.f;
}

__global__ void k_tree_dist_finish(int n, float *td) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && td[i] < 0.f) td[i] = FLT_MAX;   // unreachable from any root
}

extern "C" int st_tree_distances(const float *points, const int32_t *pred, const uint8_t *is_root, int64_t n, float *tree_dist,
                                 void *ctl_workspace, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    ST_REQUIRE(ctl_workspace != nullptr, "ctl_workspace (256 + 8n bytes) required");
    int device = 0;
    ST_CHECK_CUDA(cudaGetDevice(&device));
    SsspCtl *ctl = (SsspCtl *)ctl_workspace;
    ST_CHECK_CUDA(cudaMemsetAsync(ctl, 0, sizeof(SsspCtl), s));
    unsigned g = (unsigned)cdiv(n, 256);
    if (!getenv("ST_TREE_DIST_HOPWISE")) {
        int2 *pe = (int2 *)((char *)ctl_workspace + 256);
        k_tree_edges<<<g, 256, 0, s>>>(points, pred, is_root, (int)n, pe, tree_dist);
        ST_CHECK_LAUNCH();
        int blocks = 0;
        int rc = coop_grid((const void *)k_tree_dist_walk, 256, device, blocks);
        if (rc) return rc;
        if (blocks > (int)g) blocks = (int)g;
        int nn = (int)n;
        const int2 *pe_c = pe;
        void *args[] = {(void *)&pe_c, (void *)&nn, (void *)&tree_dist};
        ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_tree_dist_walk, dim3(blocks), dim3(256), args, 0, s));
        return ST_OK;
    }
    k_tree_dist_init<<<g, 256, 0, s>>>(pred, is_root, (int)n, tree_dist);
    ST_CHECK_LAUNCH();
    int blocks = 0;
    int rc = coop_grid((const void *)k_tree_dist, 256, device, blocks);
    if (rc) return rc;
    if (blocks > (int)g) blocks = (int)g;
    int nn = (int)n;
    void *args[] = {(void *)&points, (void *)&pred, (void *)&nn, (void *)&tree_dist, (void *)&ctl};
    ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_tree_dist, dim3(blocks), dim3(256), args, 0, s));
    k_tree_dist_finish<<<g, 256, 0, s>>>((int)n, tree_dist);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ point -> tube (repair)
// queries.py:89-133: t = clip(ap.ab / ab.ab, 0, 1) ; proj = a + t ab ; dist = ||proj - p|| ;
// r = (1-t) r1 + t r2 ; argmin |dist - r| (first minimum)
__global__ void __launch_bounds__(128) k_points_to_tubes(const float *__restrict__ pts, int nq, const float *__restrict__ a,
                                                         const float *__restrict__ b, const float *__restrict__ r1,
                                                         const float *__restrict__ r2, const int32_t *__restrict__ off,
                                                         float *__restrict__ out_vec, int32_t *__restrict__ out_idx,
                                                         float *__restrict__ out_r) {
    // one warp per query; lanes stride over the query's tubes, then an argmin reduction (first minimum)
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float px = pts[3 * q], py = pts[3 * q + 1], pz = pts[3 * q + 2];
    float best = ST_INF, bvx = 0, bvy = 0, bvz = 0, br = 0;
    int bidx = INT_MAX;
    const int m0 = off[q], m1 = off[q + 1];
    for (int m = m0 + lane; m < m1; m += 32) {
        float ax = a[3 * m], ay = a[3 * m + 1], az = a[3 * m + 2];
        float abx = b[3 * m] - ax, aby = b[3 * m + 1] - ay, abz = b[3 * m + 2] - az;
        float apx = px - ax, apy = py - ay, apz = pz - az;
        float num = apx * abx + apy * aby + apz * abz;
        float den = abx * abx + aby * aby + abz * abz;
        float t = fminf(fmaxf(num / den, 0.f), 1.f);
        float qx = ax + t * abx, qy = ay + t * aby, qz = az + t * abz;
        float dx = qx - px, dy = qy - py, dz = qz - pz;
        float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        float r = (1.f - t) * r1[m] + t * r2[m];
        float score = fabsf(dist - r);
        if (score < best) { best = score; bidx = m - m0; bvx = dx; bvy = dy; bvz = dz; br = r; }
    }
    for (int o = 16; o; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        float ox = __shfl_xor_sync(0xffffffffu, bvx, o), oy = __shfl_xor_sync(0xffffffffu, bvy, o);
        float oz = __shfl_xor_sync(0xffffffffu, bvz, o), orr = __shfl_xor_sync(0xffffffffu, br, o);
        if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bvx = ox; bvy = oy; bvz = oz; br = orr; }
    }
    if (lane == 0) {
        out_vec[3 * q] = bvx; out_vec[3 * q + 1] = bvy; out_vec[3 * q + 2] = bvz;
        out_idx[q] = bidx == INT_MAX ? -1 : bidx;
        out_r[q] = br;
    }
}

extern "C" int st_points_to_tubes(const float *pts, int64_t n_q, const float *a, const float *b, const float *r1, const float *r2,
                                  const int32_t *tube_off, float *out_vec, int32_t *out_idx, float *out_r, void *stream) {
    if (n_q == 0) return ST_OK;
    k_points_to_tubes<<<(unsigned)cdiv(n_q * 32, 128), 128, 0, (cudaStream_t)stream>>>(pts, (int)n_q, a, b, r1, r2, tube_off, out_vec, out_idx, out_r);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ repair, all levels in one launch
// TreeSkeleton.repair (smart_tree/data_types/tree.py:73-92) on the shared node array nodes[R,4] (xyz, radius;
// branch b owns rows row[b]+1 .. row[b]+len[b], row[b] is its spare row).  Branches are given sorted by tree
// depth (level_off); a branch's connection point is the nearest-tube projection of its first node onto its
// parent's CURRENT polyline (which includes the parent's own connection point iff parent_repaired).
// One CTA walks the levels with a block barrier in between; one warp per branch.
__global__ void __launch_bounds__(1024) k_repair(float *nodes, const int32_t *__restrict__ row, const int32_t *__restrict__ len,
                                                 const int32_t *__restrict__ parent, const uint8_t *__restrict__ parent_repaired,
                                                 const int32_t *__restrict__ level_off, int n_levels) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int lv = 0; lv < n_levels; ++lv) {
        for (int b = level_off[lv] + warp; b < level_off[lv + 1]; b += nwarp) {
            const int pb = parent[b];
            const int r0 = row[b] + 1;
            const float px = nodes[4 * r0], py = nodes[4 * r0 + 1], pz = nodes[4 * r0 + 2];
            const int t0 = parent_repaired[b] ? row[pb] : row[pb] + 1;     // first node row of the parent polyline
            const int t1 = row[pb] + len[pb];                              // last node row
            float best = ST_INF, bx = 0, by = 0, bz = 0;
            int bidx = INT_MAX;
            for (int m = t0 + lane; m < t1; m += 32) {
                float ax = nodes[4 * m], ay = nodes[4 * m + 1], az = nodes[4 * m + 2], r1 = nodes[4 * m + 3];
                float cx = nodes[4 * m + 4], cy = nodes[4 * m + 5], cz = nodes[4 * m + 6], r2 = nodes[4 * m + 7];
                float abx = cx - ax, aby = cy - ay, abz = cz - az;
                float apx = px - ax, apy = py - ay, apz = pz - az;
                float t = fminf(fmaxf((apx * abx + apy * aby + apz * abz) / (abx * abx + aby * aby + abz * abz), 0.f), 1.f);
                float dx = ax + t * abx - px, dy = ay + t * aby - py, dz = az + t * abz - pz;
                float score = fabsf(sqrtf(dx * dx + dy * dy + dz * dz) - ((1.f - t) * r1 + t * r2));
                if (score < best) { best = score; bidx = m; bx = dx; by = dy; bz = dz; }
            }
            for (int o = 16; o; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                float ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
                if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bx = ox; by = oy; bz = oz; }
            }
            if (lane == 0) {
                const int rs = row[b];
                nodes[4 * rs] = px + bx; nodes[4 * rs + 1] = py + by; nodes[4 * rs + 2] = pz + bz;
                nodes[4 * rs + 3] = nodes[4 * r0 + 3];                      // radius of the first node (tree.py:92)
            }
        }
        __syncthreads();
    }
}

extern "C" int st_repair_branches(float *nodes, const int32_t *row, const int32_t *len, const int32_t *parent,
                                  const uint8_t *parent_repaired, const int32_t *level_off, int32_t n_levels, void *stream) {
    if (n_levels <= 0) return ST_OK;
    k_repair<<<1, 1024, 0, (cudaStream_t)stream>>>(nodes, row, len, parent, parent_repaired, level_off, n_levels);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// debug only: [chunks, evaluations, improvements, wakeups(unused), lane-mode groups] of the last st_sssp on this workspace
extern "C" int st_debug_sssp_stats(const void *ctl_workspace, unsigned *out_host) {
    const SsspCtl *c = (const SsspCtl *)ctl_workspace;
    ST_CHECK_CUDA(cudaMemcpy(out_host, &c->chunks, sizeof(unsigned) * 5, cudaMemcpyDeviceToHost));
    return ST_OK;
}
