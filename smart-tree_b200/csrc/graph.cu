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

// Hooking pass over the edges e = first, first + step, ...  With skip_equal the pass follows a compression (every vertex
// points at its root of the moment): two vertices with the same parent are in the same set already, which is the case for
// almost every edge once a sample of the edges has been hooked -- two loads instead of two root searches.
__global__ void k_cc_hook(const int32_t *__restrict__ edges, int64_t ne, int32_t *parent, int step, int skip_equal) {
    int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * step;
    if (e >= ne) return;
    int u = edges[2 * e], v = edges[2 * e + 1];
    if (u == v) return;
    if (skip_equal && __ldcg(parent + u) == __ldcg(parent + v)) return;
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
    // one atomic per distinct root and warp (a tree cloud is ONE component: 245 k increments of the same word otherwise)
    const unsigned peers = __match_any_sync(__activemask(), r);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(size + r, __popc(peers));
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
        // Sampled first (every 8th edge), compressed, then all edges with the same-parent shortcut (the idea of Afforest,
        // Sutton et al. 2018): a tree cloud is one giant component, after the sample nearly every vertex already points at
        // its root.  Bench graph (245 k vertices, 3.9 M edges): 512 us for one plain pass -> see profiles/README.md.
        int sample = getenv("ST_CC_NO_SAMPLE") ? 0 : 8;
        if (const char *e = getenv("ST_CC_SAMPLE")) { int v = atoi(e); if (v >= 2 && v <= 1024) sample = v; }
        if (sample && n_edges >= 4096) {
            k_cc_hook<<<(unsigned)cdiv(cdiv(n_edges, sample), 256), 256, 0, s>>>(edges, n_edges, label, sample, 0);
            ST_CHECK_LAUNCH();
            k_cc_relabel<<<g, 256, 0, s>>>(label, (int)n);
            ST_CHECK_LAUNCH();
            k_cc_hook<<<(unsigned)cdiv(n_edges, 256), 256, 0, s>>>(edges, n_edges, label, 1, 1);
        } else {
            k_cc_hook<<<(unsigned)cdiv(n_edges, 256), 256, 0, s>>>(edges, n_edges, label, 1, 0);
        }
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
                                              const float *__restrict__ w, int n, float *dist, int *dirty, SsspCtl *ctl, float delta, int npass, int adv, int pairs) {
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
                    if (pairs && mask) {
                        // two woken vertices at a time, one per half warp (a kNN vertex has ~16 arcs: half of the lanes of a
                        // whole-warp relaxation idle, and the vertices of a group wake in bunches as the wave front passes)
                        const int l1 = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int hl = lane & 15, lm = (lane >> 4) ? l1 : l;
                        const int vv = (g << 5) + lm;
                        const int b = __shfl_sync(0xffffffffu, rb[k], lm), e = __shfl_sync(0xffffffffu, re[k], lm);
                        const float cur = __ldcg(dist + vv);
                        float best = cur;
                        int u0 = -1, u1 = -1;
                        float c0 = ST_INF, c1 = ST_INF, d0 = 0.f, d1 = 0.f, ww0 = 0.f, ww1 = 0.f;
                        if (b + hl < e) { u0 = __ldg(col + b + hl); ww0 = __ldg(w + b + hl); d0 = __ldcg(dist + u0); c0 = __fadd_rn(d0, ww0); }
                        if (b + 16 + hl < e) { u1 = __ldg(col + b + 16 + hl); ww1 = __ldg(w + b + 16 + hl); d1 = __ldcg(dist + u1); c1 = __fadd_rn(d1, ww1); }
                        best = fminf(best, fminf(c0, c1));
                        for (int a = b + 32 + hl; a < e; a += 16)
                            best = fminf(best, __fadd_rn(__ldcg(dist + __ldg(col + a)), __ldg(w + a)));
                        for (int o = 8; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));      // (stays inside the half)
                        const bool acc = best < cur && best <= T;
                        if (acc && hl == 0) __stcg(dist + vv, best);
                        if (__any_sync(0xffffffffu, acc)) {
                            consumed = true;
                            __threadfence();
                            __syncwarp();
                            if (acc) {
                                if (u0 >= 0 && __fadd_rn(best, ww0) < d0) atomicAdd(dirty + u0, 1);
                                if (u1 >= 0 && __fadd_rn(best, ww1) < d1) atomicAdd(dirty + u1, 1);
                                for (int a = b + 32 + hl; a < e; a += 16) {
                                    const int u = __ldg(col + a);
                                    if (__fadd_rn(best, __ldg(w + a)) < __ldcg(dist + u)) atomicAdd(dirty + u, 1);
                                }
                            }
                        }
                        const float b0 = __shfl_sync(0xffffffffu, best, 0), b1 = __shfl_sync(0xffffffffu, best, 16);
                        const float q0 = __shfl_sync(0xffffffffu, cur, 0), q1 = __shfl_sync(0xffffffffu, cur, 16);
                        if (lane == l && b0 < q0 && b0 > T) pend[k] = b0;        // parked until the threshold reaches it
                        if (lane == l1 && b1 < q1 && b1 > T) pend[k] = b1;
                        continue;
                    }
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
// Same relaxation and the same distance-ordered ("near-far") acceptance as k_sssp, organised around WHERE a shortest-path
// chain spends its time: every hop is a dependent round trip, and through L2 (poll a counter, relax, fence, notify) it
// costs microseconds.  Here every CTA keeps the distances and wake-up counters of its own contiguous vertex range in SHARED
// memory.  With the vertices numbered in spatial (Z) order -- the caller builds the CSR that way and passes the map back
// to its own numbering as `orig_id` -- a range is a stretch of branch some tens of hops long (bench tree, 2048 vertices per
// CTA: 7 % of the arcs and 39 of the 363 hops of the deepest path leave their range; in the caller's numbering 90 % and
// 343); a hop inside it costs a shared-memory round trip, only the hops across a range boundary go through L2.
//   * one LANE per vertex (a group of 32 consecutive vertices is one cross-section of a branch: the wave reaches its
//     vertices together, so they are relaxed together, every lane walking its own arcs);
//   * a vertex' distance lives in sdist (owner CTA) and is written through to dist[] (what remote CTAs read);
//   * arcs to vertices of the own range are evaluated whenever the vertex is woken; arcs to REMOTE vertices only when the
//     vertex' GLOBAL counter moved -- a remote neighbour that improved bumps exactly that counter (judged against a
//     distance of ours it has read, which can only be staler = larger than the truth, so no wake-up is ever missed);
//   * improvements are accepted while they lie below the threshold T of the current EPOCH, larger candidates stay parked
//     in a register of the owning lane.  Without this order the relaxation is chaotic: every tiny improvement near the
//     root re-sums the whole subtree behind it (measured on the bench tree: thousands of correction waves, 33 ms);
//   * the warps of a CTA run free -- no block barrier while anything moves in the CTA (s_nactive counts the warps that
//     found work in their last poll);
//   * no grid barrier, and NO gpu-scope fence on the hot path (on sm_100 __threadfence() = MEMBAR.SC + CCTL.IVALL: it
//     empties the SM's L1, where the arcs of the range live).  Every CTA has a MAILBOX word: notification count << 2 |
//     epoch parity << 1 | idle bit.  A remote notification = relaxed bumps of the vertices' counters, then ONE
//     release-add on the owner's mailbox (MEMBAR.ALL.GPU + atom: distance and counters are performed first).  A CTA in
//     which nothing moves below T publishes its smallest parked candidate (atomicMin), ORs the idle bit into its own
//     mailbox and compares the count it gets back with the count it has consumed: atomics on one word are totally ordered,
//     so either the owner sees the notification or the notifier sees the idle bit -- then it clears the bit and takes the
//     CTA out of the epoch's idle count on its behalf.  The count reaches the number of CTAs only when no CTA is active
//     and no notification is pending; the CTA whose increment completes it opens the next epoch with T = (smallest parked
//     candidate) + delta -- or ends the kernel if nothing is parked: the fixed point.
// Any schedule converges to the same fp32 fixed point (see k_sssp); T and delta only decide how much work it takes.
struct SsspBlobCtl {
    unsigned long long epoch_T;  // (epoch << 32) | float bits of T: one word, so a reader never pairs a new epoch with an old T
    int idle_count[2];           // by epoch parity
    unsigned gmin[2];            // smallest parked candidate published in the epoch (float bits), by epoch parity
    int rounds;                  // rounds of CTA 0 (diagnostics)
    int pad0;
    unsigned long long stats[4]; // relaxations, accepted improvements, remote notifications, epochs
    unsigned long long phase[6]; // thread 0 of every CTA, summed (cycles): local polls, global step, barriers, idle spin, whole kernel; [5] = busiest CTA
    int pad[36];
    int mbox[1];                 // [gridDim.x] mailboxes
};
constexpr int ST_EPOCH_DONE = 0x7FFFFFFF;

__device__ __forceinline__ int atom_add_release(int *p, int v) {
    int old;
    asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__global__ void k_sssp_blob_init(SsspBlobCtl *ctl, int nblocks, float T0) {
    if (threadIdx.x == 0) {
        ctl->epoch_T = (unsigned long long)__float_as_uint(T0);
        ctl->idle_count[0] = ctl->idle_count[1] = 0;
        ctl->gmin[0] = ctl->gmin[1] = 0x7F800000u;
        ctl->rounds = 0;
        for (int i = 0; i < 4; ++i) ctl->stats[i] = 0;
        for (int i = 0; i < 6; ++i) ctl->phase[i] = 0;
    }
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) ctl->mbox[i] = 0;
}

// debug (ST_SSSP_BLOB_FLAGS & 4): %globaltimer >> 3 of the last accepted improvement of every vertex (graph numbering)
__device__ unsigned g_sssp_tfinal[1 << 20];
__device__ unsigned long long g_sssp_epoch_log[1024][4];    // per epoch: time of the advance, advancing CTA, its rounds, time it woke up in this epoch
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

constexpr int SSSP_OW = 960;       // threads of a k_sssp_blob CTA that own vertices (30 warps); the last two warps publish

template <int G>
__global__ void __launch_bounds__(1024, 1) k_sssp_blob(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                       const float *__restrict__ w, int n, float *dist, int *dirty, SsspBlobCtl *ctl,
                                                       float delta, int nlocal, int flags) {
    constexpr int VB = SSSP_OW * G;                      // vertices per CTA
    constexpr int NW = (VB + 31) / 32;                   // words of the publish bitmap
    extern __shared__ __align__(16) unsigned char sssp_smem[];
    float *sdist = reinterpret_cast<float *>(sssp_smem);
    int *sflag = reinterpret_cast<int *>(sssp_smem + (size_t)VB * sizeof(float));
    unsigned *rbits = reinterpret_cast<unsigned *>(sssp_smem + (size_t)VB * 8);      // bit lv: improved, remote neighbours not yet told
    __shared__ int s_nactive;                            // owner warps that found work in their last poll
    __shared__ unsigned s_minpend[2];                    // by round parity: the slot of the next round is reset between this round's barriers
    __shared__ int s_state;
    __shared__ int s_mbox;                               // notification count of our mailbox as the publisher warp last saw it
    const int tid = threadIdx.x, lane = tid & 31;
    const bool owner = tid < SSSP_OW;
    const int v0 = blockIdx.x * VB;
    const int NB = (int)gridDim.x;
    int *const mybox = ctl->mbox + blockIdx.x;
    for (int i = tid; i < VB; i += 1024) {
        sdist[i] = v0 + i < n ? __ldcg(dist + v0 + i) : ST_INF;
        sflag[i] = 0;
    }
    for (int i = tid; i < NW; i += 1024) rbits[i] = 0;
    if (tid == 0) { s_nactive = 0; s_minpend[0] = s_minpend[1] = 0x7F800000u; s_state = 0; s_mbox = 0; }
    int seenL[G], seenG[G], rb[G], re[G];
    float pend[G];
    unsigned hasrem = 0;                                 // bit k: vertex k has arcs that leave the range
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const int v = v0 + k * SSSP_OW + tid;
        const bool in = owner && v < n;
        seenL[k] = 0; seenG[k] = 0;
        pend[k] = ST_INF;
        rb[k] = in ? __ldg(row_ptr + v) : 0;
        re[k] = in ? __ldg(row_ptr + v + 1) : 0;
        for (int a = rb[k]; a < re[k]; ++a)
            if ((unsigned)(__ldg(col + a) - v0) >= (unsigned)VB) hasrem |= 1u << k;
    }
    __syncthreads();
    volatile float *vdist = sdist;
    volatile int *vflag = sflag;
    volatile int *vnact = &s_nactive;
    volatile unsigned *vbits = rbits;
    volatile int *vmbox = &s_mbox;
    float T = 0.f;
    int epoch = 0;
    int epoch_prev = -1;
    unsigned long long t_epoch_wake = 0;
    int mb_seen = -1;                                    // notification count of our mailbox that has been consumed (-1: look at the counters once)
    unsigned n_relax = 0, n_acc = 0, n_remote = 0;
    long long ph[4] = {0, 0, 0, 0};
    const long long t_begin = clock64();

    // One relaxation of local vertex lv (arcs [b, e)); gw: its global counter moved, remote arcs are evaluated too.
    // pd (in/out): the vertex' parked candidate -- a realisable path length, possibly through a remote neighbour, so it
    // takes part in the minimum even when only the local arcs are walked.
    auto relax = [&](int lv, int b, int e, bool gw, bool rem, float &pd) -> bool {
        ++n_relax;
        const float cur = vdist[lv];
        float best = fminf(cur, pd);
        pd = ST_INF;
        for (int a = b; a < e; a += 4) {
            int u[4]; float ww[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = a + i < e;
                u[i] = in ? __ldg(col + a + i) : -1;
                ww[i] = in ? __ldg(w + a + i) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (u[i] < 0) continue;
                const unsigned lu = (unsigned)(u[i] - v0);
                float du = ST_INF;
                if (lu < (unsigned)VB) du = vdist[lu];
                else if (gw) du = __ldcg(dist + u[i]);
                best = fminf(best, __fadd_rn(du, ww[i]));
            }
        }
        if (!(best < cur)) return false;
        if (best > T) { pd = best; return false; }       // parked until the threshold reaches it
        ++n_acc;
        vdist[lv] = best;
        __stcg(dist + v0 + lv, best);
        if ((flags & 4) && v0 + lv < (1 << 20)) g_sssp_tfinal[v0 + lv] = (unsigned)(global_timer() >> 3);
        __threadfence_block();                           // the new distance is visible in the CTA before any counter of it moves
        // local neighbours that can still improve through this vertex (d + w < d[u]) are woken right away; the remote ones
        // are notified by the publisher warps (bit in rbits): the wave inside the range never waits for L2
        for (int a = b; a < e; a += 4) {
            int u[4]; float c[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = a + i < e;
                u[i] = in ? __ldg(col + a + i) : -1;
                c[i] = in ? __fadd_rn(best, __ldg(w + a + i)) : ST_INF;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned lu = (unsigned)(u[i] - v0);
                if (u[i] >= 0 && lu < (unsigned)VB && c[i] < vdist[lu]) atomicAdd(sflag + lu, 1);
            }
        }
        if (rem) atomicOr(rbits + (lv >> 5), 1u << (lv & 31));
        return true;
    };

    // Remote notifications of local vertex lv at its CURRENT distance (several improvements since the last call are one
    // notification): counters of the remote vertices it can still improve (relaxed), then one release-add per owner mailbox.
    auto publish = [&](int lv) {
        const int b = __ldg(row_ptr + v0 + lv), e = __ldg(row_ptr + v0 + lv + 1);
        const float d = vdist[lv];
        unsigned long long rmask = 0ull;
        bool rover = false;
        for (int a = b; a < e; a += 4) {
            int u[4]; float c[4], du[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = a + i < e;
                u[i] = in ? __ldg(col + a + i) : -1;
                c[i] = in ? __fadd_rn(d, __ldg(w + a + i)) : ST_INF;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {               // the remote distances of the four arcs are in flight together
                const unsigned lu = (unsigned)(u[i] - v0);
                du[i] = (u[i] >= 0 && lu >= (unsigned)VB) ? __ldcg(dist + u[i]) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned lu = (unsigned)(u[i] - v0);
                if (u[i] >= 0 && lu >= (unsigned)VB && c[i] < du[i]) {
                    atomicAdd(dirty + u[i], 1);
                    if (a + i - b < 64) rmask |= 1ull << (a + i - b); else rover = true;
                }
            }
        }
        if (!(rmask || rover)) return;
        ++n_remote;
        int lastX = -1;
        for (int a = b; a < e; ++a) {
            const int i = a - b;
            if (i < 64 && !((rmask >> i) & 1ull)) continue;
            const int u = __ldg(col + a);
            if ((unsigned)(u - v0) < (unsigned)VB) continue;
            const int X = u / VB;
            if (X == lastX) continue;                    // (a repeated bump of the same mailbox would only cost the owner another look)
            lastX = X;
            const int old = atom_add_release(ctl->mbox + X, 4);
            if (old & 1) {                               // the owner is idle: wake it and take it out of the idle count of its epoch
                const int old2 = atomicAnd(ctl->mbox + X, ~3);
                if (old2 & 1) atomicSub(&ctl->idle_count[(old2 >> 1) & 1], 1);
            }
        }
    };

    for (int round = 0;; ++round) {
        // epoch / threshold (they only change while every CTA is idle)
        {
            const unsigned long long et = *(volatile unsigned long long *)&ctl->epoch_T;
            epoch = (int)(et >> 32);
            T = __uint_as_float((unsigned)et);
        }
        bool did = false;
        long long t0 = clock64();
        if (epoch != epoch_prev) { epoch_prev = epoch; t_epoch_wake = global_timer(); }
        // One polling loop per round, left only when the whole CTA is quiet: no warp has found work for two polls, the
        // publish bitmap is empty and the mailbox count every warp has acted on is the current one.  Warps that have left
        // wait at the barrier below -- deaf to their wake-up counters -- so leaving must be rare.
        if (owner) {
            bool awake = false;                          // this warp is counted in s_nactive
            int quiet = 0;
            for (int q = 0; q < nlocal; ++q) {
                bool woke[G];
                bool anyw_lane = false;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    const int lv = k * SSSP_OW + tid;
                    const int cnt = vflag[lv];
                    woke[k] = v0 + lv < n && (cnt != seenL[k] || pend[k] <= T);
                    seenL[k] = cnt;
                    anyw_lane |= woke[k];
                }
                if (__any_sync(0xffffffffu, anyw_lane)) {
                    if (!awake) { awake = true; if (lane == 0) atomicAdd(&s_nactive, 1); }
#pragma unroll
                    for (int k = 0; k < G; ++k)
                        if (woke[k] && relax(k * SSSP_OW + tid, rb[k], re[k], false, (hasrem >> k) & 1u, pend[k])) did = true;
                    quiet = 0;
                    continue;
                }
                // the mailbox count as the publisher warp last saw it: new remote notifications -> look at our vertices' counters
                const int mbx = __shfl_sync(0xffffffffu, *vmbox, 0);
                if (mbx != mb_seen) {
                    if (!awake) { awake = true; if (lane == 0) atomicAdd(&s_nactive, 1); }
                    mb_seen = mbx;
                    int cg[G];
#pragma unroll
                    for (int k = 0; k < G; ++k) {
                        const int v = v0 + k * SSSP_OW + tid;
                        cg[k] = v < n ? __ldcg(dirty + v) : seenG[k];
                    }
#pragma unroll
                    for (int k = 0; k < G; ++k) {
                        const int lv = k * SSSP_OW + tid;
                        if (cg[k] != seenG[k]) {
                            seenG[k] = cg[k];
                            seenL[k] = vflag[lv];
                            relax(lv, rb[k], re[k], true, (hasrem >> k) & 1u, pend[k]);
                            did = true;                  // (a consumed notification counts as work: the round is not quiescent)
                        }
                    }
                    quiet = 0;
                    continue;
                }
                if (awake) { awake = false; __syncwarp(); if (lane == 0) atomicSub(&s_nactive, 1); }
                const int na = __shfl_sync(0xffffffffu, *vnact, 0);
                quiet = na == 0 ? quiet + 1 : 0;
                if (quiet >= 2) break;
            }
            if (awake) { __syncwarp(); if (lane == 0) atomicSub(&s_nactive, 1); }
            float pmin = ST_INF;
#pragma unroll
            for (int k = 0; k < G; ++k) pmin = fminf(pmin, pend[k]);
            for (int o = 16; o; o >>= 1) pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
            if (lane == 0 && pmin < ST_INF) atomicMin(&s_minpend[round & 1], __float_as_uint(pmin));
        } else {
            // publisher warps: keep the CTA's copy of the mailbox count current and drain the publish bitmap (lane = one word)
            const int wi = tid - SSSP_OW;
            bool awake = false;
            int quiet = 0;
            for (int q = 0; q < nlocal; ++q) {
                if (wi == 0) {
                    const int mbc = *(volatile int *)mybox >> 2;
                    if (mbc != *vmbox) { *vmbox = mbc; did = true; }
                }
                bool found = false;                      // (only the words this warp drains: the other publisher warp may already have left the round)
                for (int j = ((tid - SSSP_OW) >> 5) + 2 * lane; j < NW; j += 64) found |= vbits[j] != 0u;
                if (__any_sync(0xffffffffu, found)) {
                    if (!awake) { awake = true; if (lane == 0) atomicAdd(&s_nactive, 1); }
                    // a word = 32 consecutive vertices = the ones that improve together: the warp takes one word at a time,
                    // lane i publishes vertex i of it
                    const int pw = (tid - SSSP_OW) >> 5;
                    for (int j = pw; j < NW; j += (1024 - SSSP_OW) / 32) {
                        unsigned bits = 0;
                        if (lane == 0 && vbits[j] != 0u) bits = atomicExch(rbits + j, 0u);
                        bits = __shfl_sync(0xffffffffu, bits, 0);
                        if ((bits >> lane) & 1u) publish(j * 32 + lane);
                    }
                    did = true;
                    quiet = 0;
                    continue;
                }
                if (awake) { awake = false; __syncwarp(); if (lane == 0) atomicSub(&s_nactive, 1); }
                const int na = __shfl_sync(0xffffffffu, *vnact, 0);
                quiet = na == 0 ? quiet + 1 : 0;
                if (quiet >= 2) break;
            }
            if (awake) { __syncwarp(); if (lane == 0) atomicSub(&s_nactive, 1); }
        }
        { const long long t = clock64(); ph[0] += t - t0; t0 = t; }
        __syncthreads();
        // a warp that left before the last mailbox update has not looked at its counters for it: not quiescent
        const int mb_final = *vmbox;
        if (owner && mb_seen != mb_final) did = true;
        const int anyw = __syncthreads_or(did);
        const unsigned mp = s_minpend[round & 1];
        if (tid == 0) s_minpend[(round + 1) & 1] = 0x7F800000u;
        __syncthreads();
        { const long long t = clock64(); ph[2] += t - t0; t0 = t; }
        if (anyw) continue;
        // nothing moved below T: publish the smallest parked candidate and go idle (see the kernel comment)
        if (tid == 0) {
            const int par = epoch & 1;
            if (mp != 0x7F800000u) atomicMin(&ctl->gmin[par], mp);
            const int old = atomicOr(mybox, 1 | (par << 1));
            if ((old >> 2) != mb_final) {
                // a notification arrived before the idle bit: back to work.  If a notifier has already cleared the bit it
                // has also taken us out of the count -- which we never entered
                const int old2 = atomicAnd(mybox, ~3);
                if (!(old2 & 1)) atomicAdd(&ctl->idle_count[par], 1);
                s_state = 0;
            } else if (atom_add_release(&ctl->idle_count[par], 1) == NB - 1) {
                // every CTA is idle in this epoch and no notification is pending: next threshold, or the fixed point
                const unsigned gm = atomicExch(&ctl->gmin[par], 0x7F800000u);
                if (gm == 0x7F800000u) {
                    *(volatile unsigned long long *)&ctl->epoch_T = (unsigned long long)ST_EPOCH_DONE << 32;
                    s_state = 1;
                } else {
                    const float Tn = fmaxf(T, __uint_as_float(gm)) + delta;
                    if ((flags & 4) && epoch < 1024) {
                        g_sssp_epoch_log[epoch][0] = global_timer();
                        g_sssp_epoch_log[epoch][1] = blockIdx.x;
                        g_sssp_epoch_log[epoch][2] = (unsigned long long)round;
                        g_sssp_epoch_log[epoch][3] = t_epoch_wake;
                    }
                    *(volatile unsigned long long *)&ctl->epoch_T = ((unsigned long long)(unsigned)(epoch + 1) << 32) | __float_as_uint(Tn);
                    const int old2 = atomicAnd(mybox, ~3);
                    if (old2 & 1) atomicSub(&ctl->idle_count[(old2 >> 1) & 1], 1);
                    s_state = 0;
                }
            } else {
                for (;;) {
                    const int mb = *(volatile int *)mybox;
                    const int ep = (int)(*(volatile unsigned long long *)&ctl->epoch_T >> 32);
                    if (ep == ST_EPOCH_DONE) { s_state = 1; break; }
                    if (!(mb & 1)) { s_state = 0; break; }                                        // woken by a notifier
                    if (ep != epoch) {                                                              // next threshold
                        const int old2 = atomicAnd(mybox, ~3);
                        if (old2 & 1) atomicSub(&ctl->idle_count[(old2 >> 1) & 1], 1);
                        s_state = 0;
                        break;
                    }
                    __nanosleep(32);
                }
            }
            if (blockIdx.x == 0) ctl->rounds = round + 1;
        }
        __syncthreads();
        { const long long t = clock64(); ph[3] += t - t0; t0 = t; }
        if (s_state == 1) break;
    }
    // diagnostics
    for (int o = 16; o; o >>= 1) {
        n_relax += __shfl_xor_sync(0xffffffffu, n_relax, o);
        n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
        n_remote += __shfl_xor_sync(0xffffffffu, n_remote, o);
    }
    if (lane == 0) {
        atomicAdd(&ctl->stats[0], (unsigned long long)n_relax);
        atomicAdd(&ctl->stats[1], (unsigned long long)n_acc);
        atomicAdd(&ctl->stats[2], (unsigned long long)n_remote);
    }
    if (blockIdx.x == 0 && tid == 0) ctl->stats[3] = (unsigned long long)epoch + 1;
    if (tid == 0) {
        const unsigned long long tot = (unsigned long long)(clock64() - t_begin);
        for (int i = 0; i < 4; ++i) atomicAdd(&ctl->phase[i], (unsigned long long)ph[i]);
        atomicAdd(&ctl->phase[4], tot);
        atomicMax(&ctl->phase[5], (unsigned long long)(ph[0] + ph[1] + ph[2]));      // busiest CTA: cycles outside the idle spin
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

constexpr int ST_SSSP_MAX_CTAS = 960;      // idle flags of k_sssp_blob in the workspace tail (4 KB)

template <int G>
static int launch_sssp_blob(int blocks_needed, int device, void **args, cudaStream_t s, bool &launched) {
    const size_t smem = (size_t)8 * SSSP_OW * G + 4 * (size_t)((SSSP_OW * G + 31) / 32);
    static bool attr = false;
    if (!attr) { ST_CHECK_CUDA(cudaFuncSetAttribute((const void *)k_sssp_blob<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    int per_sm = 0, sms = 0;
    ST_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k_sssp_blob<G>, 1024, smem));
    ST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    launched = false;
    if (per_sm * sms < blocks_needed) return ST_OK;
    ST_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_sssp_blob<G>, dim3(blocks_needed), dim3(1024), args, smem, s));
    launched = true;
    return ST_OK;
}

extern "C" int st_sssp(const int32_t *row_ptr, const int32_t *col, const float *w, int64_t n, const int32_t *sources,
                       int32_t n_sources, float delta, float *dist_out, int32_t *pred, int32_t *sweeps_host, void *ctl_workspace,
                       const int32_t *orig_id, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (sweeps_host) *sweeps_host = 0;
    if (n == 0) return ST_OK;
    ST_REQUIRE(ctl_workspace != nullptr, "ctl_workspace (4608 + 16n bytes) required");
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
    int pairs = 1;                         // relax two woken vertices at a time, one per half warp (2.65 -> 2.43 ms on the bench tree)
    if (const char *e = getenv("ST_SSSP_PAIRS")) pairs = atoi(e) != 0;
    void *args[] = {(void *)&row_ptr, (void *)&col, (void *)&w, (void *)&nn, (void *)&dist, (void *)&dirty, (void *)&ctl, (void *)&delta, (void *)&npass, (void *)&adv,
                    (void *)&pairs};
    // near-far counter variant: poll state in registers when every resident warp can own its vertices there, in global
    // memory otherwise (ctl_workspace holds dirty[n], seen[n], pend[n]); ST_SSSP_FLAGS=1 forces the old flag variant
    const bool small = (int64_t)blocks * 8 * SSSP_G * 32 >= n && !getenv("ST_SSSP_FORCE_BIG");      // (env: tests exercise the large-graph kernel)
    // CTA-local propagation (k_sssp_blob): the smallest range per CTA for which the whole graph is resident
    bool local_done = false;
    // Opt-in (ST_SSSP_LOCAL=1): measured on the bench tree 13 ms against 4.4 ms of k_sssp -- a hop inside a range costs ~7 us
    // (the two walks over a vertex' arcs through L1/L2), a hop across a range boundary ~80 us, and every epoch waits for
    // its slowest range (profiles/README.md, tools/sssp_trace.py).
    const char *loc_env = getenv("ST_SSSP_LOCAL");
    if (loc_env && atoi(loc_env) != 0 && !getenv("ST_SSSP_FORCE_BIG") && !getenv("ST_SSSP_FLAGS")) {
        int nlocal = 1 << 30;                 // cap of the polls per round (a round ends when the CTA is quiet; tests force short rounds)
        if (const char *e = getenv("ST_SSSP_NLOCAL")) { int v = atoi(e); if (v >= 1) nlocal = v; }
        float bdelta = delta;                 // threshold step per epoch (any value is exact)
        if (const char *e = getenv("ST_SSSP_BLOB_DELTA")) { float v = (float)atof(e); if (v > 0.f) bdelta = v; }
        SsspBlobCtl *bctl = (SsspBlobCtl *)((char *)ctl_workspace + 256 + 16 * (size_t)n);
        int bflags = 0;
        if (const char *e = getenv("ST_SSSP_BLOB_FLAGS")) bflags = atoi(e);
        void *largs[] = {(void *)&row_ptr, (void *)&col, (void *)&w, (void *)&nn, (void *)&dist, (void *)&dirty, (void *)&bctl, (void *)&bdelta,
                         (void *)&nlocal, (void *)&bflags};
        int gsel = 0;
        if (const char *e = getenv("ST_SSSP_LOCAL_G")) gsel = atoi(e);
        for (int G : {1, 2, 4, 8}) {
            if (local_done || (gsel && G != gsel)) continue;
            const int need = (int)cdiv(n, SSSP_OW * (int64_t)G);
            if (need > ST_SSSP_MAX_CTAS) continue;
            k_sssp_blob_init<<<1, 256, 0, s>>>(bctl, need, bdelta);
            ST_CHECK_LAUNCH();
            int rc2 = ST_OK;
            if (G == 1) rc2 = launch_sssp_blob<1>(need, device, largs, s, local_done);
            else if (G == 2) rc2 = launch_sssp_blob<2>(need, device, largs, s, local_done);
            else if (G == 4) rc2 = launch_sssp_blob<4>(need, device, largs, s, local_done);
            else rc2 = launch_sssp_blob<8>(need, device, largs, s, local_done);
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
        if (local_done) ST_CHECK_CUDA(cudaMemcpy(&chunks, &((SsspBlobCtl *)((char *)ctl_workspace + 256 + 16 * (size_t)n))->rounds, 4, cudaMemcpyDeviceToHost));
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
extern "C" int st_debug_sssp_blob_stats(const void *ctl_workspace, int64_t n, unsigned long long *out_host) {
    const SsspBlobCtl *c = (const SsspBlobCtl *)((const char *)ctl_workspace + 256 + 16 * (size_t)n);
    ST_CHECK_CUDA(cudaMemcpy(out_host, c->stats, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost));
    int r = 0;
    ST_CHECK_CUDA(cudaMemcpy(&r, &c->rounds, 4, cudaMemcpyDeviceToHost));
    out_host[4] = (unsigned long long)r;
    ST_CHECK_CUDA(cudaMemcpy(out_host + 5, c->phase, sizeof(unsigned long long) * 6, cudaMemcpyDeviceToHost));
    return ST_OK;
}

extern "C" int st_debug_sssp_epoch_log(unsigned long long *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_sssp_epoch_log, sizeof(unsigned long long) * 4096));
    return ST_OK;
}

extern "C" int st_debug_sssp_tfinal(unsigned *out_host, int64_t n) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_sssp_tfinal, sizeof(unsigned) * (size_t)(n < (1 << 20) ? n : (1 << 20))));
    return ST_OK;
}

extern "C" int st_debug_sssp_stats(const void *ctl_workspace, unsigned *out_host) {
    const SsspCtl *c = (const SsspCtl *)ctl_workspace;
    ST_CHECK_CUDA(cudaMemcpy(out_host, &c->chunks, sizeof(unsigned) * 5, cudaMemcpyDeviceToHost));
    return ST_OK;
}
