// Fixed-radius kNN on a uniform grid (counting sort of the destination points into cells,
// one thread per query with a register-resident sorted top-K), outlier mask, edge list.
// Exactness contract (oracle/skeleton_ref.py): d2 = (dx*dx+dy*dy)+dz*dz in fp32 without FMA
// contraction, candidates need d2 < fl32(r*r), order = ascending (d2, index).
#include "grid.cuh"

using namespace st;

__global__ void k_fill_f32(float *p, int64_t n, float v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__device__ __forceinline__ bool cand_less(float d2a, int ia, float d2b, int ib) { return d2a < d2b || (d2a == d2b && ia < ib); }

template <int K>
__global__ void __launch_bounds__(128) k_knn(const float *__restrict__ src, int n, Grid g, const int32_t *__restrict__ cell_start,
                                             const float4 *__restrict__ sorted, float r, const float *__restrict__ qrad,
                                             int32_t *__restrict__ idx, float *__restrict__ d2out, int self_query) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float qx, qy, qz;
    if (self_query) {
        // queries == grid points: walk them in cell order, so a warp's 32 queries scan (nearly) the same
        // cells -- coherent branches, L1-resident candidates -- and write to their own rows
        const float4 q = __ldg(sorted + i);
        qx = q.x; qy = q.y; qz = q.z;
        i = __float_as_int(q.w);
    } else {
        qx = src[3 * (size_t)i]; qy = src[3 * (size_t)i + 1]; qz = src[3 * (size_t)i + 2];
    }
    const float r2 = __fmul_rn(r, r);
    const float qr = qrad ? qrad[i] : 0.f;
    float reach = qrad ? fminf(r, qr) : r;
    float bd[K];
    int bi[K];
#pragma unroll
    for (int s = 0; s < K; ++s) { bd[s] = FLT_MAX; bi[s] = INT_MAX; }
    // Rows of cells nearest first; once K candidates are known the search radius shrinks to the K-th best, so in
    // the dense parts of the cloud (thousands of medial points within r of a trunk vertex) only the few cells
    // around the query are read.  Ties at the K-th distance stay reachable (rows are skipped only when strictly
    // farther), so the result is the exact ascending (d2, index) top-K.
    float rr = reach * 1.0001f + 1e-7f;
    float lim2 = rr * rr;
    for_rows_near_first(g, cell_start, qx, qy, qz, reach, lim2, [&](int beg, int end) {
        for (int t = beg; t < end; ++t) {
            float4 p = __ldg(sorted + t);
            float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
            if (!(d2 < r2)) continue;
            if (qrad && sqrtf(d2) > qr) continue;
            int j = __float_as_int(p.w);
            if (!cand_less(d2, j, bd[K - 1], bi[K - 1])) continue;
#pragma unroll
            for (int s = K - 1; s >= 0; --s) {
                bool lt_prev = (s > 0) && cand_less(d2, j, bd[s > 0 ? s - 1 : 0], bi[s > 0 ? s - 1 : 0]);
                bool lt_cur = cand_less(d2, j, bd[s], bi[s]);
                float nd = lt_prev ? bd[s > 0 ? s - 1 : 0] : (lt_cur ? d2 : bd[s]);
                int ni = lt_prev ? bi[s > 0 ? s - 1 : 0] : (lt_cur ? j : bi[s]);
                bd[s] = nd;
                bi[s] = ni;
            }
        }
        if (bi[K - 1] != INT_MAX) lim2 = fminf(lim2, bd[K - 1]);
    });
#pragma unroll
    for (int s = 0; s < K; ++s) {
        bool ok = bi[s] != INT_MAX;
        idx[(size_t)i * K + s] = ok ? bi[s] : -1;
        d2out[(size_t)i * K + s] = ok ? bd[s] : -1.f;
    }
}

extern "C" size_t st_knn_workspace_bytes(int64_t m) { return grid_ws_bytes(m); }

extern "C" int st_knn(const float *src, int64_t n, const float *dst, int64_t m, int K, float r, const float *query_radius,
                      int32_t *idx, float *d2, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31) && m < (1ll << 31), "sizes");
    if (m == 0 || !(r > 0.f)) {
        ST_CHECK_CUDA(cudaMemsetAsync(idx, 0xFF, n * K * 4, s));
        // d2 = -1.0f
        k_fill_f32<<<(unsigned)cdiv(n * K, 256), 256, 0, s>>>(d2, n * K, -1.f);
        ST_CHECK_LAUNCH();
        return ST_OK;
    }
    Carver cv(workspace, workspace_bytes);
    GridBuild gb;
    // with per-query radii a finer grid pays: thin branches search a few small cells
    float h = query_radius ? r * 0.25f : r;
    int rc = build_grid(dst, m, h, cv, gb, s);
    if (rc) return rc;
    unsigned g = (unsigned)cdiv(n, 128);
    const int self_query = (src == dst && n == m) ? 1 : 0;
#define ST_K(KK) case KK: k_knn<KK><<<g, 128, 0, s>>>(src, (int)n, gb.g, gb.cell_start, gb.sorted, r, query_radius, idx, d2, self_query); break;
    switch (K) {
        ST_K(1) ST_K(2) ST_K(4) ST_K(8) ST_K(16) ST_K(24) ST_K(32) ST_K(48) ST_K(64)
        default: set_error("st_knn: K=%d not instantiated (1,2,4,8,16,24,32,48,64; the Python wrapper rounds up and slices)", K); return ST_ERR_UNSUPPORTED;
    }
#undef ST_K
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ outlier mask
__global__ void __launch_bounds__(128) k_outlier(const float *__restrict__ pts, int n, Grid g, const int32_t *__restrict__ cell_start,
                                                 const float4 *__restrict__ sorted, float r, const float *__restrict__ radii, int nb,
                                                 uint8_t *__restrict__ keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 qq = __ldg(sorted + i);          // cell order: coherent warps (see k_knn)
    const float qx = qq.x, qy = qq.y, qz = qq.z;
    i = __float_as_int(qq.w);
    const float r2 = __fmul_rn(r, r);
    const float qr = radii[i];
    // the nb nearest all lie strictly inside radius_i  <=>  at least nb points with d2 < r^2 and sqrt(d2) < radius_i
    float reach = fminf(r, qr);
    int cnt = 0;
    if (reach > 0.f) {
        const float rr = reach * 1.0001f + 1e-7f;
        float lim2 = rr * rr;
        for_rows_near_first(g, cell_start, qx, qy, qz, reach, lim2, [&](int beg, int end) {
            for (int t = beg; t < end && cnt < nb; ++t) {
                float4 p = __ldg(sorted + t);
                float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                if (d2 < r2 && sqrtf(d2) < qr) ++cnt;
            }
            if (cnt >= nb) lim2 = -1.f;       // enough neighbours found: skip every remaining row
        });
    }
    keep[i] = cnt >= nb ? 1 : 0;
}

extern "C" int st_outlier_mask(const float *points, int64_t n, const float *radii, float r_max, int nb, uint8_t *keep,
                               void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    if (!(r_max > 0.f)) { ST_CHECK_CUDA(cudaMemsetAsync(keep, 0, n, s)); return ST_OK; }
    Carver cv(workspace, workspace_bytes);
    GridBuild gb;
    int rc = build_grid(points, n, r_max * 0.25f, cv, gb, s);
    if (rc) return rc;
    k_outlier<<<(unsigned)cdiv(n, 128), 128, 0, s>>>(points, (int)n, gb.g, gb.cell_start, gb.sorted, r_max, radii, nb, keep);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ edges from kNN
__global__ void k_edge_flags(const int32_t *__restrict__ idx, const float *__restrict__ d2, int64_t total, int K,
                             const float *__restrict__ radii, int32_t *__restrict__ flag) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int j = idx[t];
    bool ok = j > 0;  // sic: graph.py:59 `idxs > 0`
    if (ok && radii) ok = !(sqrtf(d2[t]) > radii[t / K]);
    flag[t] = ok ? 1 : 0;
}

__global__ void k_edge_emit(const int32_t *__restrict__ idx, const float *__restrict__ d2, int64_t total, int K,
                            const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, int32_t *__restrict__ edges,
                            float *__restrict__ weights, int64_t *__restrict__ n_edges) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    if (flag[t]) {
        int p = pos[t];
        edges[2 * (size_t)p] = (int)(t / K);
        edges[2 * (size_t)p + 1] = idx[t];
        weights[p] = sqrtf(d2[t]);
    }
    if (t == total - 1) *n_edges = (int64_t)pos[t] + flag[t];
}

extern "C" size_t st_edges_workspace_bytes(int64_t n, int K) {
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (int *)nullptr, (int *)nullptr, (int)(n * K));
    return align_up(scan) + 2 * align_up(n * K * 4) + 1024;
}

extern "C" int st_edges_from_knn(const int32_t *idx, const float *d2, int64_t n, int K, const float *radii, int32_t *edges,
                                 float *weights, int64_t *n_edges_host, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    *n_edges_host = 0;
    int64_t total = n * K;
    if (total == 0) return ST_OK;
    ST_REQUIRE(total < (1ll << 31), "n*K");
    Carver cv(workspace, workspace_bytes);
    int32_t *flag = cv.take<int32_t>(total);
    int32_t *pos = cv.take<int32_t>(total);
    int64_t *n_dev = cv.take<int64_t>(1);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, pos, (int)total);
    void *scan_ws = cv.take<char>(scan_bytes);
    if (!cv.ok()) { set_error("st_edges_from_knn: workspace too small"); return ST_ERR_WORKSPACE; }
    unsigned g = (unsigned)cdiv(total, 256);
    k_edge_flags<<<g, 256, 0, s>>>(idx, d2, total, K, radii, flag);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, flag, pos, (int)total, s));
    k_edge_emit<<<g, 256, 0, s>>>(idx, d2, total, K, flag, pos, edges, weights, n_dev);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cudaMemcpyAsync(n_edges_host, n_dev, 8, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    return ST_OK;
}
