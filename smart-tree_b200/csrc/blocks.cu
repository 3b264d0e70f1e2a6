// Block tiling of a cloud for inference (reference: SingleTreeInference.compute_blocks,
// smart_tree/dataset/dataset.py:166-190 + cube_filter, smart_tree/util/maths.py:145-155).
// The reference builds one O(N) boolean mask per block in a host loop; here the kept blocks are found
// with one sort + run-length pass and every (block, point) membership pair is emitted by one kernel,
// block-major with the original point order inside a block.
#include <cub/cub.cuh>

#include "common.cuh"

using namespace st;

constexpr int BK_OFF = 1 << 20;

// torch.div(a, b, rounding_mode="floor") for floats (c10::div_floor_floating)
__device__ __forceinline__ float div_floor(float a, float b) {
    if (b == 0.f) return a / b;
    float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.f && ((b < 0.f) != (mod < 0.f))) div = __fsub_rn(div, 1.f);
    float fl;
    if (div != 0.f) {
        fl = floorf(div);
        if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.f);
    } else {
        fl = copysignf(0.f, a / b);
    }
    return fl;
}

__device__ __forceinline__ unsigned long long bkey(int x, int y, int z) {
    return ((unsigned long long)(unsigned)(x + BK_OFF) << 42) | ((unsigned long long)(unsigned)(y + BK_OFF) << 21) |
           (unsigned long long)(unsigned)(z + BK_OFF);
}

__global__ void k_blk_keys(const float *__restrict__ xyz, int n, float bs, unsigned long long *__restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int qx = (int)div_floor(xyz[3 * (size_t)i], bs), qy = (int)div_floor(xyz[3 * (size_t)i + 1], bs), qz = (int)div_floor(xyz[3 * (size_t)i + 2], bs);
    keys[i] = bkey(qx, qy, qz);
}

__global__ void k_blk_filter(const unsigned long long *__restrict__ ukeys, const int *__restrict__ counts, const int *__restrict__ nruns,
                             int min_points, unsigned long long *__restrict__ kept, float *__restrict__ ids, int cap, int *n_kept) {
    // single CTA: ordered compaction of the (already sorted) runs
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int runs = *nruns;
    for (int start = 0; start < runs; start += blockDim.x) {
        int r = start + threadIdx.x;
        bool ok = r < runs && counts[r] > min_points;
        // block-wide exclusive scan of ok via ballots
        unsigned bal = __ballot_sync(0xffffffffu, ok);
        __shared__ int s_w[32];
        int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        if (lane == 0) s_w[wp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < wp; ++k) off += s_w[k];
        int pos = off + __popc(bal & ((1u << lane) - 1));
        if (ok && pos < cap) {
            unsigned long long k = ukeys[r];
            kept[pos] = k;
            ids[3 * pos] = (float)((int)((k >> 42) & 0x1FFFFF) - BK_OFF);
            ids[3 * pos + 1] = (float)((int)((k >> 21) & 0x1FFFFF) - BK_OFF);
            ids[3 * pos + 2] = (float)((int)(k & 0x1FFFFF) - BK_OFF);
        }
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_w[k]; s_base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_kept = s_base;
}

static size_t blk_cub_bytes(int64_t n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr,
                                       (int *)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (int *)nullptr, (int *)nullptr, (int)n);
    return (a > b ? (a > c ? a : c) : (b > c ? b : c));
}

extern "C" size_t st_block_workspace_bytes(int64_t n, int64_t n_pairs) {
    int64_t m = n > n_pairs ? n : n_pairs;
    return align_up(blk_cub_bytes(m)) + 3 * align_up(m * 8) + 2 * align_up(n * 4) + 8192;
}

// Step 1: kept blocks (count > min_points), sorted by (x,y,z) block id like torch.unique(dim=0).
// block_ids[cap,3] receives the floor-divided coordinates as floats; kept_keys[cap] is an opaque
// sorted key list for step 2.
extern "C" int st_block_list(const float *xyz, int64_t n, float block_size, int min_points, float *block_ids,
                             uint64_t *kept_keys, int32_t cap, int64_t *n_blocks_host, void *workspace, size_t workspace_bytes,
                             void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    *n_blocks_host = 0;
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31), "n");
    Carver cv(workspace, workspace_bytes);
    unsigned long long *keys = cv.take<unsigned long long>(n);
    unsigned long long *sorted = cv.take<unsigned long long>(n);
    unsigned long long *ukeys = cv.take<unsigned long long>(n);
    int *counts = cv.take<int>(n);
    int *nruns = cv.take<int>(2);
    size_t cb = blk_cub_bytes(n);
    void *cub_ws = cv.take<char>(cb);
    if (!cv.ok()) { set_error("st_block_list: workspace too small"); return ST_ERR_WORKSPACE; }
    k_blk_keys<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(xyz, (int)n, block_size, keys);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(cub_ws, cb, keys, sorted, (int)n, 0, 63, s));
    ST_CHECK_CUDA(cub::DeviceRunLengthEncode::Encode(cub_ws, cb, sorted, ukeys, counts, nruns, (int)n, s));
    k_blk_filter<<<1, 1024, 0, s>>>(ukeys, counts, nruns, min_points, (unsigned long long *)kept_keys, block_ids, cap, nruns + 1);
    ST_CHECK_LAUNCH();
    int h = 0;
    ST_CHECK_CUDA(cudaMemcpyAsync(&h, nruns + 1, 4, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    if (h > cap) { set_error("st_block_list: %d blocks exceed capacity %d", h, cap); return ST_ERR_WORKSPACE; }
    *n_blocks_host = h;
    return ST_OK;
}

struct BlkArgs {
    const float *xyz;
    int n;
    const unsigned long long *kept;
    int nb;
    float bs, half_bs, half_cube;
    int reach;
};

__device__ __forceinline__ int find_block(const unsigned long long *__restrict__ kept, int nb, unsigned long long key) {
    int lo = 0, hi = nb;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(kept + mid) < key) lo = mid + 1; else hi = mid; }
    return (lo < nb && __ldg(kept + lo) == key) ? lo : -1;
}

// cube_filter(points, centre, cube): centre - cube/2 <= p < centre + cube/2, all in fp32
__device__ __forceinline__ bool in_cube(float p, int q, float bs, float half_bs, float half_cube) {
    float centre = __fadd_rn(__fmul_rn((float)q, bs), half_bs);
    return p >= __fsub_rn(centre, half_cube) && p < __fadd_rn(centre, half_cube);
}

template <bool EMIT>
__global__ void k_blk_members(BlkArgs a, const int *__restrict__ offs, int *__restrict__ cnt, unsigned long long *__restrict__ pairs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float px = a.xyz[3 * (size_t)i], py = a.xyz[3 * (size_t)i + 1], pz = a.xyz[3 * (size_t)i + 2];
    int qx = (int)div_floor(px, a.bs), qy = (int)div_floor(py, a.bs), qz = (int)div_floor(pz, a.bs);
    int c = 0;
    int o = EMIT ? offs[i] : 0;
    for (int dx = -a.reach; dx <= a.reach; ++dx) {
        if (!in_cube(px, qx + dx, a.bs, a.half_bs, a.half_cube)) continue;
        for (int dy = -a.reach; dy <= a.reach; ++dy) {
            if (!in_cube(py, qy + dy, a.bs, a.half_bs, a.half_cube)) continue;
            for (int dz = -a.reach; dz <= a.reach; ++dz) {
                if (!in_cube(pz, qz + dz, a.bs, a.half_bs, a.half_cube)) continue;
                int b = find_block(a.kept, a.nb, bkey(qx + dx, qy + dy, qz + dz));
                if (b < 0) continue;
                if (EMIT) pairs[o + c] = ((unsigned long long)(unsigned)b << 32) | (unsigned)i;
                ++c;
            }
        }
    }
    if (!EMIT) cnt[i] = c;
}

__global__ void k_blk_unpack(const unsigned long long *__restrict__ pairs, int64_t t, const float *__restrict__ xyz,
                             int64_t *__restrict__ point_index, int32_t *__restrict__ point_block, int *__restrict__ lo, int *__restrict__ hi) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = k < t;
    int b = -1, i = 0;
    if (live) {
        unsigned long long p = pairs[k];
        b = (int)(p >> 32);
        i = (int)(p & 0xFFFFFFFFull);
        point_index[k] = i;
        point_block[k] = b;
    }
    // per-block bounding box of the member points; one atomic per (warp, block)
    unsigned peers = __match_any_sync(0xffffffffu, b);
    int lane = threadIdx.x & 31;
    bool leader = (__ffs(peers) - 1) == lane;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        float v = live ? xyz[3 * (size_t)i + ax] : 0.f;
        int ov = __float_as_int(v);
        ov = ov >= 0 ? ov : ov ^ 0x7FFFFFFF;
        int mn = __reduce_min_sync(peers, ov), mx = __reduce_max_sync(peers, ov);
        if (live && leader) { atomicMin(lo + 3 * b + ax, mn); atomicMax(hi + 3 * b + ax, mx); }
    }
}

__global__ void k_blk_bbox_init(int *lo, int *hi, int n3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) { lo[i] = INT_MAX; hi[i] = INT_MIN; }
}
__global__ void k_blk_bbox_fin(int *lo, int *hi, int n3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    int l = lo[i], h = hi[i];
    lo[i] = l >= 0 ? l : l ^ 0x7FFFFFFF;     // back to float bits, in place
    hi[i] = h >= 0 ? h : h ^ 0x7FFFFFFF;
}

// Step 2a: number of (block, point) membership pairs.
extern "C" int st_block_count(const float *xyz, int64_t n, const uint64_t *kept_keys, int32_t n_blocks, float block_size,
                              float half_block, float half_cube, int reach, int32_t *offsets /*[n]*/, int64_t *n_pairs_host,
                              void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    *n_pairs_host = 0;
    if (n == 0 || n_blocks == 0) return ST_OK;
    Carver cv(workspace, workspace_bytes);
    int *cnt = cv.take<int>(n);
    size_t cb = blk_cub_bytes(n);
    void *cub_ws = cv.take<char>(cb);
    if (!cv.ok()) { set_error("st_block_count: workspace too small"); return ST_ERR_WORKSPACE; }
    BlkArgs a{xyz, (int)n, (const unsigned long long *)kept_keys, n_blocks, block_size, half_block, half_cube, reach};
    k_blk_members<false><<<(unsigned)cdiv(n, 256), 256, 0, s>>>(a, nullptr, cnt, nullptr);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cb, cnt, offsets, (int)n, s));
    int last_off = 0, last_cnt = 0;
    ST_CHECK_CUDA(cudaMemcpyAsync(&last_off, offsets + n - 1, 4, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaMemcpyAsync(&last_cnt, cnt + n - 1, 4, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    *n_pairs_host = (int64_t)last_off + last_cnt;
    return ST_OK;
}

// Step 2b: the pairs, block-major with ascending point index inside a block, plus each block's
// member bounding box (block_lo / block_hi [n_blocks,3] fp32).
extern "C" int st_block_emit(const float *xyz, int64_t n, const uint64_t *kept_keys, int32_t n_blocks, float block_size,
                             float half_block, float half_cube, int reach, const int32_t *offsets, int64_t n_pairs,
                             int64_t *point_index, int32_t *point_block, float *block_lo, float *block_hi, void *workspace,
                             size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0 || n_blocks == 0 || n_pairs == 0) return ST_OK;
    ST_REQUIRE(n_pairs < (1ll << 31), "n_pairs");
    Carver cv(workspace, workspace_bytes);
    unsigned long long *pairs = cv.take<unsigned long long>(n_pairs);
    unsigned long long *sorted = cv.take<unsigned long long>(n_pairs);
    size_t cb = blk_cub_bytes(n_pairs);
    void *cub_ws = cv.take<char>(cb);
    if (!cv.ok()) { set_error("st_block_emit: workspace too small"); return ST_ERR_WORKSPACE; }
    BlkArgs a{xyz, (int)n, (const unsigned long long *)kept_keys, n_blocks, block_size, half_block, half_cube, reach};
    k_blk_members<true><<<(unsigned)cdiv(n, 256), 256, 0, s>>>(a, offsets, nullptr, pairs);
    ST_CHECK_LAUNCH();
    int bits = 1;
    while ((1ll << bits) < n_blocks) ++bits;
    // pairs are emitted in ascending point order; a stable sort on the block bits alone makes them block-major
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(cub_ws, cb, pairs, sorted, (int)n_pairs, 32, 32 + bits, s));
    int n3 = 3 * n_blocks;
    k_blk_bbox_init<<<(unsigned)cdiv(n3, 256), 256, 0, s>>>((int *)block_lo, (int *)block_hi, n3);
    ST_CHECK_LAUNCH();
    k_blk_unpack<<<(unsigned)cdiv(n_pairs, 256), 256, 0, s>>>(sorted, n_pairs, xyz, point_index, point_block, (int *)block_lo, (int *)block_hi);
    ST_CHECK_LAUNCH();
    k_blk_bbox_fin<<<(unsigned)cdiv(n3, 256), 256, 0, s>>>((int *)block_lo, (int *)block_hi, n3);
    ST_CHECK_LAUNCH();
    return ST_OK;
}
