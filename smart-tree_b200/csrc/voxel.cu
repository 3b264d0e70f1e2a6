// Voxelise, coordinate hash table, sub-manifold neighbour map, strided/inverse conv maps.
// HBM-bound integer work: one thread per (tap, voxel) with the voxel index fastest so every
// map row is written coalesced; the hash table (16 B/slot at load factor <= 0.5) lives in L2.
#include <cub/cub.cuh>
#include <stdarg.h>

#include "common.cuh"

namespace st {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace st

using namespace st;

extern "C" int st_version(void) { return 100; }
extern "C" const char *st_last_error(void) { return st::g_err; }
extern "C" int st_device_check(int device) {
    cudaDeviceProp p;
    ST_CHECK_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        set_error("device %d is sm_%d%d; libst_b200 is built for sm_100a only", device, p.major, p.minor);
        return ST_ERR_UNSUPPORTED;
    }
    return ST_OK;
}
extern "C" int st_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return n;
}

// ------------------------------------------------------------------------------------ hash
extern "C" int64_t st_hash_capacity(int64_t n) {
    int64_t c = 1024;
    while (c < 2 * n) c <<= 1;
    return c;
}

__global__ void k_hash_insert(const int4 *__restrict__ coords, int n, uint64_t *keys, int32_t *vals, uint32_t mask, int32_t *status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = coords[i];
    // the packed key has 16 bits per field (15 for the batch): anything outside would alias another voxel
    if ((unsigned)c.x >= ST_MAX_BATCH || (unsigned)c.y > ST_MAX_COORD || (unsigned)c.z > ST_MAX_COORD || (unsigned)c.w > ST_MAX_COORD) {
        if (status) atomicOr(status, 1);
        return;
    }
    uint64_t key = pack_key(c.x, c.y, c.z, c.w);
    uint32_t s = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS((unsigned long long *)(keys + s), KEY_EMPTY, key);
        if (prev == KEY_EMPTY || prev == key) {
            atomicMin(vals + s, i);  // duplicate coordinates: lowest row wins (deterministic)
            return;
        }
        s = (s + 1) & mask;
    }
}

extern "C" int st_hash_build(const int32_t *coords, int64_t n, uint64_t *keys, int32_t *vals, int64_t capacity, int32_t *status,
                             void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (status) ST_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    ST_REQUIRE(capacity >= 2 * n && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 2n");
    ST_REQUIRE(n < (1ll << 31), "n");
    ST_CHECK_CUDA(cudaMemsetAsync(keys, 0xFF, capacity * sizeof(uint64_t), s));
    ST_CHECK_CUDA(cudaMemsetAsync(vals, 0x7F, capacity * sizeof(int32_t), s));
    if (n) {
        k_hash_insert<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const int4 *)coords, (int)n, keys, vals, (uint32_t)(capacity - 1), status);
        ST_CHECK_LAUNCH();
    }
    return ST_OK;
}

// ------------------------------------------------------------------------------------ subm map
__global__ void k_subm_map(const int4 *__restrict__ coords, int n, const uint64_t *__restrict__ keys,
                           const int32_t *__restrict__ vals, uint32_t mask, int32_t *__restrict__ nbr,
                           const int32_t *__restrict__ shape) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    int dz = k / 9 - 1, dy = (k / 3) % 3 - 1, dx = k % 3 - 1;
    int r;
    if (k == 13) {
        r = i;
    } else {
        const int qz = c.y + dz, qy = c.z + dy, qx = c.w + dx;
        // strict_spconv_bounds: a neighbour location outside the declared spatial shape is not queried (spconv's bound check)
        const bool clipped = shape && (qz >= __ldg(shape) || qy >= __ldg(shape + 1) || qx >= __ldg(shape + 2));
        r = (clipped || qz < 0 || qy < 0 || qx < 0) ? -1 : hash_lookup(keys, vals, mask, pack_key(c.x, qz, qy, qx));
    }
    nbr[(size_t)k * n + i] = r;
}

extern "C" int st_subm_map(const int32_t *coords, int64_t n, const uint64_t *keys, const int32_t *vals,
                           int64_t capacity, const int32_t *spatial_shape, int32_t *nbr, void *stream) {
    if (n == 0) return ST_OK;
    dim3 grid((unsigned)cdiv(n, 256), 27);
    k_subm_map<<<grid, 256, 0, (cudaStream_t)stream>>>((const int4 *)coords, (int)n, keys, vals, (uint32_t)(capacity - 1), nbr, spatial_shape);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ strided coords
// Each input voxel p feeds outputs o = (p+1-k)/2 for taps k with p+1-k even: per axis one
// candidate for even p (k=1), two for odd p (k=0,2) -> at most 8 candidates per voxel.
__global__ void k_strided_candidates(const int4 *__restrict__ coords, int n, uint64_t *__restrict__ cand, int morton) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    int p[3] = {c.y, c.z, c.w};
    int lo[3], cnt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (p[a] & 1) { lo[a] = (p[a] - 1) >> 1; cnt[a] = 2; }
        else          { lo[a] = p[a] >> 1;       cnt[a] = 1; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int jz = j >> 2, jy = (j >> 1) & 1, jx = j & 1;
        uint64_t key = KEY_EMPTY;
        if (jz < cnt[0] && jy < cnt[1] && jx < cnt[2])
            key = morton ? morton_key(c.x, lo[0] + jz, lo[1] + jy, lo[2] + jx) : pack_key(c.x, lo[0] + jz, lo[1] + jy, lo[2] + jx);
        cand[(size_t)j * n + i] = key;
    }
}

// Same candidate set with most duplicates removed at the source.  In shifted coordinates q = p + 1 the outputs fed by
// an input are G - e with G = q >> 1 and e ranging over the subsets of the axes on which q is even.  When the rows
// are in Z-order (the engine's layout: keys interleave the bits of q), the inputs sharing G are CONTIGUOUS, so the
// first row of each run emits the union of the run's candidates once: ~2.5 keys per distinct G instead of 8 per
// input (4 M -> ~0.5 M keys to sort at level 0 of the bench tree).  Any other row order is still correct -- a G
// split over several runs just emits duplicates, which the sort + unique pass removes as before.
__global__ void k_strided_candidates_runs(const int4 *__restrict__ coords, int n, uint64_t *__restrict__ cand, int *__restrict__ n_cand,
                                          int morton, const int32_t *__restrict__ out_shape) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(coords + i);
    const int gz = (c.y + 1) >> 1, gy = (c.z + 1) >> 1, gx = (c.w + 1) >> 1;
    auto same = [&](const int4 &o) { return o.x == c.x && ((o.y + 1) >> 1) == gz && ((o.z + 1) >> 1) == gy && ((o.w + 1) >> 1) == gx; };
    if (i > 0 && same(__ldg(coords + i - 1))) return;
    unsigned offs = 0;                  // bit e (ez,ey,ex) set: output G - e is fed by some row of the run
    for (int j = i; j < n && j < i + 8; ++j) {
        const int4 o = j == i ? c : __ldg(coords + j);
        if (!same(o)) break;
        const unsigned ev = ((~(o.y + 1) & 1) << 2) | ((~(o.z + 1) & 1) << 1) | (~(o.w + 1) & 1);
#pragma unroll
        for (unsigned e = 0; e < 8; ++e)
            if ((e & ~ev) == 0) offs |= 1u << e;
    }
    if (out_shape) {                    // strict_spconv_bounds: outputs outside the declared output shape do not exist
        const int sz = __ldg(out_shape), sy = __ldg(out_shape + 1), sx = __ldg(out_shape + 2);
#pragma unroll
        for (unsigned e = 0; e < 8; ++e)
            if (gz - (int)(e >> 2) >= sz || gy - (int)((e >> 1) & 1) >= sy || gx - (int)(e & 1) >= sx) offs &= ~(1u << e);
    }
    const int cnt = __popc(offs);
    if (cnt == 0) return;
    int pos = atomicAdd(n_cand, cnt);
#pragma unroll
    for (unsigned e = 0; e < 8; ++e) {
        if (!(offs & (1u << e))) continue;
        const int oz = gz - (int)(e >> 2), oy = gy - (int)((e >> 1) & 1), ox = gx - (int)(e & 1);
        cand[pos++] = morton ? morton_key(c.x, oz, oy, ox) : pack_key(c.x, oz, oy, ox);
    }
}

__global__ void k_unpack_coords(const uint64_t *__restrict__ keys, const int *__restrict__ n_sel, int4 *__restrict__ out,
                                int64_t *n_out, int morton) {
    int m = *n_sel;
    if (m > 0 && keys[m - 1] == KEY_EMPTY) --m;  // the sentinel sorts last
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_out = m;
    for (; i < m; i += gridDim.x * blockDim.x) {
        int b, z, y, x;
        if (morton) unpack_morton(keys[i], b, z, y, x); else unpack_key(keys[i], b, z, y, x);
        out[i] = make_int4(b, z, y, x);
    }
}

static size_t strided_cub_bytes(int64_t total) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)total);
    cub::DeviceSelect::Unique(nullptr, b, (uint64_t *)nullptr, (uint64_t *)nullptr, (int *)nullptr, (int)total);
    return a > b ? a : b;
}

extern "C" size_t st_strided_coords_workspace_bytes(int64_t n) {
    int64_t total = 8 * n;
    return align_up(strided_cub_bytes(total)) + 3 * align_up(total * sizeof(uint64_t)) + 1024;
}

extern "C" int st_strided_coords(const int32_t *coords, int64_t n, int morton_order, const int32_t *out_shape, int32_t *out_coords,
                                 int64_t *n_out_host, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    *n_out_host = 0;
    if (n == 0) return ST_OK;
    ST_REQUIRE(8 * n < (1ll << 31), "n too large for one call");
    int64_t total = 8 * n;
    Carver cv(workspace, workspace_bytes);
    uint64_t *cand = cv.take<uint64_t>(total);
    uint64_t *sorted = cv.take<uint64_t>(total);
    uint64_t *uniq = cv.take<uint64_t>(total);
    int *n_sel = cv.take<int>(1);
    int64_t *n_out_dev = cv.take<int64_t>(1);
    size_t cub_bytes = strided_cub_bytes(total);
    void *cub_ws = cv.take<char>(cub_bytes);
    if (!cv.ok()) { set_error("st_strided_coords: workspace too small"); return ST_ERR_WORKSPACE; }
    // run-deduplicated candidates (one small read-back of their number: the sort then handles ~8x fewer keys)
    ST_CHECK_CUDA(cudaMemsetAsync(n_sel, 0, sizeof(int), s));
    k_strided_candidates_runs<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const int4 *)coords, (int)n, cand, n_sel, morton_order, out_shape);
    ST_CHECK_LAUNCH();
    int n_cand = 0;
    ST_CHECK_CUDA(cudaMemcpyAsync(&n_cand, n_sel, sizeof(int), cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    if (n_cand == 0 && out_shape) return ST_OK;      // everything clipped
    ST_REQUIRE(n_cand > 0 && n_cand <= total, "candidate count");
    total = n_cand;
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(cub_ws, cub_bytes, cand, sorted, (int)total, 0, 64, s));
    ST_CHECK_CUDA(cub::DeviceSelect::Unique(cub_ws, cub_bytes, sorted, uniq, n_sel, (int)total, s));
    k_unpack_coords<<<296, 256, 0, s>>>(uniq, n_sel, (int4 *)out_coords, n_out_dev, morton_order);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cudaMemcpyAsync(n_out_host, n_out_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    return ST_OK;
}

__global__ void k_strided_maps(const int4 *__restrict__ coords, int n, int m, const uint64_t *__restrict__ keys,
                               const int32_t *__restrict__ vals, uint32_t mask, int32_t *__restrict__ down,
                               int32_t *__restrict__ up) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    int tz = c.y + 1 - k / 9, ty = c.z + 1 - (k / 3) % 3, tx = c.w + 1 - k % 3;
    int o = -1;
    if (!((tz | ty | tx) & 1)) {
        o = hash_lookup(keys, vals, mask, pack_key(c.x, tz >> 1, ty >> 1, tx >> 1));
        if (o >= 0) down[(size_t)k * m + o] = i;
    }
    up[(size_t)k * n + i] = o;
}

extern "C" int st_strided_maps(const int32_t *coords, int64_t n, int64_t n_out, const uint64_t *out_keys,
                               const int32_t *out_vals, int64_t out_capacity, int32_t *down, int32_t *up, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    ST_CHECK_CUDA(cudaMemsetAsync(down, 0xFF, (size_t)27 * n_out * sizeof(int32_t), s));
    dim3 grid((unsigned)cdiv(n, 256), 27);
    k_strided_maps<<<grid, 256, 0, s>>>((const int4 *)coords, (int)n, (int)n_out, out_keys, out_vals,
                                        (uint32_t)(out_capacity - 1), down, up);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ Morton permutation
__global__ void k_morton_keys(const int4 *__restrict__ coords, int n, uint64_t *__restrict__ keys, int32_t *__restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = __ldg(coords + i);
    keys[i] = morton_key(c.x, c.y, c.z, c.w);
    idx[i] = i;
}

extern "C" size_t st_morton_workspace_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)n);
    return align_up(b) + 2 * align_up(n * 8) + align_up(n * 4) + 1024;
}

// perm[k] = row of the k-th voxel in (batch, Z-order) order; stable for duplicate coordinates.
extern "C" int st_morton_perm(const int32_t *coords, int64_t n, int32_t *perm, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    Carver cv(workspace, workspace_bytes);
    uint64_t *keys = cv.take<uint64_t>(n);
    uint64_t *keys2 = cv.take<uint64_t>(n);
    int32_t *idx = cv.take<int32_t>(n);
    size_t cb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cb, keys, keys2, idx, perm, (int)n);
    void *cub_ws = cv.take<char>(cb);
    if (!cv.ok()) { set_error("st_morton_perm: workspace too small"); return ST_ERR_WORKSPACE; }
    k_morton_keys<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const int4 *)coords, (int)n, keys, idx);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cb, keys, keys2, idx, perm, (int)n, 0, 64, s));
    return ST_OK;
}

// ------------------------------------------------------------------------------------ voxelise
__device__ __forceinline__ bool voxel_of(const float *__restrict__ p, const float *__restrict__ lo,
                                         const int32_t *__restrict__ grid, float vsize, int &cx, int &cy, int &cz) {
    // floor((p - lo) / vsize) in fp32, IEEE sub and div (matches the CPU loop of spconv)
    float fx = floorf(__fdiv_rn(__fsub_rn(p[0], lo[0]), vsize));
    float fy = floorf(__fdiv_rn(__fsub_rn(p[1], lo[1]), vsize));
    float fz = floorf(__fdiv_rn(__fsub_rn(p[2], lo[2]), vsize));
    if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f)) return false;
    if (!(fx < (float)grid[0] && fy < (float)grid[1] && fz < (float)grid[2])) return false;
    cx = (int)fx; cy = (int)fy; cz = (int)fz;
    return true;
}

__global__ void k_vox_insert(const float *__restrict__ pts, int n, int ld, const int32_t *__restrict__ pblock,
                             const float *__restrict__ blo, const int32_t *__restrict__ bgrid, float vsize,
                             uint64_t *keys, int32_t *vals, uint32_t mask, int32_t *__restrict__ slot_of_point) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = pblock ? __ldg(pblock + i) : 0;
    int cx, cy, cz;
    int slot = -1;
    if (voxel_of(pts + (size_t)i * ld, blo + 3 * b, bgrid + 3 * b, vsize, cx, cy, cz)) {
        uint64_t key = pack_key(b, cz, cy, cx);
        uint32_t s = hash_key(key) & mask;
        while (true) {
            unsigned long long prev = atomicCAS((unsigned long long *)(keys + s), KEY_EMPTY, key);
            if (prev == KEY_EMPTY || prev == key) { atomicMin(vals + s, i); slot = (int)s; break; }
            s = (s + 1) & mask;
        }
    }
    slot_of_point[i] = slot;
}

__global__ void k_vox_flag(const int32_t *__restrict__ slot_of_point, const int32_t *__restrict__ vals, int n,
                           int32_t *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = slot_of_point[i];
    flag[i] = (s >= 0 && vals[s] == i) ? 1 : 0;
}

__global__ void k_vox_emit(const int32_t *__restrict__ slot_of_point, const int32_t *__restrict__ vals,
                           const uint64_t *__restrict__ keys, const int32_t *__restrict__ rank, int n,
                           int32_t *__restrict__ pc_voxel_id, int32_t *__restrict__ rep_point, int4 *__restrict__ coords) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = slot_of_point[i];
    if (s < 0) { pc_voxel_id[i] = -1; return; }
    int rep = vals[s];
    int vid = rank[rep];  // exclusive scan of the representative flags = first-appearance order
    pc_voxel_id[i] = vid;
    if (rep == i) {
        rep_point[vid] = i;
        int b, z, y, x;
        unpack_key(keys[s], b, z, y, x);
        coords[vid] = make_int4(b, z, y, x);
    }
}

extern "C" size_t st_voxelize_workspace_bytes(int64_t n) {
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (int *)nullptr, (int *)nullptr, (int)n);
    int64_t cap = st_hash_capacity(n);
    return align_up(scan) + align_up(cap * 8) + align_up(cap * 4) + 3 * align_up(n * 4) + 2048;
}

extern "C" int st_voxelize(const float *points, int64_t n, int ld, const int32_t *point_block, const float *block_lo,
                           const int32_t *block_grid, int32_t n_blocks, float vsize, int32_t *pc_voxel_id,
                           int32_t *rep_point, int32_t *coords, int64_t *n_voxels_host, void *workspace,
                           size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    *n_voxels_host = 0;
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 30), "n");
    ST_REQUIRE(n_blocks < 32768, "n_blocks");
    int64_t cap = st_hash_capacity(n);
    Carver cv(workspace, workspace_bytes);
    uint64_t *keys = cv.take<uint64_t>(cap);
    int32_t *vals = cv.take<int32_t>(cap);
    int32_t *slot = cv.take<int32_t>(n);
    int32_t *flag = cv.take<int32_t>(n);
    int32_t *rank = cv.take<int32_t>(n);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, rank, (int)n);
    void *scan_ws = cv.take<char>(scan_bytes);
    if (!cv.ok()) { set_error("st_voxelize: workspace too small"); return ST_ERR_WORKSPACE; }
    ST_CHECK_CUDA(cudaMemsetAsync(keys, 0xFF, cap * sizeof(uint64_t), s));
    ST_CHECK_CUDA(cudaMemsetAsync(vals, 0x7F, cap * sizeof(int32_t), s));
    unsigned g = (unsigned)cdiv(n, 256);
    k_vox_insert<<<g, 256, 0, s>>>(points, (int)n, ld, point_block, block_lo, block_grid, vsize, keys, vals,
                                   (uint32_t)(cap - 1), slot);
    ST_CHECK_LAUNCH();
    k_vox_flag<<<g, 256, 0, s>>>(slot, vals, (int)n, flag);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_bytes, flag, rank, (int)n, s));
    k_vox_emit<<<g, 256, 0, s>>>(slot, vals, keys, rank, (int)n, pc_voxel_id, rep_point, (int4 *)coords);
    ST_CHECK_LAUNCH();
    int last_rank = 0, last_flag = 0;
    ST_CHECK_CUDA(cudaMemcpyAsync(&last_rank, rank + n - 1, 4, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaMemcpyAsync(&last_flag, flag + n - 1, 4, cudaMemcpyDeviceToHost, s));
    ST_CHECK_CUDA(cudaStreamSynchronize(s));
    *n_voxels_host = (int64_t)last_rank + last_flag;
    return ST_OK;
}

// ------------------------------------------------------------------------------------ devoxelise
// Per-voxel predictions broadcast back to EVERY input point (the reference computes
// pc_voxel_id in dataset.py:214 and then drops it).  A point can sit in several blocks (buffers overlap); it takes
// the prediction of the block whose inner half-open cube contains it (util/maths.py:135-155), which is unique.
// Points of dropped blocks / dropped by the voxeliser keep class -1, voxel -1 and a zero vector.
__global__ void k_devoxelize(const float *__restrict__ xyz, const int64_t *__restrict__ pair_point, const int32_t *__restrict__ pair_block,
                             const int32_t *__restrict__ pair_voxel, int64_t n_pairs, const float *__restrict__ centres, float half,
                             const float *__restrict__ vmedial, const int32_t *__restrict__ vclass, float *__restrict__ pmedial,
                             int32_t *__restrict__ pclass, int32_t *__restrict__ pvoxel) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int v = __ldg(pair_voxel + t);
    if (v < 0) return;
    const int64_t p = __ldg(pair_point + t);
    const int b = __ldg(pair_block + t);
    bool inside = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float c = __ldg(centres + 3 * (size_t)b + a), x = __ldg(xyz + 3 * p + a);
        inside = inside && x >= __fsub_rn(c, half) && x < __fadd_rn(c, half);
    }
    if (!inside) return;
    pmedial[3 * p] = __ldg(vmedial + 3 * (size_t)v);
    pmedial[3 * p + 1] = __ldg(vmedial + 3 * (size_t)v + 1);
    pmedial[3 * p + 2] = __ldg(vmedial + 3 * (size_t)v + 2);
    pclass[p] = __ldg(vclass + v);
    pvoxel[p] = v;
}

extern "C" int st_devoxelize(const float *xyz, int64_t n_points, const int64_t *pair_point, const int32_t *pair_block,
                             const int32_t *pair_voxel, int64_t n_pairs, const float *block_centres, float block_size,
                             const float *voxel_medial, const int32_t *voxel_class, float *point_medial, int32_t *point_class,
                             int32_t *point_voxel, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_points == 0) return ST_OK;
    ST_CHECK_CUDA(cudaMemsetAsync(point_medial, 0, (size_t)n_points * 3 * sizeof(float), s));
    ST_CHECK_CUDA(cudaMemsetAsync(point_class, 0xFF, (size_t)n_points * sizeof(int32_t), s));
    ST_CHECK_CUDA(cudaMemsetAsync(point_voxel, 0xFF, (size_t)n_points * sizeof(int32_t), s));
    if (n_pairs == 0) return ST_OK;
    k_devoxelize<<<(unsigned)cdiv(n_pairs, 256), 256, 0, s>>>(xyz, pair_point, pair_block, pair_voxel, n_pairs, block_centres,
                                                              (float)(block_size / 2), voxel_medial, voxel_class, point_medial,
                                                              point_class, point_voxel);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// ------------------------------------------------------------------------------------ row gather
// dst[r, :] = src[idx[r], :] for rows of `words` 32-bit words (int32 index): the permutations of the path (Z-order
// rows, voxel representatives) -- torch's 16-byte vectorised gather took 259 us for 495 k rows.
__global__ void k_gather_rows(const uint32_t *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, int words, uint32_t *__restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * words) return;
    const int64_t r = t / words;
    const int w = (int)(t - r * words);
    dst[t] = __ldg(src + (int64_t)__ldg(idx + r) * words + w);
}
__global__ void k_gather_rows16(const int4 *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, int4 *__restrict__ dst) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) dst[r] = __ldg(src + __ldg(idx + r));
}
extern "C" int st_gather_rows(const void *src, const int32_t *idx, int64_t n, int row_bytes, void *dst, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    ST_REQUIRE(row_bytes > 0 && row_bytes % 4 == 0, "rows must be a whole number of 32-bit words");
    if (row_bytes == 16 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) {
        k_gather_rows16<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const int4 *)src, idx, n, (int4 *)dst);
    } else {
        const int words = row_bytes / 4;
        k_gather_rows<<<(unsigned)cdiv(n * words, 256), 256, 0, s>>>((const uint32_t *)src, idx, n, words, (uint32_t *)dst);
    }
    ST_CHECK_LAUNCH();
    return ST_OK;
}
