// Sub-manifold 3x3x3 convolution of the NARROW level (8 or 16 input channels, 8 output channels: level 0 of the UNet) on
// 4x4x4 BRICKS -- no gather map.
//
// Why: with 8 channels a voxel row is 32 bytes.  The gather-map kernels (conv.cu) read 27 scattered 32-byte rows per voxel
// through the L1 (one wavefront per distinct line: the bound that ncu showed, L1 wavefronts 65-80 % of peak) plus a
// 108-byte map column per voxel from HBM -- more than the 64 bytes of features the voxel reads and writes.  Rows are kept in
// (batch, Z-order), so the voxels of an aligned 4x4x4 brick are CONTIGUOUS rows, and all 27 neighbours of its voxels lie
// in the 6x6x6 window around it.  One WARP per brick:
//   * a 216-byte cell table of the window in shared memory (cell -> local row, 0 = absent = an all-zero row): filled from
//     a 1-byte cell code per own row and a per-brick HALO LIST (source row + cell code of the window's occupied cells
//     outside the brick), both built once per level (st_brick_plan_build: 152 hash probes per brick instead of 27 per
//     voxel) and shared by every conv of the level: ~8 bytes of index per voxel and launch instead of 108;
//   * the window's rows (own rows: one contiguous range; halo rows: gathered) are staged once with cp.async; the 27 taps
//     then read neighbours with LDS.128 from shared memory -- absent neighbours read the zero row, no predication;
//   * lanes = (row slot, tap group): a slot owns two rows, the T = 1/2/4/8 lanes of a slot split the 27 taps between them
//     (so that bricks with few voxels still fill the warp) and add their partial sums with shuffles; FFMA2 arithmetic with
//     the per-tap weight tiles in shared memory (tile stride padded: the T tiles of one step sit in different banks);
//   * fused epilogue as in conv.cu: BN affine, residual, ResBlock identity 1x1 conv (in2 . w2), ReLU, column-slice store.
// 16 input channels run as two passes of 8 over the same staging buffer.
#include <cub/cub.cuh>

#include "common.cuh"

using namespace st;

namespace {

constexpr int BR_WARPS = 8;                 // warps (= bricks in flight) per CTA
constexpr int BR_WIN = 216;                 // cells of the 6x6x6 window
constexpr int BR_ROWS = BR_WIN + 1;         // staged rows: all window cells + the zero row
constexpr int BR_TAB_BYTES = 224;
constexpr int BR_FEAT_FLOATS = BR_ROWS * 8;
constexpr int BR_WARP_BYTES = BR_TAB_BYTES + BR_FEAT_FLOATS * 4;      // 7168

struct BrickPlan {                           // views into the caller's plan buffer (layout fixed by n)
    int32_t *hdr;                            // [0] bricks, [1] halo entries, [2] status (bit 0: rows not in (batch, Z-order))
    int32_t *brick_row;                      // [n + 1] first row of each brick (+ n at the end)
    int32_t *hoff, *hcnt;                    // [n] halo list segment of each brick
    int32_t *hsrc;                           // [7 n] source row of a halo entry
    uint8_t *code;                           // [n] window cell of each row
    uint8_t *hcode;                          // [7 n] window cell of a halo entry
    int32_t *scan;                           // [n] scratch (brick index of each row)
};

__host__ __device__ inline size_t plan_al(size_t x) { return (x + 255) / 256 * 256; }

__host__ __device__ inline BrickPlan plan_views(void *buf, int64_t n) {
    char *p = (char *)buf;
    BrickPlan v;
    v.hdr = (int32_t *)p; p += 256;
    v.brick_row = (int32_t *)p; p += plan_al(4 * (size_t)(n + 1));
    v.hoff = (int32_t *)p; p += plan_al(4 * (size_t)n);
    v.hcnt = (int32_t *)p; p += plan_al(4 * (size_t)n);
    v.hsrc = (int32_t *)p; p += plan_al(28 * (size_t)n);
    v.code = (uint8_t *)p; p += plan_al((size_t)n);
    v.hcode = (uint8_t *)p; p += plan_al(7 * (size_t)n);
    v.scan = (int32_t *)p;
    return v;
}

// brick of a voxel in SHIFTED coordinates q = p + 1 (the Morton keys interleave q): q >> 2 per axis
__device__ __forceinline__ bool same_brick(int4 a, int4 b) {
    return a.x == b.x && ((a.y + 1) >> 2) == ((b.y + 1) >> 2) && ((a.z + 1) >> 2) == ((b.z + 1) >> 2) && ((a.w + 1) >> 2) == ((b.w + 1) >> 2);
}

__global__ void k_brick_flags(const int4 *__restrict__ coords, int n, int32_t *__restrict__ flag, uint8_t *__restrict__ code,
                              int32_t *__restrict__ hdr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(coords + i);
    int f = 1;
    if (i > 0) {
        const int4 p = __ldg(coords + i - 1);
        f = same_brick(c, p) ? 0 : 1;
        if (!(morton_key(p.x, p.y, p.z, p.w) < morton_key(c.x, c.y, c.z, c.w))) atomicOr(hdr + 2, 1);
    }
    flag[i] = f;
    code[i] = (uint8_t)((((c.y + 1) & 3) + 1) * 36 + (((c.z + 1) & 3) + 1) * 6 + (((c.w + 1) & 3) + 1));
}

// inclusive scan of the flags in `scan` -> brick_row
__global__ void k_brick_rows(const int4 *__restrict__ coords, int n, const int32_t *__restrict__ scan, int32_t *__restrict__ brick_row,
                             int32_t *__restrict__ hdr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = scan[i] - 1;
    const bool first = i == 0 || scan[i - 1] != scan[i];
    if (first) brick_row[b] = i;
    if (i == n - 1) { brick_row[b + 1] = n; hdr[0] = b + 1; }
}

// halo lists: one warp per brick probes the 152 window cells outside the brick
__global__ void __launch_bounds__(256) k_brick_halo(const int4 *__restrict__ coords, const uint64_t *__restrict__ keys, const int32_t *__restrict__ vals,
                                                    uint32_t mask, BrickPlan pl) {
    const int lane = threadIdx.x & 31;
    const int nbk = pl.hdr[0];
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nbk; b += nwarps) {
        const int4 c = __ldg(coords + pl.brick_row[b]);
        // window cell (wz, wy, wx) in [0, 6)^3 <-> coordinate base + w, base = first cell of the brick - 1 (unshifted: q0 - 2)
        const int bz = (((c.y + 1) >> 2) << 2) - 2, by = (((c.z + 1) >> 2) << 2) - 2, bx = (((c.w + 1) >> 2) << 2) - 2;
        int src[7];
        int cnt = 0;
        unsigned got = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int cell = lane + 32 * i;
            src[i] = -1;
            if (cell < BR_WIN) {
                const int wz = cell / 36, wy = (cell / 6) % 6, wx = cell % 6;
                const bool halo = wz == 0 || wz == 5 || wy == 0 || wy == 5 || wx == 0 || wx == 5;
                const int z = bz + wz, y = by + wy, x = bx + wx;
                if (halo && z >= 0 && y >= 0 && x >= 0) src[i] = hash_lookup(keys, vals, mask, pack_key(c.x, z, y, x));
            }
            if (src[i] >= 0) { ++cnt; got |= 1u << i; }
        }
        // warp-exclusive prefix of cnt
        int pre = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
        }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        pre -= cnt;
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(pl.hdr + 1, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane == 0) { pl.hoff[b] = base; pl.hcnt[b] = total; }
        int o = base + pre;
#pragma unroll
        for (int i = 0; i < 7; ++i)
            if (got & (1u << i)) { pl.hsrc[o] = src[i]; pl.hcode[o] = (uint8_t)(lane + 32 * i); ++o; }
    }
}

struct BrickArgs {
    const float *in;
    int in_ld;
    BrickPlan pl;
    int n;
    const float *w;              // [27, cin, 8]
    const float *scale, *shift;
    const float *res;
    int res_ld;
    const float *in2;
    int in2_ld;
    const float *w2;             // [cin2, 8]
    int cin2;
    float *out;
    int out_ld;
    int act;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int CIN>
__global__ void __launch_bounds__(BR_WARPS * 32, 3) k_conv_brick(BrickArgs a) {
    constexpr int NH = CIN / 8;                       // passes of 8 input channels
    constexpr int TILE = CIN * 8 + 4;                 // floats per tap tile (padded: consecutive tiles start 4 banks apart)
    extern __shared__ __align__(16) unsigned char smem[];
    float *sw = reinterpret_cast<float *>(smem);                                   // [27][TILE]
    float *sw2 = sw + 27 * TILE;                                                   // [16][8]
    float *saff = sw2 + 128;                                                       // scale[8], shift[8]
    unsigned char *warp_base = smem + (27 * TILE + 128 + 16) * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *tab = warp_base + warp * BR_WARP_BYTES;
    float *feat = reinterpret_cast<float *>(tab + BR_TAB_BYTES);

    for (int i = tid; i < 27 * CIN * 8; i += BR_WARPS * 32) sw[(i / (CIN * 8)) * TILE + i % (CIN * 8)] = __ldg(a.w + i);
    if (a.in2) for (int i = tid; i < a.cin2 * 8; i += BR_WARPS * 32) sw2[i] = __ldg(a.w2 + i);
    if (tid < 8) { saff[tid] = a.scale ? __ldg(a.scale + tid) : 1.f; saff[8 + tid] = a.shift ? __ldg(a.shift + tid) : 0.f; }
    if (lane < 8) feat[lane] = 0.f;                   // row 0 of every warp's buffer: the all-zero row
    __syncthreads();

    const int nbk = a.pl.hdr[0];
    const int nwarps = gridDim.x * BR_WARPS;
    for (int b = blockIdx.x * BR_WARPS + warp; b < nbk; b += nwarps) {
        const int r0 = a.pl.brick_row[b];
        const int R = a.pl.brick_row[b + 1] - r0;     // 1 .. 64 own rows
        const int h0 = a.pl.hoff[b], H = a.pl.hcnt[b];
        // ---- cell table
        for (int i = lane; i < BR_TAB_BYTES / 4; i += 32) reinterpret_cast<uint32_t *>(tab)[i] = 0u;
        __syncwarp();
        for (int i = lane; i < R; i += 32) tab[__ldg(a.pl.code + r0 + i)] = (uint8_t)(1 + i);
        for (int i = lane; i < H; i += 32) tab[__ldg(a.pl.hcode + h0 + i)] = (uint8_t)(1 + R + i);
        // ---- thread mapping: slot = two rows, T lanes per slot split the taps
        const int S = (R + 1) >> 1;
        const int T = S > 16 ? 1 : S > 8 ? 2 : S > 4 ? 4 : 8;
        const int tg = lane & (T - 1), slot = lane / T;
        const int row[2] = {2 * slot, 2 * slot + 1};
        int cell[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) cell[r] = row[r] < R ? (int)__ldg(a.pl.code + r0 + row[r]) : -1;
        unsigned long long acc[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0ull;
#pragma unroll 1
        for (int hh = 0; hh < NH; ++hh) {
            // ---- stage the window's rows (8 channels of this pass): own rows, then halo rows
            __syncwarp();
            const float *inh = a.in + 8 * hh;
            for (int i = lane; i < 2 * R; i += 32)
                cp_async16(feat + 8 + (i >> 1) * 8 + (i & 1) * 4, inh + (size_t)(r0 + (i >> 1)) * a.in_ld + (i & 1) * 4);
            for (int i = lane; i < 2 * H; i += 32) {
                const int src = __ldg(a.pl.hsrc + h0 + (i >> 1));
                cp_async16(feat + 8 + (R + (i >> 1)) * 8 + (i & 1) * 4, inh + (size_t)src * a.in_ld + (i & 1) * 4);
            }
            cp_async_wait_all();
            __syncwarp();
            // ---- taps tg, tg + T, ...
            for (int k = tg; k < 27; k += T) {
                const int off = (k / 9 - 1) * 36 + ((k / 3) % 3 - 1) * 6 + (k % 3 - 1);
                float4 x[2][2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int li = cell[r] >= 0 ? (int)tab[cell[r] + off] : 0;
                    const float4 *p = reinterpret_cast<const float4 *>(feat + li * 8);
                    x[r][0] = p[0]; x[r][1] = p[1];
                }
                const float *wt = sw + k * TILE + hh * 64;
#pragma unroll
                for (int ci = 0; ci < 8; ++ci) {
                    const float4 wa = *reinterpret_cast<const float4 *>(wt + ci * 8), wb = *reinterpret_cast<const float4 *>(wt + ci * 8 + 4);
                    const unsigned long long w01 = f2_pack(wa.x, wa.y), w23 = f2_pack(wa.z, wa.w), w45 = f2_pack(wb.x, wb.y), w67 = f2_pack(wb.z, wb.w);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float4 xq = x[r][ci >> 2];
                        const float xv = (ci & 3) == 0 ? xq.x : (ci & 3) == 1 ? xq.y : (ci & 3) == 2 ? xq.z : xq.w;
                        const unsigned long long xx = f2_pack(xv, xv);
                        acc[r][0] = ffma2(xx, w01, acc[r][0]);
                        acc[r][1] = ffma2(xx, w23, acc[r][1]);
                        acc[r][2] = ffma2(xx, w45, acc[r][2]);
                        acc[r][3] = ffma2(xx, w67, acc[r][3]);
                    }
                }
            }
        }
        // ---- partial sums of the slot's T lanes
        float v[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) f2_unpack(acc[r][c], v[r][2 * c], v[r][2 * c + 1]);
        for (int o = 1; o < T; o <<= 1)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) v[r][c] += __shfl_xor_sync(0xffffffffu, v[r][c], o);
        // ---- epilogue: affine, residual, identity 1x1 conv (its input channels split over the T lanes), activation
        float e[2][8];
        if (a.in2) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) e[r][c] = 0.f;
                if (row[r] < R) {
                    const float *xr = a.in2 + (size_t)(r0 + row[r]) * a.in2_ld;
                    for (int ci = tg; ci < a.cin2; ci += T) {
                        const float xv = __ldg(xr + ci);
                        const float4 wa = *reinterpret_cast<const float4 *>(sw2 + ci * 8), wb = *reinterpret_cast<const float4 *>(sw2 + ci * 8 + 4);
                        e[r][0] = fmaf(xv, wa.x, e[r][0]); e[r][1] = fmaf(xv, wa.y, e[r][1]); e[r][2] = fmaf(xv, wa.z, e[r][2]); e[r][3] = fmaf(xv, wa.w, e[r][3]);
                        e[r][4] = fmaf(xv, wb.x, e[r][4]); e[r][5] = fmaf(xv, wb.y, e[r][5]); e[r][6] = fmaf(xv, wb.z, e[r][6]); e[r][7] = fmaf(xv, wb.w, e[r][7]);
                    }
                }
            }
            for (int o = 1; o < T; o <<= 1)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) e[r][c] += __shfl_xor_sync(0xffffffffu, e[r][c], o);
        }
        if (tg == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (row[r] >= R) continue;
                const size_t gr = (size_t)(r0 + row[r]);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[r][c] = fmaf(v[r][c], saff[c], saff[8 + c]);
                if (a.res) {
                    const float4 ra = __ldg(reinterpret_cast<const float4 *>(a.res + gr * a.res_ld)), rbv = __ldg(reinterpret_cast<const float4 *>(a.res + gr * a.res_ld) + 1);
                    v[r][0] += ra.x; v[r][1] += ra.y; v[r][2] += ra.z; v[r][3] += ra.w;
                    v[r][4] += rbv.x; v[r][5] += rbv.y; v[r][6] += rbv.z; v[r][7] += rbv.w;
                }
                if (a.in2) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[r][c] += e[r][c];
                }
                if (a.act & ST_ACT_RELU) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[r][c] = fmaxf(v[r][c], 0.f);
                }
                float4 *o = reinterpret_cast<float4 *>(a.out + gr * a.out_ld);
                o[0] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
                o[1] = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
            }
        }
        __syncwarp();
    }
}

template <int CIN>
int launch_brick(const BrickArgs &a, cudaStream_t s) {
    constexpr size_t smem = (27 * (CIN * 8 + 4) + 128 + 16) * 4 + (size_t)BR_WARPS * BR_WARP_BYTES;
    static bool attr = false;
    if (!attr) {
        ST_CHECK_CUDA(cudaFuncSetAttribute(k_conv_brick<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int dev = 0, sms = 0;
    ST_CHECK_CUDA(cudaGetDevice(&dev));
    ST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // resident grid (3 CTAs per SM), bricks dealt round-robin to the warps; never more warps than rows / 2
    int64_t grid = (int64_t)sms * 3;
    const int64_t cap = cdiv(a.n, 2 * BR_WARPS);
    if (grid > cap) grid = cap < 1 ? 1 : cap;
    k_conv_brick<CIN><<<(unsigned)grid, BR_WARPS * 32, smem, s>>>(a);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

}  // namespace

extern "C" size_t st_brick_plan_bytes(int64_t n) {
    if (n < 1) n = 1;
    size_t scan_tmp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
    return 256 + plan_al(4 * (size_t)(n + 1)) + 2 * plan_al(4 * (size_t)n) + plan_al(28 * (size_t)n) + plan_al((size_t)n) + plan_al(7 * (size_t)n) +
           plan_al(4 * (size_t)n) + plan_al(4 * (size_t)n) + plan_al(scan_tmp);
}

extern "C" int st_brick_plan_build(const int32_t *coords, int64_t n, const uint64_t *keys, const int32_t *vals, int64_t capacity, void *plan,
                                   size_t plan_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ST_REQUIRE(n >= 0 && n < (1ll << 31) / 8, "n");
    ST_REQUIRE(plan != nullptr && plan_bytes >= st_brick_plan_bytes(n), "plan buffer too small (st_brick_plan_bytes)");
    BrickPlan pl = plan_views(plan, n);
    ST_CHECK_CUDA(cudaMemsetAsync(pl.hdr, 0, 256, s));
    if (n == 0) return ST_OK;
    int32_t *flag = pl.scan + plan_al(4 * (size_t)n) / 4;
    void *tmp = (char *)flag + plan_al(4 * (size_t)n);
    size_t scan_tmp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
    const unsigned g = (unsigned)cdiv(n, 256);
    k_brick_flags<<<g, 256, 0, s>>>((const int4 *)coords, (int)n, flag, pl.code, pl.hdr);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceScan::InclusiveSum(tmp, scan_tmp, flag, pl.scan, (int)n, s));
    k_brick_rows<<<g, 256, 0, s>>>((const int4 *)coords, (int)n, pl.scan, pl.brick_row, pl.hdr);
    ST_CHECK_LAUNCH();
    int dev = 0, sms = 0;
    ST_CHECK_CUDA(cudaGetDevice(&dev));
    ST_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t hg = (int64_t)sms * 8;
    if (hg > cdiv(n, 8)) hg = cdiv(n, 8);
    k_brick_halo<<<(unsigned)hg, 256, 0, s>>>((const int4 *)coords, keys, vals, (uint32_t)(capacity - 1), pl);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

extern "C" int st_brick_plan_info(const void *plan, int64_t n, int32_t *info_host /* [3]: bricks, halo entries, status */) {
    BrickPlan pl = plan_views((void *)plan, n);
    ST_CHECK_CUDA(cudaMemcpy(info_host, pl.hdr, 12, cudaMemcpyDeviceToHost));
    return ST_OK;
}

extern "C" int st_conv_brick(const float *in, int in_ld, const void *plan, int64_t n, const float *w, int cin, int cout, const float *scale,
                             const float *shift, const float *residual, int res_ld, const float *in2, int in2_ld, const float *w2, int cin2,
                             float *out, int out_ld, int act, void *stream) {
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31) / 8, "n");
    auto al16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    if (!((cin == 8 || cin == 16) && cout == 8)) {
        set_error("st_conv_brick: supports 8 or 16 input channels and 8 output channels (got %d -> %d)", cin, cout);
        return ST_ERR_UNSUPPORTED;
    }
    ST_REQUIRE(al16(in) && al16(out) && in_ld % 4 == 0 && out_ld % 4 == 0 && (!residual || (al16(residual) && res_ld % 4 == 0)),
               "16-byte aligned rows");
    ST_REQUIRE(!in2 || (w2 && cin2 >= 1 && cin2 <= 16), "in2 needs w2 and at most 16 channels");
    BrickArgs a{in, in_ld, plan_views((void *)plan, n), (int)n, w, scale, shift, residual, res_ld, in2, in2_ld, w2, cin2, out, out_ld, act};
    return cin == 8 ? launch_brick<8>(a, (cudaStream_t)stream) : launch_brick<16>(a, (cudaStream_t)stream);
}
