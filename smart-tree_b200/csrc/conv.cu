// Gather convolution (sub-manifold / strided / inverse / 1x1) with fused BN-affine, residual,
// fused ResBlock-identity 1x1 conv, ReLU and concat-slice store; fused prediction heads.
//
// fp32 FMA path: out[i,:] = act(scale * sum_k W[k] . in[map[k,i],:] + shift + res[i,:] + W2 . in2[i,:])
//   - one thread owns RT output rows x CT output channels (accumulators in registers)
//   - the [CIN x COUT] weight tile of each tap is staged in shared memory with cp.async,
//     double buffered, and read as warp-broadcast LDS.128
//   - neighbour rows are read with 128-bit loads straight from L2/L1 (every feature array of
//     the network is L2 resident on B200: <= 16 MB per tensor at 250k voxels)
//   - taps for which no lane of the warp has a neighbour are skipped (decoder: <= 8 of 27)
#include "common.cuh"

using namespace st;

struct ConvArgs {
    const float *in;
    int in_ld;
    const int32_t *map;
    int n_out;
    int ntaps;
    const float *w;
    int cin, cout;
    const float *scale, *shift;
    const float *res;
    int res_ld;
    const float *in2;
    int in2_ld;
    const float *w2;
    int cin2;
    float *out;
    int out_ld;
    int act;
    const int32_t *out_index;     // optional: launch row i is output row out_index[i] (parity-sorted inverse conv)
    const uint32_t *tile_mask;    // optional: per 128 launch rows, bit k set iff tap k occurs in the tile
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// FFMA2 (fma.rn.f32x2, sm_100): two fp32 FMAs per lane and issue slot; each component rounds exactly like fmaf.
// Written as PTX on 64-bit register pairs: the compiler otherwise splits most __ffma2_rn calls back into scalar FFMAs
// when one operand is a broadcast.
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

template <int CIN, int COUT>
struct ConvCfg {
    static constexpr int THREADS = 256;
    static constexpr int CT = COUT >= 16 ? 16 : COUT;   // output channels per thread
    static constexpr int TPR = COUT / CT;               // threads per output row
    static constexpr int RT = 2;                        // rows per thread
    static constexpr int ROW_SLOTS = THREADS / TPR;
    static constexpr int ROWS = ROW_SLOTS * RT;         // rows per CTA
    static constexpr int TAP_FLOATS = CIN * COUT;
    static constexpr int TPS_RAW = 4096 / TAP_FLOATS;   // taps per 16 KB stage
    static constexpr int TPS = TPS_RAW < 1 ? 1 : (TPS_RAW > 27 ? 27 : TPS_RAW);
    static constexpr int STAGE_FLOATS = TPS * TAP_FLOATS;
    static constexpr int SMEM_BYTES = 2 * STAGE_FLOATS * 4;
};

// ---- epilogue shared by the FMA kernels: affine, residual, fused identity 1x1 conv, activation, slice store
template <int RT, int CT, int COUT>
__device__ __forceinline__ void conv_epilogue(const ConvArgs &a, const int (&rows)[RT], const unsigned long long (&acc)[RT][CT / 2], int co0) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        if (rows[r] >= a.n_out) continue;
        const int row = a.out_index ? __ldg(a.out_index + rows[r]) : rows[r];      // output row (residual / in2 follow it)
        float v[CT];
#pragma unroll
        for (int c = 0; c < CT / 2; ++c) f2_unpack(acc[r][c], v[2 * c], v[2 * c + 1]);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            float sc = a.scale ? __ldg(a.scale + co0 + c) : 1.f;
            float sh = a.shift ? __ldg(a.shift + co0 + c) : 0.f;
            v[c] = fmaf(v[c], sc, sh);
        }
        if (a.res) {
#pragma unroll
            for (int c4 = 0; c4 < CT / 4; ++c4) {
                float4 rr = __ldg((const float4 *)(a.res + (size_t)row * a.res_ld + co0) + c4);
                v[c4 * 4 + 0] += rr.x; v[c4 * 4 + 1] += rr.y; v[c4 * 4 + 2] += rr.z; v[c4 * 4 + 3] += rr.w;
            }
        }
        if (a.in2) {
            float e[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) e[c] = 0.f;
            const float *xr = a.in2 + (size_t)row * a.in2_ld;
            for (int ci = 0; ci < a.cin2; ++ci) {
                float xv = __ldg(xr + ci);
#pragma unroll
                for (int c4 = 0; c4 < CT / 4; ++c4) {
                    float4 w4 = __ldg((const float4 *)(a.w2 + (size_t)ci * COUT + co0) + c4);
                    e[c4 * 4 + 0] = fmaf(xv, w4.x, e[c4 * 4 + 0]);
                    e[c4 * 4 + 1] = fmaf(xv, w4.y, e[c4 * 4 + 1]);
                    e[c4 * 4 + 2] = fmaf(xv, w4.z, e[c4 * 4 + 2]);
                    e[c4 * 4 + 3] = fmaf(xv, w4.w, e[c4 * 4 + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] += e[c];
        }
        if (a.act & ST_ACT_RELU) {
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = fmaxf(v[c], 0.f);
        }
        float4 *o = (float4 *)(a.out + (size_t)row * a.out_ld + co0);
#pragma unroll
        for (int c4 = 0; c4 < CT / 4; ++c4) o[c4] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
    }
}

// One 32-byte row per lane and instruction (LDG.256, sm_100): a gathered 8-channel row is one sector, so the L1 data
// pipe serves it in one pass instead of two half-used ones.
__device__ __forceinline__ void ldg256(const float *p, float4 &a, float4 &b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

__device__ __align__(32) float g_zero_row[16];      // all zeros: where the gathers of absent neighbours point

// ---- narrow layers (8 or 16 input channels: level 0 and the encoder / decoder between levels 0 and 1) -----------------
// ncu on the round-1 kernel (8 -> 8, level 0): L1 data-pipe wavefronts 65-80 % of peak, FMA pipe 37 %, DRAM 11 % -- these
// layers are bound by what goes through the L1 data pipe: the gathered rows (27 x 32 B per voxel) and the per-tap weight
// tile every warp re-reads from shared memory.  Hence
//   * a thread owns FOUR output rows x 8 output channels: the 16 broadcast LDS.128 of an 8 x 8 weight tile serve 128 rows
//     of a warp instead of 64;
//   * a gathered row is ONE 32-byte load per lane (LDG.256): one sector, one pass of the data pipe instead of two
//     half-used ones;
//   * 16 input channels are two "virtual taps" of 8 (same map entry, second half of the row, second half of the weight
//     tile); 16 / 32 output channels are 2 / 4 threads per row, each with its own 8 channels -- lanes of a row group issue
//     the same gather address (one sector) and the weight LDS of the groups multicast from different banks;
//   * map entries run two virtual taps ahead and feature rows one ahead of the arithmetic (registers), FFMA2 arithmetic.
template <int CIN, int COUT>
struct NarrowCfg {
    static constexpr int THREADS = 128;
    static constexpr int H = CIN / 8;                   // virtual taps per tap
    static constexpr int CT = 8;                        // output channels per thread
    static constexpr int TPR = COUT / 8;                // threads per output row
    static constexpr int RT = 4;                        // rows per thread
    static constexpr int ROW_SLOTS = THREADS / TPR;
    static constexpr int ROWS = ROW_SLOTS * RT;         // rows per CTA: 512 / 256 / 128
    static constexpr int TAP_FLOATS = CIN * COUT;
    static constexpr int TPS_RAW = 4096 / TAP_FLOATS;   // taps per 16 KB stage
    static constexpr int TPS = TPS_RAW < 1 ? 1 : (TPS_RAW > 27 ? 27 : TPS_RAW);
    static constexpr int STAGE_FLOATS = TPS * TAP_FLOATS;
    static constexpr int SMEM_BYTES = 2 * STAGE_FLOATS * 4;
};

template <int CIN, int COUT, bool V8>
__global__ void __launch_bounds__(128, 4) k_conv_narrow(ConvArgs a) {
    using C = NarrowCfg<CIN, COUT>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int cg = tid % C::TPR;
    const int slot = tid / C::TPR;
    const int row0 = blockIdx.x * C::ROWS + slot;
    int rows[C::RT];
#pragma unroll
    for (int r = 0; r < C::RT; ++r) rows[r] = row0 + r * C::ROW_SLOTS;
    unsigned long long acc[C::RT][4];
#pragma unroll
    for (int r = 0; r < C::RT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0ull;

    // taps this CTA has to visit (see k_conv_fma)
    __shared__ int s_taps[32];
    __shared__ int s_ntaps;
    if (tid < 32) {
        unsigned m = a.ntaps >= 32 ? 0xffffffffu : ((1u << a.ntaps) - 1u);
        if (a.tile_mask) {
            unsigned mm = 0;
            const int tile0 = blockIdx.x * (C::ROWS / 128), ntile = (a.n_out + 127) / 128;
            for (int t = 0; t < C::ROWS / 128; ++t)
                if (tile0 + t < ntile) mm |= __ldg(a.tile_mask + tile0 + t);
            m &= mm;
        }
        if (m & (1u << tid)) s_taps[__popc(m & ((1u << tid) - 1u))] = tid;
        if (tid == 0) s_ntaps = __popc(m);
    }
    __syncthreads();
    const int nlist = s_ntaps;

    const int nstages = (a.ntaps + C::TPS - 1) / C::TPS;
    auto issue = [&](int s) {
        int t0 = s * C::TPS;
        int nt = min(C::TPS, a.ntaps - t0);
        const float *src = a.w + (size_t)t0 * C::TAP_FLOATS;
        float *dst = smem + (s & 1) * C::STAGE_FLOATS;
        for (int i = tid * 4; i < nt * C::TAP_FLOATS; i += C::THREADS * 4) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    issue(0);
    int li = 0;
    for (int s = 0; s < nstages; ++s) {
        if (s + 1 < nstages) { issue(s + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float *ws = smem + (s & 1) * C::STAGE_FLOATS;
        const int t0 = s * C::TPS;
        const int nt = min(C::TPS, a.ntaps - t0);
        int lend = li;
        while (lend < nlist && s_taps[lend] < t0 + nt) ++lend;
        // virtual taps vi = H * (list position) + (half of the input row), for list positions [li, lend)
        const int v0 = li * C::H, v1 = lend * C::H;
        // Entries of the gather map for virtual tap vi (the tap offset k * n_out is uniform: one 64-bit multiply per tap).
        auto map_at = [&](int vi, int (&j)[C::RT]) {
            const int k = vi < v1 ? s_taps[vi / C::H] : -1;
            const int32_t *mk = a.map ? a.map + (size_t)max(k, 0) * a.n_out : nullptr;
#pragma unroll
            for (int r = 0; r < C::RT; ++r) {
                j[r] = -1;
                if (k >= 0 && rows[r] < a.n_out) j[r] = mk ? __ldg(mk + rows[r]) : rows[r];
            }
        };
        // Rows of virtual tap vi: an absent neighbour reads the all-zero row instead of being predicated off (no register
        // zeroing, no predicate per load; the lanes without a neighbour all hit the same L1 sector).  Returns "any present".
        auto rows_at = [&](int vi, const int (&j)[C::RT], float4 (&x)[C::RT][2]) -> bool {
            const int half = C::H == 1 ? 0 : vi % C::H;
            bool any = false;
#pragma unroll
            for (int r = 0; r < C::RT; ++r) {
                any |= j[r] >= 0;
                const float *p = j[r] >= 0 ? a.in + (size_t)j[r] * a.in_ld + 8 * half : g_zero_row;
                if (V8) ldg256(p, x[r][0], x[r][1]);
                else { x[r][0] = __ldg((const float4 *)p); x[r][1] = __ldg((const float4 *)p + 1); }
            }
            return any;
        };
        auto compute = [&](int vi, const float4 (&x)[C::RT][2]) {
            const int k = s_taps[vi / C::H];
            const int half = C::H == 1 ? 0 : vi % C::H;
            const float *wt = ws + (k - t0) * C::TAP_FLOATS + half * 8 * COUT + cg * 8;
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const float4 wa = *(const float4 *)(wt + ci * COUT), wb = *(const float4 *)(wt + ci * COUT + 4);
                const unsigned long long w01 = f2_pack(wa.x, wa.y), w23 = f2_pack(wa.z, wa.w), w45 = f2_pack(wb.x, wb.y), w67 = f2_pack(wb.z, wb.w);
#pragma unroll
                for (int r = 0; r < C::RT; ++r) {
                    const float4 xq = x[r][ci >> 2];
                    const float xv = (ci & 3) == 0 ? xq.x : (ci & 3) == 1 ? xq.y : (ci & 3) == 2 ? xq.z : xq.w;
                    const unsigned long long xx = f2_pack(xv, xv);
                    acc[r][0] = ffma2(xx, w01, acc[r][0]);
                    acc[r][1] = ffma2(xx, w23, acc[r][1]);
                    acc[r][2] = ffma2(xx, w45, acc[r][2]);
                    acc[r][3] = ffma2(xx, w67, acc[r][3]);
                }
            }
        };
        // software pipeline, unrolled by two so that the two row buffers swap roles instead of being copied:
        // map entries run two virtual taps ahead of the arithmetic, feature rows one
        int ja[C::RT], jb[C::RT];
        float4 xa[C::RT][2], xb[C::RT][2];
        map_at(v0, ja);
        map_at(v0 + 1, jb);
        bool any_a = rows_at(v0, ja, xa), any_b = false;
        for (int vi = v0; vi < v1; vi += 2) {
            any_b = rows_at(vi + 1, jb, xb);           // rows of vi + 1 (entries loaded one step ago)
            map_at(vi + 2, ja);                        // entries of vi + 2
            if (__any_sync(0xffffffffu, any_a)) compute(vi, xa);
            if (vi + 1 >= v1) break;
            any_a = rows_at(vi + 2, ja, xa);           // rows of vi + 2
            map_at(vi + 3, jb);                        // entries of vi + 3
            if (__any_sync(0xffffffffu, any_b)) compute(vi + 1, xb);
        }
        li = lend;
        __syncthreads();
    }
    conv_epilogue<C::RT, 8, COUT>(a, rows, acc, cg * 8);
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(256, (CIN <= 8 && COUT <= 8 ? 4 : 1)) k_conv_fma(ConvArgs a) {
    using C = ConvCfg<CIN, COUT>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int cg = tid % C::TPR;
    const int slot = tid / C::TPR;
    const int row0 = blockIdx.x * C::ROWS + slot;
    int rows[C::RT];
#pragma unroll
    for (int r = 0; r < C::RT; ++r) rows[r] = row0 + r * C::ROW_SLOTS;

    // accumulators as channel pairs (FFMA2)
    unsigned long long acc[C::RT][C::CT / 2];
#pragma unroll
    for (int r = 0; r < C::RT; ++r)
#pragma unroll
        for (int c = 0; c < C::CT / 2; ++c) acc[r][c] = 0ull;

    // taps this CTA has to visit: all of them, or (inverse conv on parity-sorted rows) the <= 8 of 27 that occur in its
    // 128-row tiles -- the other taps' map entries are never read
    __shared__ int s_taps[32];
    __shared__ int s_ntaps;
    if (tid < 32) {
        unsigned m = a.ntaps >= 32 ? 0xffffffffu : ((1u << a.ntaps) - 1u);
        if (a.tile_mask) {
            unsigned mm = 0;
            const int tile0 = blockIdx.x * (C::ROWS / 128), ntile = (a.n_out + 127) / 128;
            for (int t = 0; t < C::ROWS / 128; ++t)
                if (tile0 + t < ntile) mm |= __ldg(a.tile_mask + tile0 + t);
            m &= mm;
        }
        if (m & (1u << tid)) s_taps[__popc(m & ((1u << tid) - 1u))] = tid;
        if (tid == 0) s_ntaps = __popc(m);
    }
    __syncthreads();
    const int nlist = s_ntaps;

    const int nstages = (a.ntaps + C::TPS - 1) / C::TPS;
    auto issue = [&](int s) {
        int t0 = s * C::TPS;
        int nt = min(C::TPS, a.ntaps - t0);
        const float *src = a.w + (size_t)t0 * C::TAP_FLOATS;
        float *dst = smem + (s & 1) * C::STAGE_FLOATS;
        for (int i = tid * 4; i < nt * C::TAP_FLOATS; i += C::THREADS * 4) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    issue(0);
    int li = 0;                     // position in the tap list (ascending taps: stages consume it in order)
    for (int s = 0; s < nstages; ++s) {
        if (s + 1 < nstages) { issue(s + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float *ws = smem + (s & 1) * C::STAGE_FLOATS;
        const int t0 = s * C::TPS;
        const int nt = min(C::TPS, a.ntaps - t0);
        int lend = li;
        while (lend < nlist && s_taps[lend] < t0 + nt) ++lend;      // list entries [li, lend) belong to this stage
        // Narrow layers (CIN <= 16: the 8 -> 8 convs of level 0) are bound by the dependent load chain of a tap
        // (gather-map entry, cold from HBM -> feature row -> FMAs) rather than by the FMAs: the map entries run two taps
        // and the feature rows one tap ahead of the arithmetic, in registers.
        constexpr bool PIPE = CIN <= 16;
        constexpr int XV = CIN / 4;
        auto map_at = [&](int i, int (&j)[C::RT]) {
            const int k = i < lend ? s_taps[i] : -1;
#pragma unroll
            for (int r = 0; r < C::RT; ++r) {
                j[r] = -1;
                if (k >= 0 && rows[r] < a.n_out) j[r] = a.map ? __ldg(a.map + (size_t)k * a.n_out + rows[r]) : rows[r];
            }
        };
        auto rows_at = [&](const int (&j)[C::RT], float4 (&x)[C::RT][PIPE ? XV : 1]) {
#pragma unroll
            for (int r = 0; r < C::RT; ++r)
#pragma unroll
                for (int q = 0; q < (PIPE ? XV : 1); ++q)
                    x[r][q] = j[r] >= 0 ? __ldg((const float4 *)(a.in + (size_t)j[r] * a.in_ld) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        int jn1[C::RT], jn2[C::RT];
        float4 xn[C::RT][PIPE ? XV : 1];
        if (PIPE) {
            int j0[C::RT];
            map_at(li, j0);
            map_at(li + 1, jn1);
            rows_at(j0, xn);
#pragma unroll
            for (int r = 0; r < C::RT; ++r) jn2[r] = j0[r];          // jn2 = entries of the tap whose rows sit in xn
        }
        for (int i = li; i < lend; ++i) {
            const int k = s_taps[i];
            int j[C::RT];
            float4 xc[C::RT][PIPE ? XV : 1];
            bool any = false;
            if (PIPE) {
#pragma unroll
                for (int r = 0; r < C::RT; ++r) {
                    j[r] = jn2[r];
                    any |= j[r] >= 0;
#pragma unroll
                    for (int q = 0; q < XV; ++q) xc[r][q] = xn[r][q];
                    jn2[r] = jn1[r];
                }
                rows_at(jn1, xn);                     // rows of the next listed tap (its entries were loaded a tap ago)
                map_at(i + 2, jn1);                   // entries of the tap after that
            } else {
#pragma unroll
                for (int r = 0; r < C::RT; ++r) {
                    j[r] = -1;
                    if (rows[r] < a.n_out) j[r] = a.map ? __ldg(a.map + (size_t)k * a.n_out + rows[r]) : rows[r];
                    any |= j[r] >= 0;
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            const float *wt = ws + (k - t0) * C::TAP_FLOATS + cg * C::CT;
#pragma unroll
            for (int ci4 = 0; ci4 < CIN / 4; ++ci4) {
                float4 x[C::RT];
#pragma unroll
                for (int r = 0; r < C::RT; ++r) {
                    if (PIPE) x[r] = xc[r][ci4];
                    else x[r] = j[r] >= 0 ? __ldg((const float4 *)(a.in + (size_t)j[r] * a.in_ld) + ci4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int co4 = 0; co4 < C::CT / 4; ++co4) {
                        float4 w4 = *(const float4 *)(wt + (ci4 * 4 + c) * COUT + co4 * 4);
                        const unsigned long long w01 = f2_pack(w4.x, w4.y), w23 = f2_pack(w4.z, w4.w);
#pragma unroll
                        for (int r = 0; r < C::RT; ++r) {
                            const float xv = c == 0 ? x[r].x : c == 1 ? x[r].y : c == 2 ? x[r].z : x[r].w;
                            const unsigned long long xx = f2_pack(xv, xv);
                            acc[r][co4 * 2 + 0] = ffma2(xx, w01, acc[r][co4 * 2 + 0]);
                            acc[r][co4 * 2 + 1] = ffma2(xx, w23, acc[r][co4 * 2 + 1]);
                        }
                    }
                }
            }
        }
        li = lend;
        __syncthreads();
    }

    conv_epilogue<C::RT, C::CT, COUT>(a, rows, acc, cg * C::CT);
}

// Generic fallback for channel counts outside the tuned set (e.g. the 3->8 stem):
// one thread per (row, output channel).
__global__ void k_conv_generic(ConvArgs a) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.n_out * a.cout) return;
    int row = (int)(idx / a.cout), co = (int)(idx % a.cout);
    float acc = 0.f;
    for (int k = 0; k < a.ntaps; ++k) {
        int j = a.map ? __ldg(a.map + (size_t)k * a.n_out + row) : row;
        if (j < 0) continue;
        const float *x = a.in + (size_t)j * a.in_ld;
        const float *w = a.w + (size_t)k * a.cin * a.cout + co;
        for (int ci = 0; ci < a.cin; ++ci) acc = fmaf(__ldg(x + ci), __ldg(w + (size_t)ci * a.cout), acc);
    }
    float v = fmaf(acc, a.scale ? a.scale[co] : 1.f, a.shift ? a.shift[co] : 0.f);
    if (a.res) v += a.res[(size_t)row * a.res_ld + co];
    if (a.in2) {
        float e = 0.f;
        for (int ci = 0; ci < a.cin2; ++ci) e = fmaf(a.in2[(size_t)row * a.in2_ld + ci], a.w2[(size_t)ci * a.cout + co], e);
        v += e;
    }
    if (a.act & ST_ACT_RELU) v = fmaxf(v, 0.f);
    a.out[(size_t)(a.out_index ? a.out_index[row] : row) * a.out_ld + co] = v;
}

template <int CIN, int COUT>
static int launch_fma(const ConvArgs &a, cudaStream_t s) {
    using C = ConvCfg<CIN, COUT>;
    static bool attr_set = false;
    if (!attr_set && C::SMEM_BYTES > 48 * 1024) {
        ST_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fma<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    unsigned grid = (unsigned)cdiv(a.n_out, C::ROWS);
    k_conv_fma<CIN, COUT><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(a);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

template <int CIN, int COUT>
static int launch_narrow(const ConvArgs &a, cudaStream_t s) {
    using C = NarrowCfg<CIN, COUT>;
    static_assert(C::SMEM_BYTES <= 48 * 1024, "narrow conv stages must fit the default shared-memory limit");
    unsigned grid = (unsigned)cdiv(a.n_out, C::ROWS);
    // 32-byte gathers when every input row starts on a 32-byte boundary
    if (((uintptr_t)a.in & 31) == 0 && a.in_ld % 8 == 0 && !getenv("ST_CONV_NO_V8"))
        k_conv_narrow<CIN, COUT, true><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(a);
    else
        k_conv_narrow<CIN, COUT, false><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(a);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }
static int conv_dispatch(const ConvArgs &a, cudaStream_t s);

extern "C" int st_conv_gather(const float *in, int in_ld, const int32_t *map, int64_t n_out, int ntaps, const float *w,
                              int cin, int cout, const float *scale, const float *shift, const float *residual,
                              int res_ld, const float *in2, int in2_ld, const float *w2, int cin2, float *out,
                              int out_ld, int act, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_out == 0) return ST_OK;
    ST_REQUIRE(n_out < (1ll << 31), "n_out");
    ST_REQUIRE(ntaps >= 1 && (map != nullptr || ntaps == 1), "identity map requires ntaps == 1");
    ST_REQUIRE(!in2 || w2, "in2 needs w2");
    ST_REQUIRE(ntaps <= 32, "at most 32 taps");
    ConvArgs a{in, in_ld, map, (int)n_out, ntaps, w, cin, cout, scale, shift, residual, res_ld, in2, in2_ld, w2, cin2, out, out_ld, act,
               nullptr, nullptr};
    return conv_dispatch(a, s);
}

static int conv_dispatch(const ConvArgs &a, cudaStream_t s) {
    const float *in = a.in, *w = a.w, *residual = a.res, *w2 = a.w2;
    float *out = a.out;
    const int in_ld = a.in_ld, out_ld = a.out_ld, res_ld = a.res_ld, cin = a.cin, cout = a.cout;
    const int64_t n_out = a.n_out;
    bool fast = aligned16(in) && aligned16(out) && aligned16(w) && (in_ld % 4 == 0) && (out_ld % 4 == 0) &&
                (!residual || (aligned16(residual) && res_ld % 4 == 0)) && (!w2 || aligned16(w2));
    if (fast) {
        if (!getenv("ST_CONV_NO_NARROW")) {
#define ST_NARROW(CI, CO) if (cin == CI && cout == CO) return launch_narrow<CI, CO>(a, s);
            ST_NARROW(8, 8) ST_NARROW(8, 16) ST_NARROW(16, 8) ST_NARROW(16, 16) ST_NARROW(16, 32)
#undef ST_NARROW
        }
#define ST_CASE(CI, CO) if (cin == CI && cout == CO) return launch_fma<CI, CO>(a, s);
        ST_CASE(8, 8) ST_CASE(8, 16) ST_CASE(16, 8) ST_CASE(16, 16) ST_CASE(16, 32) ST_CASE(32, 16)
        ST_CASE(32, 32) ST_CASE(32, 64) ST_CASE(64, 32) ST_CASE(64, 64) ST_CASE(8, 4) ST_CASE(4, 4)
#undef ST_CASE
    }
    int64_t total = n_out * cout;
    k_conv_generic<<<(unsigned)cdiv(total, 256), 256, 0, s>>>(a);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// Stem: the 1x1 input conv (3 -> 8 in the shipped checkpoints) + BN + ReLU, fused with the row permutation that puts the
// network's rows into Z-order: out[i,:] = act(scale * (W . in[row_index[i], :cin]) + shift).  One thread per row, weights in
// registers via shared memory, two 16-byte stores; `in` may be a column slice of a wider array (in_ld).
template <int COUT>
__global__ void __launch_bounds__(256) k_stem(const float *__restrict__ in, int in_ld, const int32_t *__restrict__ row_index, int n,
                                              const float *__restrict__ w, int cin, const float *__restrict__ scale,
                                              const float *__restrict__ shift, float *__restrict__ out, int out_ld, int act) {
    __shared__ float sw[8 * COUT + 2 * COUT];
    for (int i = threadIdx.x; i < cin * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
        sw[8 * COUT + i] = scale ? scale[i] : 1.f;
        sw[9 * COUT + i] = shift ? shift[i] : 0.f;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *x = in + (size_t)(row_index ? __ldg(row_index + i) : i) * in_ld;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    for (int ci = 0; ci < cin; ++ci) {
        const float xv = __ldg(x + ci);
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = fmaf(xv, sw[ci * COUT + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
        acc[c] = fmaf(acc[c], sw[8 * COUT + c], sw[9 * COUT + c]);
        if (act & ST_ACT_RELU) acc[c] = fmaxf(acc[c], 0.f);
    }
    float4 *o = (float4 *)(out + (size_t)i * out_ld);
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) o[c4] = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
}

extern "C" int st_stem_conv(const float *in, int in_ld, const int32_t *row_index, int64_t n, const float *w, int cin, int cout,
                            const float *scale, const float *shift, float *out, int out_ld, int act, void *stream) {
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31), "n");
    if (!(cin >= 1 && cin <= 8 && (cout == 8 || cout == 16) && aligned16(out) && out_ld % 4 == 0)) {
        set_error("st_stem_conv: supports cin <= 8, cout in {8, 16}, 16-byte aligned output rows");
        return ST_ERR_UNSUPPORTED;
    }
    const unsigned g = (unsigned)cdiv(n, 256);
    if (cout == 8) k_stem<8><<<g, 256, 0, (cudaStream_t)stream>>>(in, in_ld, row_index, (int)n, w, cin, scale, shift, out, out_ld, act);
    else k_stem<16><<<g, 256, 0, (cudaStream_t)stream>>>(in, in_ld, row_index, (int)n, w, cin, scale, shift, out, out_ld, act);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// Inverse (decoder) conv on parity-sorted rows through the FMA kernel: FMA counterpart of st_conv_gather_tc_inv for the
// narrow layers (16 -> 8 at level 0), where a fine voxel has 3.4 of 27 taps on average and the tensor-core tile switch
// costs more than the arithmetic.  The CTA visits only the taps its tiles' masks name.
extern "C" int st_conv_gather_inv(const float *in, int in_ld, const int32_t *up_sorted, const int32_t *row_index,
                                  const uint32_t *tile_mask, int64_t n_out, int ntaps, const float *w, int cin, int cout,
                                  const float *scale, const float *shift, float *out, int out_ld, int act, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_out == 0) return ST_OK;
    ST_REQUIRE(n_out < (1ll << 31), "n_out");
    ST_REQUIRE(ntaps >= 1 && ntaps <= 32 && up_sorted && row_index, "inverse conv needs up_sorted / row_index and at most 32 taps");
    ConvArgs a{in, in_ld, up_sorted, (int)n_out, ntaps, w, cin, cout, scale, shift, nullptr, 0, nullptr, 0, nullptr, 0, out, out_ld, act,
               row_index, tile_mask};
    return conv_dispatch(a, s);
}

// ------------------------------------------------------------------------------------ fused heads
// params (floats): for head h in {radius(k=1), direction(k=3), class(k=2)}:
//   W1[8][8] (ci-major), s1[8], b1[8], W2[8][4], s2[4], b2[4], W3[4][k], b3[k]
constexpr int HEAD_FIXED = 64 + 8 + 8 + 32 + 4 + 4;
__host__ __device__ constexpr int head_size(int k) { return HEAD_FIXED + 5 * k; }
constexpr int HEADS_TOTAL = head_size(1) + head_size(3) + head_size(2);

template <int K>
__device__ __forceinline__ void run_head(const float *__restrict__ p, const float x[8], float out[K]) {
    float h1[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(x[i], p[i * 8 + o], acc);
        h1[o] = fmaxf(fmaf(acc, p[64 + o], p[72 + o]), 0.f);
    }
    const float *q = p + 80;
    float h2[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(h1[i], q[i * 4 + o], acc);
        h2[o] = fmaxf(fmaf(acc, q[32 + o], q[36 + o]), 0.f);
    }
    const float *t = q + 40;
#pragma unroll
    for (int o = 0; o < K; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc = fmaf(h2[i], t[i * K + o], acc);
        out[o] = acc + t[4 * K + o];
    }
}

__global__ void __launch_bounds__(256) k_heads(const float *__restrict__ in, int in_ld, int n, const float *__restrict__ params,
                                               float *__restrict__ radius, float *__restrict__ direction,
                                               float *__restrict__ logits, float *__restrict__ medial, int32_t *__restrict__ cls,
                                               const int32_t *__restrict__ out_index) {
    __shared__ float sp[HEADS_TOTAL];
    for (int i = threadIdx.x; i < HEADS_TOTAL; i += blockDim.x) sp[i] = params[i];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[8];
    const float4 *xr = (const float4 *)(in + (size_t)i * in_ld);
    float4 a = __ldg(xr), b = __ldg(xr + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    float r[1], d[3], c[2];
    run_head<1>(sp, x, r);
    run_head<3>(sp + head_size(1), x, d);
    run_head<2>(sp + head_size(1) + head_size(3), x, c);
    if (out_index) i = __ldg(out_index + i);      // un-permute: internal (Z-order) row -> caller's row
    // F.normalize(p=2, dim=1, eps=1e-12)
    float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    float den = fmaxf(nrm, 1e-12f);
    d[0] = __fdiv_rn(d[0], den); d[1] = __fdiv_rn(d[1], den); d[2] = __fdiv_rn(d[2], den);
    if (radius) radius[i] = r[0];
    if (direction) { direction[3 * (size_t)i] = d[0]; direction[3 * (size_t)i + 1] = d[1]; direction[3 * (size_t)i + 2] = d[2]; }
    if (logits) { logits[2 * (size_t)i] = c[0]; logits[2 * (size_t)i + 1] = c[1]; }
    if (medial) {
        float e = expf(r[0]);
        medial[3 * (size_t)i] = e * d[0]; medial[3 * (size_t)i + 1] = e * d[1]; medial[3 * (size_t)i + 2] = e * d[2];
    }
    if (cls) cls[i] = c[1] > c[0] ? 1 : 0;
}

extern "C" int st_heads_fused(const float *in, int in_ld, int64_t n, const float *params, const int32_t *out_index, float *radius,
                              float *direction, float *class_logits, float *medial_vector, int32_t *class_l, void *stream) {
    if (n == 0) return ST_OK;
    ST_REQUIRE(aligned16(in) && in_ld % 4 == 0, "heads input must be 16-byte aligned rows");
    k_heads<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(in, in_ld, (int)n, params, radius, direction,
                                                                        class_logits, medial_vector, class_l, out_index);
    ST_CHECK_LAUNCH();
    return ST_OK;
}
