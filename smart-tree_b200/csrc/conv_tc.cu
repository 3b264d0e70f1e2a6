// Gather convolution on the 5th-generation tensor cores (tcgen05 / TMEM), fp32-grade accuracy.
//
//   out[i,:] = act(scale * sum_k W[k] . in[map[k,i],:] + shift + res[i,:] + W2 . in2[i,:])
//
// Implicit GEMM per CTA: M = 128 output rows, N = Cout (padded to a multiple of 16), K = 27*Cin
// walked in stages of 32.  Per stage
//   * 8 producer warps gather the neighbour rows (128-bit loads; rows are kept in Z-order so most hit
//     L1), split each fp32 value into a TF32 head and an fp32 tail, and write both straight into
//     TENSOR MEMORY with tcgen05.st (thread = row = TMEM lane, K along the columns): the A operand never
//     touches shared memory -- with the 3xTF32 split an smem-resident A would be read three times per
//     k-step and made the first version of this kernel shared-memory-bandwidth bound;
//   * the weight tile (head + tail, pre-arranged by st_conv_tc_prepare) arrives by one 1-D bulk
//     async copy (cp.async.bulk -> UBLKCP) into its own deeper ring, issued several stages ahead by a
//     dedicated loader lane so that its L2 latency never sits on the MMA critical path;
//   * one elected thread issues 12 tcgen05.mma.kind::tf32 (4 k-steps x {hi*hi, lo*hi, hi*lo}: the
//     3xTF32 split keeps ~21 mantissa bits, needed for the 1e-3 end-to-end tolerance) that
//     accumulate in TMEM; tcgen05.commit releases the stage back to the producers.
// CTAs are persistent (one or two per SM) and walk over 128-row tiles, so TMEM allocation, barrier
// initialisation and pipeline fill are paid once.
// Epilogue: the producer warps read their accumulator lanes with tcgen05.ld and apply the fused
// BN affine / residual / identity 1x1 conv / ReLU, writing straight into the (possibly sliced) output.
#include <cub/cub.cuh>
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include "common.cuh"

using namespace st;

namespace {

constexpr int TC_M = 128;          // rows per CTA
constexpr int TC_KS = 32;          // K elements per stage
constexpr int TC_STAGES = 3;          // A stages, in tensor memory (64 columns each: 32 hi + 32 lo)
constexpr int TC_A_COLS = 2 * TC_KS;
constexpr int TC_MAXTAPS = 28;        // gather-map entries of a row cached in shared memory
constexpr int TC_PRODUCERS = 256;  // 8 warps: 2 threads per row
constexpr int TC_THREADS = TC_PRODUCERS + 64;   // + MMA issuer warp + weight-loader warp
constexpr int TC_MAX_BSTAGES = 8;
constexpr int A_TILE_FLOATS = TC_M * TC_KS;            // 4096 floats = 16 KB (one of hi / lo)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
// non-blocking probe (its latency overlaps with whatever is issued next); true = phase `parity` has completed
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// exactly one lane of a converged warp; the compiler then emits warp-uniform instructions (UTCHMMA, UTCBAR,
// UBLKCP) directly instead of wrapping each one in an ELECT/branch serialisation loop
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate.  The issuing thread is the serial
// bottleneck of the whole CTA, so descriptors are passed as (lo, hi) words: hi is a constant and lo is
// one integer add per MMA.
__device__ __forceinline__ void umma_tf32_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.eq.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_tf32_init(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A = 128 lanes x 8 columns (row i in lane i, k along the columns)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// 16x256b.x2: thread t = 4*g + q writes {r0,r1 | r4,r5} to lane g, columns {2q,2q+1 | 8+2q,8+2q+1} and
// {r2,r3 | r6,r7} to lane g+8, same columns (layout measured with tools/tmem_probe.cu)
__device__ __forceinline__ void tmem_st_quad(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// descriptor words: lo = (addr >> 4) | LBO(128 B) << 16 ; hi = SBO(1024 B) | version 1 << 14
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16); }
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14);
// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = 128 B between the two K chunks of one
// MMA, SBO = 1024 B between 8-row groups (a stage row holds 8 chunks); descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// debug timeline of CTA 0 (cycles): [0] producer warp 0 after a_empty wait, [1] after its a_full arrive,
// [2] MMA thread after a_full wait, [3] after commit, [4] b_full wait done.  Read with st_debug_tc_trace.
__device__ long long g_tc_trace[8][128];
__device__ int g_tc_dbg;
#ifdef ST_TC_TRACE_ON
#define TC_TRACE(slot, idx) do { if (blockIdx.x == (unsigned)g_tc_dbg && (idx) < 128) g_tc_trace[slot][idx] = clock64(); } while (0)
#else
#define TC_TRACE(slot, idx) do { } while (0)
#endif

struct TcArgs {
    const float *in;
    int in_ld;
    const int32_t *map;
    int n_out;
    int ntaps;
    const float *wprep;   // [nstages][2][npad*32] core-matrix tiles (hi, lo)
    int cin, cout, npad, nstages, b_stages;
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
    int nstages_main;     // stages that walk the taps; the rest (fused identity) walk the channels of in2
    const int32_t *row_index;      // optional: launch row -> output row (parity-sorted inverse convs)
    const uint32_t *tile_mask;     // optional: per 128-row tile, the taps that have at least one neighbour
};

// Stage s of a CIN-channel conv covers these taps (bit t = tap t).
template <int CIN>
__device__ __forceinline__ uint32_t stage_tap_bits(int s) {
    if (CIN == 8) return 0xFu << (4 * s);
    if (CIN == 16) return 0x3u << (2 * s);
    return 1u << (s / (CIN / TC_KS));
}
// Bit s set = stage s has work.  tap_mask: taps through which at least one row of the tile has a neighbour
// (inverse convs on parity-sorted rows: 1-8 of 27); the fused-identity stages beyond nmain are always active.
template <int CIN>
__device__ __forceinline__ unsigned long long stage_mask_of(uint32_t tap_mask, int nmain, int nstages) {
    if (tap_mask == 0u) tap_mask = 1u;                   // an empty tile still runs one stage (the accumulator must be defined)
    unsigned long long m = 0ull;
    for (uint32_t r = tap_mask; r; r &= r - 1) {         // the 1-8 taps that occur: stages [t*CIN/32, ((t+1)*CIN-1)/32]
        const int t = __ffs((int)r) - 1;
        const int s0 = t * CIN / TC_KS, s1 = ((t + 1) * CIN - 1) / TC_KS;
        m |= ((2ull << s1) - 1ull) & ~((1ull << s0) - 1ull);
    }
    if (nstages > nmain) m |= ((nstages < 64 ? (1ull << nstages) : 0ull) - 1ull) & ~((1ull << nmain) - 1ull);
    return m & ((nstages < 64 ? (1ull << nstages) : 0ull) - 1ull);
}
__device__ __forceinline__ int next_stage(unsigned long long m, int s, int nstages) {      // smallest active stage > s
    const unsigned long long r = s + 1 < 64 ? (m >> (s + 1)) : 0ull;
    return r ? s + __ffsll((long long)r) : nstages;
}

template <int CIN, bool MASKED>
__global__ void __launch_bounds__(TC_THREADS, 2) k_conv_tc(TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // smem: sb x { B_hi | B_lo } (npad*128 B each).  TMEM: [0, npad) accumulator, then TC_STAGES x 64 A columns.
    const int b_tile_bytes = a.npad * TC_KS * 4;
    const int b_stage_bytes = 2 * b_tile_bytes;
    uint8_t *const b_ring = smem_raw;
    const int SB = a.b_stages;
    __shared__ uint64_t a_full[TC_STAGES], a_empty[TC_STAGES], b_full[TC_MAX_BSTAGES], b_empty[TC_MAX_BSTAGES], accum_bar[2];
    __shared__ uint32_t tmem_base_sh;
    __shared__ int smap[2][TC_MAXTAPS][TC_M];     // gather-map entries of the current / next tile (28 KB)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // MASKED (parity-sorted inverse conv: 1-4 stages per tile, so the tile switch dominates): two accumulators, the
    // epilogue of tile i runs after the stages of tile i+1 have been produced and overlaps with their MMAs
    constexpr uint32_t NACC = MASKED ? 2u : 1u;
    const uint32_t tmem_need = NACC * (uint32_t)a.npad + TC_STAGES * TC_A_COLS;
    const uint32_t tmem_cols = tmem_need <= 256 ? 256u : 512u;
    const uint32_t a_col0 = NACC * (uint32_t)a.npad;
    const int ntiles = (a.n_out + TC_M - 1) / TC_M;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&a_full[s], TC_PRODUCERS / 32); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TC_MAX_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(&accum_bar[0], 1); mbar_init(&accum_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(&tmem_base_sh, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_sh, 0);

    if (warp < 8) {
        // ================= producers (then epilogue), persistent over tiles =================
        const int rloc = 32 * (warp & 3) + lane;      // row within the tile == TMEM lane
        const int half = warp >> 2;                   // which 4 of the 8 16-byte chunks of a stage row
        const uint32_t row_off = (uint32_t)((rloc >> 3) * 1024 + (rloc & 7) * 16);   // bytes inside a tile
        // Per stage this thread owns 16 consecutive K elements = 4 chunks of 16 bytes:
        //   CIN= 8: two taps (2 chunks each)   CIN=16: one tap   CIN>=32: 16 channels of one tap
        constexpr int TAPS_PER_THREAD = CIN == 8 ? 2 : 1;
        auto taps_of = [&](int s, int &t0, int &c0) {
            if (CIN == 8)       { t0 = 4 * s + 2 * half; c0 = 0; }
            else if (CIN == 16) { t0 = 2 * s + half;     c0 = 0; }
            else                { constexpr int SPT = CIN / TC_KS; t0 = s / SPT; c0 = (s % SPT) * TC_KS + 16 * half; }
        };
        // Gather-map entries are read exactly once (always cold in cache), so they are fetched one whole
        // TILE ahead with 4-byte cp.async copies into shared memory and never touch a register scoreboard.
        auto prefetch_map = [&](int tile, int buf) {
            if (tile < ntiles) {
                const int row = tile * TC_M + rloc;
                if (MASKED) {
                    // only the taps that occur in the tile are fetched (1-8 of 27); load_rows treats the others as absent
                    int q = 0;
                    for (uint32_t r = __ldg(a.tile_mask + tile); r; r &= r - 1, ++q) {
                        if ((q & 1) != half) continue;              // the two producer halves alternate
                        const int t = __ffs((int)r) - 1;
                        if (row < a.n_out) {
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&smap[buf][t][rloc])),
                                         "l"(a.map + (size_t)t * a.n_out + row) : "memory");
                        } else {
                            smap[buf][t][rloc] = -1;
                        }
                    }
                } else
                for (int t = half; t < TC_MAXTAPS; t += 2) {
                    if (row < a.n_out && t < a.ntaps) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&smap[buf][t][rloc])),
                                     "l"(a.map + (size_t)t * a.n_out + row) : "memory");
                    } else {
                        smap[buf][t][rloc] = -1;
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto map_ready = [&]() {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");      // producers only
        };
        // Gather: a QUAD of lanes reads 64 contiguous bytes of one feature row, so a warp-wide load touches
        // 8 cache lines instead of 32 (the L1 data stage serialises per line: it was the busiest unit with
        // one lane per row).  Thread 4*gq + q fetches chunk q of rows gq, gq+8, gq+16, gq+24 of its warp's
        // 32-row quadrant -- exactly the fragment tcgen05.st.16x256b scatters into TMEM lanes.
        const int gq = lane >> 2, q4 = lane & 3;
        const int rq = 32 * (warp & 3) + gq;
        int tile_row0 = 0;
        uint32_t cur_tmask = 0xFFFFFFFFu;             // MASKED: taps of the current tile that were prefetched
        auto load_rows = [&](int buf, int s, float4 (&x)[4]) {
            if (s >= a.nstages_main) {
                // fused ResBlock identity: K continues over the channels of in2, gathered through the identity map
                // (row i reads in2[i]); its 1x1 weights, pre-divided by the BN scale, follow the taps in wprep
                const int c = (s - a.nstages_main) * TC_KS + 16 * half + 4 * q4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = tile_row0 + rq + 8 * j;
                    x[j] = (r < a.n_out && c < a.cin2) ? __ldg((const float4 *)(a.in2 + (size_t)r * a.in2_ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                return;
            }
            int t0, c0;
            taps_of(s, t0, c0);
            const int tap = CIN == 8 ? t0 + (q4 >> 1) : t0;
            const int c = CIN == 8 ? (q4 & 1) * 4 : c0 + q4 * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int src = (tap < TC_MAXTAPS && (!MASKED || ((cur_tmask >> tap) & 1u))) ? smap[buf][tap][rq + 8 * j] : -1;
                x[j] = src >= 0 ? __ldg((const float4 *)(a.in + (size_t)src * a.in_ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // epilogue of one tile: e_row = launch row (residual / second input are indexed by it), e_orow = output row
        auto epilogue = [&](int e_row, bool e_ok, int e_orow, int e_iter) {
            mbar_wait(&accum_bar[MASKED ? (e_iter & 1) : 0], (uint32_t)((MASKED ? (e_iter >> 1) : e_iter) & 1));
            tc_fence_after();
            const int ncols_half = a.npad / 2;             // columns owned by this warp group
            const int col0 = half * ncols_half;
            for (int cb = 0; cb < ncols_half; cb += 8) {
                const int c = col0 + cb;
                float v[8];
                tmem_ld8(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((MASKED ? (e_iter & 1) * a.npad : 0) + c), v);
                if (!e_ok || c >= a.cout) continue;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float sc = a.scale ? __ldg(a.scale + c + i) : 1.f;
                    float sh = a.shift ? __ldg(a.shift + c + i) : 0.f;
                    v[i] = fmaf(v[i], sc, sh);
                }
                if (a.res) {
                    const float4 *rp = (const float4 *)(a.res + (size_t)e_row * a.res_ld + c);
                    float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
                }
                if (a.in2 && a.w2) {
                    float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    const float *xr = a.in2 + (size_t)e_row * a.in2_ld;
                    for (int ci = 0; ci < a.cin2; ++ci) {
                        const float xv = __ldg(xr + ci);
                        const float4 *wp = (const float4 *)(a.w2 + (size_t)ci * a.cout + c);
                        float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
                        e[0] = fmaf(xv, w0.x, e[0]); e[1] = fmaf(xv, w0.y, e[1]); e[2] = fmaf(xv, w0.z, e[2]); e[3] = fmaf(xv, w0.w, e[3]);
                        e[4] = fmaf(xv, w1.x, e[4]); e[5] = fmaf(xv, w1.y, e[5]); e[6] = fmaf(xv, w1.z, e[6]); e[7] = fmaf(xv, w1.w, e[7]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += e[i];
                }
                if (a.act & ST_ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                float4 *op = (float4 *)(a.out + (size_t)e_orow * a.out_ld + c);
                op[0] = make_float4(v[0], v[1], v[2], v[3]);
                op[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            // the accumulator has been read: order those TMEM loads before this thread's next arrivals,
            // which in turn gate the next tile's first (overwriting) MMA
            tc_fence_before();
        };
        bool pend = false, p_ok = false;
        int p_row = 0, p_orow = 0, p_iter = 0;
        int g = 0, st = 0;                            // global stage counter, ring position
        uint32_t pe = 1;                              // parity of the previous use of A stage `st`
        int tile_iter = 0;
        prefetch_map(blockIdx.x, 0);
        map_ready();
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
            const int buf = tile_iter & 1;
            tile_row0 = tile * TC_M;
            const int row = tile * TC_M + rloc;
            const bool row_ok = row < a.n_out;
            prefetch_map(tile + gridDim.x, buf ^ 1);
            // software pipeline: feature rows two stages ahead of the stores (covers the tail latency of the
            // ~1000 gathers a stage consists of; the slowest one gates the whole CTA).  The three register
            // sets rotate by unrolling the stage loop three times -- no register-to-register copies.
            float4 x0[4], x1[4], x2[4];
            auto do_stage = [&](int s_fill, float4 (&cur)[4], float4 (&fill)[4]) {       // s_fill: the stage two ahead in the walk
                if (tid == 0) TC_TRACE(6, g);
                if (s_fill < a.nstages) load_rows(buf, s_fill, fill);
                if (tid == 0) TC_TRACE(7, g);
                if (g >= TC_STAGES) {
                    if (lane == 0) mbar_wait(&a_empty[st], pe);
                    __syncwarp();
                    tc_fence_after();     // the MMAs that read this TMEM stage have completed
                }
                if (tid == 0) TC_TRACE(0, g);
                // TMEM address of this thread's 16 columns of stage st: lanes 32*(warp&3).., columns a_col0 + st*64 + half*16
                const uint32_t ta = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + a_col0 + (uint32_t)st * TC_A_COLS + (uint32_t)half * 16;
                // x = hi + lo exactly; hi carries the top 10 mantissa bits (what kind::tf32 reads), lo the rest
                auto split = [](float v, uint32_t &h, uint32_t &l) {
                    h = __float_as_uint(v) & 0xFFFFE000u;
                    l = __float_as_uint(v - __uint_as_float(h));
                };
#pragma unroll
                for (int p = 0; p < 2; ++p) {             // lanes +0..15 (rows gq, gq+8) then +16..31
                    const float4 u = cur[2 * p], w = cur[2 * p + 1];
                    uint32_t hi[8], lo[8];
                    split(u.x, hi[0], lo[0]); split(u.y, hi[1], lo[1]); split(w.x, hi[2], lo[2]); split(w.y, hi[3], lo[3]);
                    split(u.z, hi[4], lo[4]); split(u.w, hi[5], lo[5]); split(w.z, hi[6], lo[6]); split(w.w, hi[7], lo[7]);
                    tmem_st_quad(ta + ((uint32_t)(16 * p) << 16), hi);
                    tmem_st_quad(ta + ((uint32_t)(16 * p) << 16) + TC_KS, lo);
                }
                tmem_st_wait();
                tc_fence_before();        // order the TMEM writes before the arrive that hands them to the MMA thread
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[st]);      // one arrival per producer warp
                if (tid == 0) TC_TRACE(1, g);
                ++g;
                if (++st == TC_STAGES) { st = 0; pe ^= 1; }
            };
            // walk the ACTIVE stages only (all of them unless a tile mask is given)
            if (MASKED) cur_tmask = __ldg(a.tile_mask + tile);
            const unsigned long long smask = MASKED ? stage_mask_of<CIN>(cur_tmask, a.nstages_main, a.nstages) : ~0ull;
            const int ns = a.nstages;
            int sa = (MASKED ? next_stage(smask, -1, a.nstages) : (-1) + 1), sb = (MASKED ? next_stage(smask, sa, a.nstages) : (sa) + 1), sc = (MASKED ? next_stage(smask, sb, a.nstages) : (sb) + 1);
            load_rows(buf, sa, x0);
            if (sb < ns) load_rows(buf, sb, x1);
            while (sa < ns) {
                do_stage(sc, x0, x2);
                sa = sb; sb = sc; sc = (MASKED ? next_stage(smask, sc, a.nstages) : (sc) + 1);
                if (sa < ns) { do_stage(sc, x1, x0); sa = sb; sb = sc; sc = (MASKED ? next_stage(smask, sc, a.nstages) : (sc) + 1); }
                if (sa < ns) { do_stage(sc, x2, x1); sa = sb; sb = sc; sc = (MASKED ? next_stage(smask, sc, a.nstages) : (sc) + 1); }
            }
            // ---------------- epilogue: of this tile, or (MASKED) of the previous one ----------------
            if (!MASKED) {
                epilogue(row, row_ok, row, tile_iter);
            } else {
                if (pend) epilogue(p_row, p_ok, p_orow, p_iter);
                pend = true; p_row = row; p_ok = row_ok; p_iter = tile_iter;
                p_orow = row_ok ? __ldg(a.row_index + row) : row;
            }
            map_ready();                   // next tile's gather map has landed; this tile's buffer may be refilled
        }
        if (MASKED && pend) epilogue(p_row, p_ok, p_orow, p_iter);
    } else if (warp == 8) {
        // ================= MMA issuer (one elected lane) =================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        {
            // The whole warp walks the stages in lock step (so every address below is warp-uniform and lives
            // in uniform registers); one elected lane issues.  Ring positions and phase bits are kept
            // incrementally: this warp is the serial bottleneck of the CTA.
            int g = 0, st = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            const uint32_t a_stage0 = tmem_base + a_col0;
            const uint32_t b_base = desc_lo(smem_u32(b_ring)), b_stage_units = (uint32_t)(b_stage_bytes >> 4), b_lo_off = (uint32_t)(b_tile_bytes >> 4);
            int tile_iter = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
                const unsigned long long smask = MASKED ? stage_mask_of<CIN>(__ldg(a.tile_mask + tile), a.nstages_main, a.nstages) : ~0ull;
                const uint32_t tmem_d = tmem_base + (MASKED ? (uint32_t)((tile_iter & 1) * a.npad) : 0u);
                const int s_first = (MASKED ? next_stage(smask, -1, a.nstages) : (-1) + 1);
                for (int s = s_first; s < a.nstages; s = (MASKED ? next_stage(smask, s, a.nstages) : (s) + 1), ++g) {
                    const bool s_last = (MASKED ? next_stage(smask, s, a.nstages) : (s) + 1) >= a.nstages;
                    mbar_wait(&b_full[sb], pb);
                    if (lane == 0) TC_TRACE(4, g);
                    mbar_wait(&a_full[st], pa);
                    if (lane == 0) TC_TRACE(2, g);
                    tc_fence_after();
                    const uint32_t ah = a_stage0 + (uint32_t)st * TC_A_COLS, al = ah + TC_KS;
                    const uint32_t bh = b_base + (uint32_t)sb * b_stage_units, bl = bh + b_lo_off;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < TC_KS / 8; ++j) {
                            const uint32_t ko = (uint32_t)j * 16;        // B: 256 bytes per k-step; A: 8 TMEM columns
                            umma_tf32_ts(tmem_d, ah + j * 8, bh + ko, DESC_HI, idesc, (s != s_first || j) ? 1u : 0u);
                            umma_tf32_ts(tmem_d, al + j * 8, bh + ko, DESC_HI, idesc, 1u);
                            umma_tf32_ts(tmem_d, ah + j * 8, bl + ko, DESC_HI, idesc, 1u);
                        }
                        umma_commit(&a_empty[st]);    // both ring slots are reusable once these MMAs have read them
                        umma_commit(&b_empty[sb]);
                        if (s_last) umma_commit(&accum_bar[MASKED ? (tile_iter & 1) : 0]);      // this tile's accumulator is complete
                    }
                    __syncwarp();
                    if (lane == 0) TC_TRACE(3, g);
                    if (++st == TC_STAGES) { st = 0; pa ^= 1; }
                    if (++sb == SB) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else {
        // ================= weight loader (one elected lane): SB stages ahead of the MMAs =================
        {
            int g = 0, sb = 0;
            uint32_t pb = 1;              // parity of the *previous* use of the slot
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const unsigned long long smask = MASKED ? stage_mask_of<CIN>(__ldg(a.tile_mask + tile), a.nstages_main, a.nstages) : ~0ull;
                for (int s = (MASKED ? next_stage(smask, -1, a.nstages) : (-1) + 1); s < a.nstages; s = (MASKED ? next_stage(smask, s, a.nstages) : (s) + 1), ++g) {
                    if (g >= SB) mbar_wait(&b_empty[sb], pb);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&b_full[sb], (uint32_t)b_stage_bytes);
                        bulk_copy_g2s(b_ring + (size_t)sb * b_stage_bytes, a.wprep + (size_t)s * 2 * a.npad * TC_KS, (uint32_t)b_stage_bytes, &b_full[sb]);
                    }
                    __syncwarp();
                    if (++sb == SB) { sb = 0; pb ^= 1; }
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// w [ntaps, cin, cout] -> wprep [nstages][2][npad*32]: per stage the K-major core-matrix tile of the
// TF32 heads followed by the tile of the tails; K index kk = tap*cin + c, zero padded.
__global__ void k_tc_prepare(const float *__restrict__ w, int ntaps, int cin, int cout, int npad, int nstages, float *__restrict__ wprep,
                             const float *__restrict__ w2, int cin2, const float *__restrict__ scale, int nstages_main) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)nstages * npad * TC_KS;
    if (idx >= total) return;
    int s = (int)(idx / (npad * TC_KS));
    int rem = (int)(idx % (npad * TC_KS));
    int n = rem / TC_KS, kk = rem % TC_KS;
    if (s >= nstages_main) {
        // fused identity stages: logical K element = channel of in2; same fragment permutation as below
        const int cc = kk & 15, q = (cc & 7) >> 1, m = (cc & 1) + ((cc >> 3) << 1);
        const int c2 = (s - nstages_main) * TC_KS + (kk & 16) + 4 * q + m;
        float v = (c2 < cin2 && n < cout) ? w2[(size_t)c2 * cout + n] / (scale ? scale[n] : 1.f) : 0.f;
        float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        float lo = v - hi;
        size_t pos = (size_t)(n >> 3) * 256 + (size_t)(kk >> 2) * 32 + (size_t)(n & 7) * 4 + (kk & 3);
        float *tile = wprep + (size_t)s * 2 * npad * TC_KS;
        tile[pos] = hi;
        tile[(size_t)npad * TC_KS + pos] = lo;
        return;
    }
    // TMEM column kk of a stage holds logical K element 16*(kk/16) + 4*q + m, where the quad fragment puts
    // (q, m) at column 2q+m (m < 2) or 8+2q+(m-2) of its 16-column half (see load_rows / tmem_st_quad)
    const int cc = kk & 15, q = (cc & 7) >> 1, m = (cc & 1) + ((cc >> 3) << 1);
    int K = s * TC_KS + (kk & 16) + 4 * q + m;
    int tap = K / cin, c = K % cin;
    float v = (tap < ntaps && n < cout) ? w[((size_t)tap * cin + c) * cout + n] : 0.f;
    float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    float lo = v - hi;
    size_t pos = (size_t)(n >> 3) * 256 + (size_t)(kk >> 2) * 32 + (size_t)(n & 7) * 4 + (kk & 3);
    float *tile = wprep + (size_t)s * 2 * npad * TC_KS;
    tile[pos] = hi;
    tile[(size_t)npad * TC_KS + pos] = lo;
}

int tc_npad(int cout) { int n = (cout + 15) / 16 * 16; return n < 16 ? 16 : n; }
int tc_nstages(int ntaps, int cin) { return (ntaps * cin + TC_KS - 1) / TC_KS; }
bool tc_supported(int cin, int cout) { return (cin == 8 || cin == 16 || cin == 32 || cin == 64 || cin == 128) && cout % 8 == 0 && cout <= 256; }


// =====================================================================================================
// Tile-plan path: the neighbour rows of a 128-row tile are staged in SHARED MEMORY once per tile.
//
// Rows are kept in Z-order, so the 27 x 128 gather-map entries of a tile name only ~2 x 128 DISTINCT source
// rows (measured on the bench tree: mean 260, max 385).  k_conv_tc above fetches every (row, tap) pair
// through the L1 with per-thread loads and is bound by the L1 line-visit rate (~2 cycles per line touched,
// 16-19 visits per source row).  Here a per-map PLAN (built once per level, shared by its 4-5 convs) lists
// the distinct source rows of every tile (sorted, so the staging copies are coalesced) and a 16-bit LOCAL
// gather map; the kernel stages those rows with 16-byte cp.async copies (every feature line crosses the
// L1 once per tile, double buffered across tiles) and the producers gather their A fragments from shared
// memory instead.  Everything downstream (A in TMEM, 3xTF32 MMAs, weight ring, epilogue) is as above.
// A tile whose distinct rows do not fit (scattered points) is split into eight 16-row sub-tiles, which
// always fit (16 x 27 < TP_NU).
constexpr int TP_NU = 512;                       // distinct source rows a (sub-)tile may stage
constexpr int TP_TAPS = 28;                      // local-map rows per tile (27 taps + one all-absent pad tap)
constexpr int TP_SUBS = 8;
constexpr int TP_HDR = 20;                       // int32 words per tile: [0] split (0 | 1), [1..8] distinct rows of sub-tile j, [9..16] runs of consecutive rows
constexpr int TP_LMAP_BYTES = TP_TAPS * TC_M * 2;
constexpr uint16_t TP_NONE = TP_NU;                // local index of the all-zero row

struct PlanLayout {
    size_t lmap_off, rows_off, total;
    explicit PlanLayout(int64_t nblocks) {
        lmap_off = align_up((size_t)nblocks * TP_HDR * 4);
        rows_off = lmap_off + align_up((size_t)nblocks * TP_LMAP_BYTES);
        total = rows_off + align_up((size_t)nblocks * TP_SUBS * TP_NU * 8);      // int2 (first local row, first source row) per run
    }
};

constexpr int PB_THREADS = 256, PB_ITEMS = 14;   // 3584 >= 27 * 128 map entries per tile

__global__ void __launch_bounds__(PB_THREADS) k_plan_build(const int32_t *__restrict__ map, int n_out, int ntaps, int n_in, int end_bit,
                                                           int32_t *__restrict__ hdr, uint16_t *__restrict__ lmap, int2 *__restrict__ runs) {
    using Sort = cub::BlockRadixSort<int, PB_THREADS, PB_ITEMS>;
    using Scan = cub::BlockScan<int, PB_THREADS>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ int s_sorted[PB_THREADS * PB_ITEMS];
    __shared__ int s_u[TP_NU];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int row0 = b * TC_M;
    // distinct source rows of tile rows [r_begin, r_begin + r_count), ascending, into s_u; returns their number
    auto build = [&](int r_begin, int r_count) -> int {
        const int entries = ntaps * r_count;
        int keys[PB_ITEMS];
#pragma unroll
        for (int i = 0; i < PB_ITEMS; ++i) {
            const int e = tid * PB_ITEMS + i;
            int v = n_in;                                   // "absent" sorts last
            if (e < entries) {
                const int tap = e / r_count, row = row0 + r_begin + e % r_count;
                if (row < n_out) { const int m = __ldg(map + (size_t)tap * n_out + row); if (m >= 0) v = m; }
            }
            keys[i] = v;
        }
        Sort(tmp.sort).Sort(keys, 0, end_bit);
#pragma unroll
        for (int i = 0; i < PB_ITEMS; ++i) s_sorted[tid * PB_ITEMS + i] = keys[i];
        __syncthreads();
        int cnt = 0;
        unsigned heads = 0;
#pragma unroll
        for (int i = 0; i < PB_ITEMS; ++i) {
            const int idx = tid * PB_ITEMS + i;
            const int prev = idx ? s_sorted[idx - 1] : -1;
            if (keys[i] != n_in && keys[i] != prev) { heads |= 1u << i; ++cnt; }
        }
        int base, total;
        Scan(tmp.scan).ExclusiveSum(cnt, base, total);
        if (total <= TP_NU) {
#pragma unroll
            for (int i = 0; i < PB_ITEMS; ++i)
                if (heads & (1u << i)) s_u[base++] = keys[i];
        }
        __syncthreads();
        return total;
    };
    __shared__ int s_nr;
    auto emit = [&](int r_begin, int r_count, int sub, int nu) {
        // runs of consecutive source rows (the list is sorted): one bulk copy each when the kernel stages them
        int2 *rdst = runs + ((size_t)b * TP_SUBS + sub) * TP_NU;
        int heads[2], cnt = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = 2 * tid + k;
            heads[k] = i < nu && (i == 0 || s_u[i] != s_u[i - 1] + 1);
            cnt += heads[k];
        }
        int base, total;
        Scan(tmp.scan).ExclusiveSum(cnt, base, total);
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (heads[k]) rdst[base++] = make_int2(2 * tid + k, s_u[2 * tid + k]);
        if (tid == 0) s_nr = total;
        const int entries = TP_TAPS * r_count;
        for (int e = tid; e < entries; e += PB_THREADS) {
            const int tap = e / r_count, r = r_begin + e % r_count, row = row0 + r;
            uint16_t li = TP_NONE;
            if (tap < ntaps && row < n_out) {
                const int m = __ldg(map + (size_t)tap * n_out + row);
                if (m >= 0) {
                    int lo = 0, hi = nu - 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_u[mid] < m) lo = mid + 1; else hi = mid; }
                    li = (uint16_t)lo;
                }
            }
            // the four rows a producer thread gathers for (r, r+8, r+16, r+24 of a 32-row quadrant) sit side by side
            lmap[(size_t)b * (TP_TAPS * TC_M) + tap * TC_M + (r & ~31) + (r & 7) * 4 + ((r >> 3) & 3)] = li;
        }
        __syncthreads();
    };
    int nu = build(0, TC_M);
    if (nu <= TP_NU) {
        emit(0, TC_M, 0, nu);
        if (tid == 0) { hdr[b * TP_HDR] = 0; hdr[b * TP_HDR + 1] = nu; hdr[b * TP_HDR + 1 + TP_SUBS] = s_nr; }
    } else {
        for (int j = 0; j < TP_SUBS; ++j) {
            nu = build(j * (TC_M / TP_SUBS), TC_M / TP_SUBS);
            emit(j * (TC_M / TP_SUBS), TC_M / TP_SUBS, j, nu);
            if (tid == 0) { hdr[b * TP_HDR + 1 + j] = nu; hdr[b * TP_HDR + 1 + TP_SUBS + j] = s_nr; }
            __syncthreads();
        }
        if (tid == 0) hdr[b * TP_HDR] = 1;
    }
}

struct TpArgs {
    const float *in;
    int in_ld;
    const int32_t *hdr;
    const uint16_t *lmap;
    const int2 *runs;
    int n_out, nblocks;
    const float *wprep;
    int cin, cout, npad, nstages, b_stages;
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
};

template <int CIN, int NBUF>
__global__ void __launch_bounds__(TC_THREADS, 2) k_conv_tp(TpArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int CH = CIN / 4;                          // 16-byte chunks per feature row
    constexpr int FEAT_BYTES = (TP_NU + 1) * CIN * 4;          // + one all-zero row: where absent neighbours point
    constexpr int SLOT_BYTES = FEAT_BYTES + TP_LMAP_BYTES;        // one staging slot: source rows + local gather map
    const int b_tile_bytes = a.npad * TC_KS * 4;
    const int b_stage_bytes = 2 * b_tile_bytes;
    const int SB = a.b_stages;
    uint8_t *const b_ring = smem_raw;
    uint8_t *const slots = smem_raw + SB * b_stage_bytes;         // NBUF x { [TP_NU][CIN] floats | [TP_TAPS][128] u16 }
    __shared__ uint64_t a_full[TC_STAGES], a_empty[TC_STAGES], b_full[TC_MAX_BSTAGES], b_empty[TC_MAX_BSTAGES], accum_bar;
    __shared__ uint64_t f_full[2], f_empty[2];
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef ST_TC_TRACE_ON
    if (tid == 0 && blockIdx.x == (unsigned)g_tc_dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_tc_trace[5][120] = (long long)gt; g_tc_trace[5][121] = clock64(); }
#endif
    const uint32_t tmem_need = 2u * (uint32_t)a.npad + TC_STAGES * TC_A_COLS;      // two accumulators: epilogue of item i overlaps the MMAs of item i+1
    const uint32_t tmem_cols = tmem_need <= 256 ? 256u : 512u;
    const uint32_t a_col0 = 2u * (uint32_t)a.npad;
    const int nblocks = a.nblocks;
    auto nsub_of = [&](int b) { return __ldg(a.hdr + (size_t)b * TP_HDR) ? TP_SUBS : 1; };

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&a_full[s], TC_PRODUCERS / 32); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TC_MAX_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&f_full[s], 1); mbar_init(&f_empty[s], TC_PRODUCERS / 32); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(&tmem_base_sh, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_sh, 0);
#ifdef ST_TC_TRACE_ON
    if (tid == 0 && blockIdx.x == (unsigned)g_tc_dbg) g_tc_trace[5][124] = clock64();
#endif

    if (warp < 8) {
        // ================= producers (then epilogue), persistent over (tile, sub-tile) items =================
        const int rloc = 32 * (warp & 3) + lane;      // row within the tile == TMEM lane (epilogue)
        const int half = warp >> 2;
        auto taps_of = [&](int s, int &t0, int &c0) {
            if (CIN == 8)       { t0 = 4 * s + 2 * half; c0 = 0; }
            else if (CIN == 16) { t0 = 2 * s + half;     c0 = 0; }
            else                { constexpr int SPT = CIN / TC_KS; t0 = s / SPT; c0 = (s % SPT) * TC_KS + 16 * half; }
        };
        // Gather from shared memory: thread 4*gq + q fetches chunk q of rows gq, gq+8, gq+16, gq+24 of its warp's
        // 32-row quadrant -- the fragment tcgen05.st.16x256b scatters into TMEM lanes (as in k_conv_tc).  The
        // local map stores those four rows' entries side by side (one 8-byte load per thread and stage).
        const int gq = lane >> 2, q4 = lane & 3;
        const int rq = 32 * (warp & 3) + gq;
        const int lm_off = (32 * (warp & 3) + 4 * gq) * 2;          // byte offset of this thread's entries inside a tap row
        // epilogue of item e_it (accumulator e_it & 1), run by the producers while the MMA warp works on the next item
        auto epilogue = [&](int e_row, bool e_ok, int e_it) {
        mbar_wait(&accum_bar, (uint32_t)(e_it & 1));
        tc_fence_after();
        const int ncols_half = a.npad / 2;
        const int col0 = half * ncols_half;
        for (int cb = 0; cb < ncols_half; cb += 8) {
            const int c = col0 + cb;
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((e_it & 1) * a.npad + c), v);
            if (!e_ok || c >= a.cout) continue;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float sc = a.scale ? __ldg(a.scale + c + i) : 1.f;
                float sh = a.shift ? __ldg(a.shift + c + i) : 0.f;
                v[i] = fmaf(v[i], sc, sh);
            }
            if (a.res) {
                const float4 *rp = (const float4 *)(a.res + (size_t)e_row * a.res_ld + c);
                float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
            }
            if (a.in2) {
                float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const float *xr = a.in2 + (size_t)e_row * a.in2_ld;
                for (int ci = 0; ci < a.cin2; ++ci) {
                    const float xv = __ldg(xr + ci);
                    const float4 *wp = (const float4 *)(a.w2 + (size_t)ci * a.cout + c);
                    float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
                    e[0] = fmaf(xv, w0.x, e[0]); e[1] = fmaf(xv, w0.y, e[1]); e[2] = fmaf(xv, w0.z, e[2]); e[3] = fmaf(xv, w0.w, e[3]);
                    e[4] = fmaf(xv, w1.x, e[4]); e[5] = fmaf(xv, w1.y, e[5]); e[6] = fmaf(xv, w1.z, e[6]); e[7] = fmaf(xv, w1.w, e[7]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += e[i];
            }
            if (a.act & ST_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            float4 *op = (float4 *)(a.out + (size_t)e_row * a.out_ld + c);
            op[0] = make_float4(v[0], v[1], v[2], v[3]);
            op[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
            tc_fence_before();
        };
        bool pend = false, p_ok = false;
        int p_row = 0, p_it = 0;
        for (int i = tid; i < NBUF * CH; i += TC_PRODUCERS)
            *(float4 *)(slots + (size_t)(i / CH) * SLOT_BYTES + (size_t)TP_NU * (CIN * 4) + (i % CH) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");      // producers only
        int g = 0, st = 0;
        uint32_t pe = 1;
        bool slot_free = false;
        int it = 0;
        int b = blockIdx.x, j = 0;
        int nsub = b < nblocks ? nsub_of(b) : 1;
        while (b < nblocks) {
            int nb = b, nj = j + 1, nnsub = nsub;
            if (nj == nsub) { nb = b + gridDim.x; nj = 0; nnsub = nb < nblocks ? nsub_of(nb) : 1; }
            const int fbuf = NBUF == 2 ? (it & 1) : 0;
            const uint8_t *const fcur = slots + (size_t)fbuf * SLOT_BYTES;
            const uint8_t *const lm = fcur + FEAT_BYTES;
            const int row = b * TC_M + rloc;
            const bool row_ok = row < a.n_out && (nsub == 1 || (rloc >> 4) == j);
            auto load_lm = [&](int s) -> uint2 {
                int t0, c0;
                taps_of(s, t0, c0);
                const int tap = CIN == 8 ? t0 + (q4 >> 1) : t0;
                return *(const uint2 *)(lm + tap * (TC_M * 2) + lm_off);
            };
            auto load_rows = [&](int s, uint2 e, float4 (&x)[4]) {
                int t0, c0;
                taps_of(s, t0, c0);
                const int c = CIN == 8 ? (q4 & 1) : (c0 >> 2) + q4;
                // Absent neighbours point at the all-zero row.  In a split tile the rows outside the current sub-tile
                // gather whatever their entries name (entries are relative to their own sub-tile's list): that only
                // reaches accumulator lanes the epilogue of this item does not read.
                const uint8_t *const fc = fcur + c * 16;
                x[0] = *(const float4 *)(fc + (e.x & 0xFFFFu) * (CIN * 4));
                x[1] = *(const float4 *)(fc + (e.x >> 16) * (CIN * 4));
                x[2] = *(const float4 *)(fc + (e.y & 0xFFFFu) * (CIN * 4));
                x[3] = *(const float4 *)(fc + (e.y >> 16) * (CIN * 4));
            };
            // the staged rows and local map of this item have landed (bulk copies issued by the loader warp)
            if (tid == 0) TC_TRACE(5, g);
            mbar_wait(&f_full[fbuf], (uint32_t)((NBUF == 2 ? (it >> 1) : it) & 1));
            if (tid == 0) TC_TRACE(5, g + 1);
            float4 x0[4], x1[4];
            uint2 l0, l1;
            auto do_stage = [&](int s, float4 (&cur)[4], float4 (&fill)[4], uint2 &lmf) {
                if (tid == 0) TC_TRACE(6, g);
                if (g >= TC_STAGES) {
                    if (!slot_free) mbar_wait(&a_empty[st], pe);      // probed a stage ago: normally already free
                    tc_fence_after();
                }
                if (tid == 0) TC_TRACE(0, g);
                const uint32_t ta = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + a_col0 + (uint32_t)st * TC_A_COLS + (uint32_t)half * 16;
                auto split = [](float v, uint32_t &h, uint32_t &l) {
                    h = __float_as_uint(v) & 0xFFFFE000u;
                    l = __float_as_uint(v - __uint_as_float(h));
                };
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const float4 u = cur[2 * p], w = cur[2 * p + 1];
                    uint32_t hi[8], lo[8];
                    split(u.x, hi[0], lo[0]); split(u.y, hi[1], lo[1]); split(w.x, hi[2], lo[2]); split(w.y, hi[3], lo[3]);
                    split(u.z, hi[4], lo[4]); split(u.w, hi[5], lo[5]); split(w.z, hi[6], lo[6]); split(w.w, hi[7], lo[7]);
                    tmem_st_quad(ta + ((uint32_t)(16 * p) << 16), hi);
                    tmem_st_quad(ta + ((uint32_t)(16 * p) << 16) + TC_KS, lo);
                }
                // the TMEM stores are in flight: fetch the fragments of the next stage (their local-map entries are
                // already in registers) and the entries of stage s+3 meanwhile
                if (s + 1 < a.nstages) load_rows(s + 1, lmf, fill);
                if (s + 3 < a.nstages) lmf = load_lm(s + 3);
                if (tid == 0) TC_TRACE(7, g);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[st]);
                if (tid == 0) TC_TRACE(1, g);
                ++g;
                if (++st == TC_STAGES) { st = 0; pe ^= 1; }
                slot_free = g >= TC_STAGES && mbar_test(&a_empty[st], pe);       // next stage's slot, consumed a stage later
            };
            {
                const uint2 e0 = load_lm(0);
                l0 = load_lm(a.nstages > 1 ? 1 : 0);
                l1 = load_lm(a.nstages > 2 ? 2 : 0);
                load_rows(0, e0, x0);
            }
            for (int s = 0; s < a.nstages; s += 2) {
                do_stage(s, x0, x1, l0);
                if (s + 1 < a.nstages) do_stage(s + 1, x1, x0, l1);
                if (s == 2 && pend) { epilogue(p_row, p_ok, p_it); pend = false; }      // previous item, four stages into this one
            }
            if (pend) { epilogue(p_row, p_ok, p_it); pend = false; }                    // (items of fewer than three stages)
            // every fragment of this warp has been consumed: hand the staging slot back to the loader
            __syncwarp();
            if (lane == 0) mbar_arrive(&f_empty[fbuf]);
            pend = true; p_row = row; p_ok = row_ok; p_it = it;
            b = nb; j = nj; nsub = nnsub;
            ++it;
        }
        if (pend) epilogue(p_row, p_ok, p_it);
    } else if (warp == 8) {
        // ================= MMA issuer (one elected lane) =================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        int g = 0, st = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        const uint32_t a_stage0 = tmem_base + a_col0;
        const uint32_t b_base = desc_lo(smem_u32(b_ring)), b_stage_units = (uint32_t)(b_stage_bytes >> 4), b_lo_off = (uint32_t)(b_tile_bytes >> 4);
        int it = 0;
        for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
            const int nsub = nsub_of(b);
            for (int j = 0; j < nsub; ++j, ++it) {
                const uint32_t tmem_d = tmem_base + (uint32_t)((it & 1) * a.npad);
                for (int s = 0; s < a.nstages; ++s, ++g) {
                    mbar_wait(&b_full[sb], pb);
                    if (lane == 0) TC_TRACE(4, g);
                    mbar_wait(&a_full[st], pa);
                    if (lane == 0) TC_TRACE(2, g);
                    tc_fence_after();
                    const uint32_t ah = a_stage0 + (uint32_t)st * TC_A_COLS, al = ah + TC_KS;
                    const uint32_t bh = b_base + (uint32_t)sb * b_stage_units, bl = bh + b_lo_off;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < TC_KS / 8; ++k) {
                            const uint32_t ko = (uint32_t)k * 16;
                            umma_tf32_ts(tmem_d, ah + k * 8, bh + ko, DESC_HI, idesc, (s | k) ? 1u : 0u);
                            umma_tf32_ts(tmem_d, al + k * 8, bh + ko, DESC_HI, idesc, 1u);
                            umma_tf32_ts(tmem_d, ah + k * 8, bl + ko, DESC_HI, idesc, 1u);
                        }
                        umma_commit(&a_empty[st]);
                        umma_commit(&b_empty[sb]);
                        if (s == a.nstages - 1) umma_commit(&accum_bar);
                    }
                    __syncwarp();
                    if (lane == 0) TC_TRACE(3, g);
                    if (++st == TC_STAGES) { st = 0; pa ^= 1; }
                    if (++sb == SB) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else {
        // ================= loader warp: weight ring (one elected lane) + staging of the next item's rows =================
        // Every distinct source row of an item is one bulk async copy (TMA, UBLKCP) into the staging slot, the local
        // map one more; all complete on the slot's mbarrier.  Issued one item ahead, so the two dependent index
        // loads (row count, row ids -- cold, they stream from HBM once per conv) never sit on the producers' path.
        const bool dense_rows = a.in_ld == CIN;          // consecutive source rows are contiguous in memory
        auto stage_item = [&](int b, int j, int itn) {
            const int buf = NBUF == 2 ? (itn & 1) : 0, use = NBUF == 2 ? (itn >> 1) : itn;
            const int nu = __ldg(a.hdr + (size_t)b * TP_HDR + 1 + j), nr = __ldg(a.hdr + (size_t)b * TP_HDR + 1 + TP_SUBS + j);
            const int2 *rns = a.runs + ((size_t)b * TP_SUBS + j) * TP_NU;
            constexpr int RPL = 4;                       // runs per lane and round (index loads batched ahead of the copies)
            uint8_t *const dst = slots + (size_t)buf * SLOT_BYTES;
            if (use > 0) mbar_wait(&f_empty[buf], (uint32_t)((use - 1) & 1));
            if (lane == 0) {
                mbar_arrive_expect_tx(&f_full[buf], (uint32_t)(nu * CIN * 4 + TP_LMAP_BYTES));
                bulk_copy_g2s(dst + FEAT_BYTES, (const uint8_t *)a.lmap + (size_t)b * TP_LMAP_BYTES, TP_LMAP_BYTES, &f_full[buf]);
            }
            __syncwarp();
            for (int r0 = 0; r0 < nr; r0 += 32 * RPL) {
                int2 rn[RPL];
                int len[RPL];
#pragma unroll
                for (int k = 0; k < RPL; ++k) {
                    const int r = r0 + lane + 32 * k;
                    rn[k] = r < nr ? __ldg(rns + r) : make_int2(0, 0);
                    len[k] = r < nr ? (r + 1 < nr ? __ldg(&rns[r + 1].x) : nu) - rn[k].x : 0;
                }
#pragma unroll
                for (int k = 0; k < RPL; ++k) {
                    if (len[k] == 0) continue;
                    if (dense_rows) {
                        bulk_copy_g2s(dst + (size_t)rn[k].x * (CIN * 4), a.in + (size_t)rn[k].y * CIN, (uint32_t)(len[k] * CIN * 4), &f_full[buf]);
                    } else {
                        for (int i = 0; i < len[k]; ++i)
                            bulk_copy_g2s(dst + (size_t)(rn[k].x + i) * (CIN * 4), a.in + (size_t)(rn[k].y + i) * a.in_ld, CIN * 4, &f_full[buf]);
                    }
                }
            }
        };
        int g = 0, sb = 0, it = 0;
        uint32_t pb = 1;
        int b = blockIdx.x, j = 0;
        int nsub = b < nblocks ? nsub_of(b) : 1;
        if (b < nblocks) stage_item(b, 0, 0);
        // the next item is staged half way through the weight stages of the current one: by then the producers have
        // released its slot, and the weight ring still holds enough stages to keep the MMA warp busy meanwhile
        const int s_stage = NBUF == 2 ? a.nstages / 2 : a.nstages - 1;
        while (b < nblocks) {
            int nb = b, nj = j + 1, nnsub = nsub;
            if (nj == nsub) { nb = b + gridDim.x; nj = 0; nnsub = nb < nblocks ? nsub_of(nb) : 1; }
            const float *src = a.wprep;
            for (int s = 0; s < a.nstages; ++s, ++g) {
                if (g >= SB) mbar_wait(&b_empty[sb], pb);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&b_full[sb], (uint32_t)b_stage_bytes);
                    bulk_copy_g2s(b_ring + (size_t)sb * b_stage_bytes, src, (uint32_t)b_stage_bytes, &b_full[sb]);
                }
                __syncwarp();
                src += 2 * a.npad * TC_KS;
                if (++sb == SB) { sb = 0; pb ^= 1; }
                if (s == s_stage && nb < nblocks) stage_item(nb, nj, it + 1);
            }
            b = nb; j = nj; nsub = nnsub;
            ++it;
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
#ifdef ST_TC_TRACE_ON
    if (tid == 0 && blockIdx.x == (unsigned)g_tc_dbg) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_tc_trace[5][122] = (long long)gt; g_tc_trace[5][123] = clock64(); }
#endif
}

bool tp_supported(int cin, int cout) { return (cin == 8 || cin == 16 || cin == 32) && cout % 8 == 0 && cout <= 32; }

}  // namespace

extern "C" int64_t st_conv_tc_weight_floats(int ntaps, int cin, int cout) {
    if (!tc_supported(cin, cout)) return -1;
    return (int64_t)tc_nstages(ntaps, cin) * 2 * tc_npad(cout) * TC_KS;
}

extern "C" int st_conv_tc_prepare(const float *w, int ntaps, int cin, int cout, float *wprep, void *stream) {
    ST_REQUIRE(tc_supported(cin, cout), "channel counts not supported by the tensor-core path");
    int npad = tc_npad(cout), nst = tc_nstages(ntaps, cin);
    int64_t total = (int64_t)nst * npad * TC_KS;
    k_tc_prepare<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, ntaps, cin, cout, npad, nst, wprep, nullptr, 0, nullptr, nst);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// Weights of a conv with a FUSED ResBlock identity (out = act(scale * conv(in) + shift + w2 . in2)): the 1x1 identity
// weights w2[cin2, cout], divided by the BN scale, are appended as extra K stages, so the tensor cores compute
// scale * (conv(in) + (w2 / scale) . in2) + shift in one accumulation (st_conv_gather_tc with in2 != NULL, w2 == NULL).
extern "C" int64_t st_conv_tc_weight_floats_fused(int ntaps, int cin, int cout, int cin2) {
    if (!tc_supported(cin, cout) || cin2 <= 0 || cin2 % 16) return -1;
    return (int64_t)(tc_nstages(ntaps, cin) + (cin2 + TC_KS - 1) / TC_KS) * 2 * tc_npad(cout) * TC_KS;
}

extern "C" int st_conv_tc_prepare_fused(const float *w, int ntaps, int cin, int cout, const float *w2, int cin2, const float *scale,
                                        float *wprep, void *stream) {
    ST_REQUIRE(tc_supported(cin, cout) && cin2 > 0 && cin2 % 16 == 0 && w2 != nullptr, "channel counts not supported by the fused tensor-core path");
    int npad = tc_npad(cout), nmain = tc_nstages(ntaps, cin), nst = nmain + (cin2 + TC_KS - 1) / TC_KS;
    int64_t total = (int64_t)nst * npad * TC_KS;
    k_tc_prepare<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, ntaps, cin, cout, npad, nst, wprep, w2, cin2, scale, nmain);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

static int conv_gather_tc_impl(const float *in, int in_ld, const int32_t *map, int64_t n_out, int ntaps, const float *wprep,
                               int cin, int cout, const float *scale, const float *shift, const float *residual, int res_ld,
                               const float *in2, int in2_ld, const float *w2, int cin2, float *out, int out_ld, int act,
                               const int32_t *row_index, const uint32_t *tile_mask, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ST_REQUIRE(!tile_mask || cin <= 64, "tile masks: at most 64 stages");
    if (n_out == 0) return ST_OK;
    ST_REQUIRE(tc_supported(cin, cout), "channel counts not supported by the tensor-core path");
    ST_REQUIRE(map != nullptr, "the tensor-core path needs an explicit gather map");
    ST_REQUIRE(((uintptr_t)in & 15) == 0 && in_ld % 4 == 0 && ((uintptr_t)out & 15) == 0 && out_ld % 4 == 0 && ((uintptr_t)wprep & 15) == 0,
               "16-byte aligned rows required");
    ST_REQUIRE(!residual || (((uintptr_t)residual & 15) == 0 && res_ld % 4 == 0), "residual alignment");
    ST_REQUIRE(!in2 || !w2 || ((uintptr_t)w2 & 15) == 0, "w2 must be 16-byte aligned");
    ST_REQUIRE(!in2 || (((uintptr_t)in2 & 15) == 0 && in2_ld % 4 == 0), "in2 alignment");
    ST_REQUIRE(!in2 || w2 || cin2 % 16 == 0, "fused identity stages need cin2 % 16 == 0");
    const int npad = tc_npad(cout), nmain = tc_nstages(ntaps, cin);
    const int nst = nmain + ((in2 && !w2) ? (cin2 + TC_KS - 1) / TC_KS : 0);      // w2 == NULL: identity stages are inside wprep
    // weight ring: deep enough that the bulk copies are issued ~2000 cycles ahead of their MMAs
    const int b_stage = 2 * npad * TC_KS * 4;
    int sb = (npad <= 32 ? 40 * 1024 : 128 * 1024) / b_stage;
    sb = sb > TC_MAX_BSTAGES ? TC_MAX_BSTAGES : (sb < 2 ? 2 : sb);
    if (sb > nst) sb = nst;
    TcArgs a{in, in_ld, map, (int)n_out, ntaps, wprep, cin, cout, npad, nst, sb, scale, shift, residual, res_ld, in2, in2_ld, w2, cin2, out, out_ld, act, nmain, row_index, tile_mask};
    const int smem = sb * b_stage + 1024;
    static int n_sms = 0;
    if (!n_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev); }
    const int64_t ntiles = cdiv(n_out, TC_M);
    // two CTAs per SM when both fit: 256 TMEM columns each (accumulator + 3 A stages) and ~75 KB of smem
    const int ctas_per_sm = ((tile_mask ? 2 : 1) * npad + TC_STAGES * TC_A_COLS <= 256 && smem <= 80 * 1024) ? 2 : 1;
    const unsigned grid = (unsigned)(ntiles < (int64_t)n_sms * ctas_per_sm ? ntiles : (int64_t)n_sms * ctas_per_sm);   // persistent CTAs
#define ST_TC_CASE(CI)                                                                                              \
    if (cin == CI) {                                                                                                \
        static int smem_set = 0;                                                                                    \
        if (smem > smem_set) {                                                                                      \
            ST_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<CI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
            ST_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<CI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
            smem_set = smem;                                                                                        \
        }                                                                                                           \
        if (tile_mask) k_conv_tc<CI, true><<<grid, TC_THREADS, smem, s>>>(a);                                       \
        else k_conv_tc<CI, false><<<grid, TC_THREADS, smem, s>>>(a);                                                \
        ST_CHECK_LAUNCH();                                                                                          \
        return ST_OK;                                                                                               \
    }
    ST_TC_CASE(8) ST_TC_CASE(16) ST_TC_CASE(32) ST_TC_CASE(64) ST_TC_CASE(128)
#undef ST_TC_CASE
    set_error("st_conv_gather_tc: cin=%d not instantiated", cin);
    return ST_ERR_UNSUPPORTED;
}


// ------------------------------------------------------------------------------------ tile-plan path, host side
extern "C" size_t st_conv_plan_bytes(int64_t n_out) { return PlanLayout(cdiv(n_out > 0 ? n_out : 1, TC_M)).total; }

extern "C" int st_conv_plan_build(const int32_t *map, int64_t n_out, int ntaps, int64_t n_in, void *plan, size_t plan_bytes, void *stream) {
    if (n_out == 0) return ST_OK;
    ST_REQUIRE(map != nullptr && ntaps >= 1 && ntaps < TP_TAPS, "tile plans take an explicit map of at most 27 taps");
    ST_REQUIRE(n_out < (1ll << 31) && n_in < (1ll << 30), "size");
    const int64_t nblocks = cdiv(n_out, TC_M);
    PlanLayout L(nblocks);
    ST_REQUIRE(plan_bytes >= L.total && ((uintptr_t)plan & 255) == 0, "plan buffer too small or misaligned (st_conv_plan_bytes, 256-byte aligned)");
    int end_bit = 1;
    while ((1ll << end_bit) <= n_in) ++end_bit;
    char *p = (char *)plan;
    k_plan_build<<<(unsigned)nblocks, PB_THREADS, 0, (cudaStream_t)stream>>>(map, (int)n_out, ntaps, (int)n_in, end_bit, (int32_t *)p,
                                                                             (uint16_t *)(p + L.lmap_off), (int2 *)(p + L.rows_off));
    ST_CHECK_LAUNCH();
    return ST_OK;
}

extern "C" int st_conv_tp_supported(int ntaps, int cin, int cout) { return ntaps < TP_TAPS && ntaps > 1 && tp_supported(cin, cout) ? 1 : 0; }

extern "C" int st_conv_gather_tp(const float *in, int in_ld, const void *plan, int64_t n_out, int ntaps, const float *wprep,
                                 int cin, int cout, const float *scale, const float *shift, const float *residual, int res_ld,
                                 const float *in2, int in2_ld, const float *w2, int cin2, float *out, int out_ld, int act,
                                 void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_out == 0) return ST_OK;
    ST_REQUIRE(tp_supported(cin, cout) && ntaps < TP_TAPS, "channel counts / taps not supported by the tile-plan path");
    ST_REQUIRE(plan != nullptr && ((uintptr_t)plan & 255) == 0, "plan");
    ST_REQUIRE(((uintptr_t)in & 15) == 0 && in_ld % 4 == 0 && ((uintptr_t)out & 15) == 0 && out_ld % 4 == 0 && ((uintptr_t)wprep & 15) == 0,
               "16-byte aligned rows required");
    ST_REQUIRE(!residual || (((uintptr_t)residual & 15) == 0 && res_ld % 4 == 0), "residual alignment");
    ST_REQUIRE(!in2 || (w2 && ((uintptr_t)w2 & 15) == 0), "in2 needs a 16-byte aligned w2");
    const int npad = tc_npad(cout), nst = tc_nstages(ntaps, cin);
    const int64_t nblocks = cdiv(n_out, TC_M);
    PlanLayout L(nblocks);
    const char *p = (const char *)plan;
    const int b_stage = 2 * npad * TC_KS * 4;
    int sb = 32 * 1024 / b_stage;
    sb = sb > TC_MAX_BSTAGES ? TC_MAX_BSTAGES : (sb < 2 ? 2 : sb);
    if (sb > nst) sb = nst;
    const int nbuf = cin <= 16 ? 2 : 1;
    TpArgs a{in, in_ld, (const int32_t *)p, (const uint16_t *)(p + L.lmap_off), (const int2 *)(p + L.rows_off), (int)n_out, (int)nblocks,
             wprep, cin, cout, npad, nst, sb, scale, shift, residual, res_ld, in2, in2_ld, w2, cin2, out, out_ld, act};
    const int smem = sb * b_stage + nbuf * ((TP_NU + 1) * cin * 4 + TP_LMAP_BYTES) + 1024;
    static int n_sms = 0;
    if (!n_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev); }
    const int ctas_per_sm = (2 * npad + TC_STAGES * TC_A_COLS <= 256 && smem <= 112 * 1024) ? 2 : 1;
    const unsigned grid = (unsigned)(nblocks < (int64_t)n_sms * ctas_per_sm ? nblocks : (int64_t)n_sms * ctas_per_sm);
#define ST_TP_CASE(CI, NB)                                                                                             \
    if (cin == CI) {                                                                                                   \
        static int smem_set = 0;                                                                                       \
        if (smem > smem_set) {                                                                                         \
            ST_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tp<CI, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            smem_set = smem;                                                                                           \
        }                                                                                                              \
        k_conv_tp<CI, NB><<<grid, TC_THREADS, smem, s>>>(a);                                                           \
        ST_CHECK_LAUNCH();                                                                                             \
        return ST_OK;                                                                                                  \
    }
    ST_TP_CASE(8, 2) ST_TP_CASE(16, 2) ST_TP_CASE(32, 1)
#undef ST_TP_CASE
    set_error("st_conv_gather_tp: cin=%d not instantiated", cin);
    return ST_ERR_UNSUPPORTED;
}

extern "C" int st_conv_gather_tc(const float *in, int in_ld, const int32_t *map, int64_t n_out, int ntaps, const float *wprep,
                                 int cin, int cout, const float *scale, const float *shift, const float *residual, int res_ld,
                                 const float *in2, int in2_ld, const float *w2, int cin2, float *out, int out_ld, int act,
                                 void *stream) {
    return conv_gather_tc_impl(in, in_ld, map, n_out, ntaps, wprep, cin, cout, scale, shift, residual, res_ld, in2, in2_ld, w2, cin2,
                               out, out_ld, act, nullptr, nullptr, stream);
}

// ------------------------------------------------------------------------------------ inverse conv on parity-sorted rows
// A fine voxel p receives from coarse voxel o through tap k iff p = 2o - 1 + k, so per axis the tap is fixed by the
// parity of p (even: k = 1; odd: k in {0, 2}): of the 27 taps at most 8 -- 3.4 on average -- can ever be non-empty,
// and WHICH ones depends only on the parity class of p.  st_inverse_plan sorts the rows of the fine level by parity
// class (stable: Z-order is kept inside a class), permutes the `up` map accordingly and records per 128-row tile the
// taps that occur; st_conv_gather_tc_inv then walks only the K stages that have work in each tile (the other
// ~85 % multiplied zeros) and writes every result to its own row through row_index.
__global__ void k_inv_keys(const int4 *__restrict__ coords, int n, uint32_t *__restrict__ keys, int32_t *__restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(coords + i);
    keys[i] = (uint32_t)(((c.y & 1) << 2) | ((c.z & 1) << 1) | (c.w & 1));
    vals[i] = i;
}
__global__ void k_inv_fill(const int32_t *__restrict__ up, const int32_t *__restrict__ row_index, int n, int32_t *__restrict__ up_sorted,
                           uint32_t *__restrict__ tile_mask) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    int v = -1;
    if (j < n) {
        v = __ldg(up + (size_t)t * n + __ldg(row_index + j));
        up_sorted[(size_t)t * n + j] = v;
    }
    // the 32 rows of a warp lie in one 128-row tile
    if (__any_sync(0xffffffffu, v >= 0) && (threadIdx.x & 31) == 0) atomicOr(tile_mask + (j >> 7), 1u << t);
}
static size_t inv_sort_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, 3);
    return b;
}
extern "C" size_t st_inverse_plan_workspace_bytes(int64_t n) { return align_up(inv_sort_bytes(n)) + 2 * align_up(n * 4) + align_up(n * 4) + 1024; }

extern "C" int st_inverse_plan(const int32_t *coords, const int32_t *up, int64_t n, int ntaps, int32_t *row_index, int32_t *up_sorted,
                               uint32_t *tile_mask, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31) && ntaps >= 1 && ntaps <= 32, "sizes");
    Carver cv(workspace, workspace_bytes);
    uint32_t *keys = cv.take<uint32_t>(n), *keys2 = cv.take<uint32_t>(n);
    int32_t *vals = cv.take<int32_t>(n);
    size_t sb = inv_sort_bytes(n);
    void *sort_ws = cv.take<char>(sb);
    if (!cv.ok()) { set_error("st_inverse_plan: workspace too small"); return ST_ERR_WORKSPACE; }
    k_inv_keys<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const int4 *)coords, (int)n, keys, vals);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_ws, sb, keys, keys2, vals, row_index, (int)n, 0, 3, s));
    ST_CHECK_CUDA(cudaMemsetAsync(tile_mask, 0, cdiv(n, TC_M) * sizeof(uint32_t), s));
    dim3 grid((unsigned)cdiv(n, 256), (unsigned)ntaps);
    k_inv_fill<<<grid, 256, 0, s>>>(up, row_index, (int)n, up_sorted, tile_mask);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

// One call for the strided maps AND the inverse plan of a level: `down` as st_strided_maps; the `up` map is produced
// directly in parity-sorted launch order (up_sorted, row_index, tile_mask as st_inverse_plan), never unsorted.
__global__ void k_strided_maps_inv(const int4 *__restrict__ coords, int n, int m, const uint64_t *__restrict__ keys,
                                   const int32_t *__restrict__ vals, uint32_t mask, const int32_t *__restrict__ row_index,
                                   int32_t *__restrict__ down, int32_t *__restrict__ up_sorted, uint32_t *__restrict__ tile_mask) {
    // thread = (launch row j, tap k): rows are walked in parity-sorted order, so up_sorted is written coalesced and the
    // 32 rows of a warp share their 128-row tile (one mask update per warp)
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    int o = -1;
    if (j < n) {
        const int i = __ldg(row_index + j);
        const int4 c = __ldg(coords + i);
        const int tz = c.y + 1 - k / 9, ty = c.z + 1 - (k / 3) % 3, tx = c.w + 1 - k % 3;
        if (!((tz | ty | tx) & 1)) {
            o = hash_lookup(keys, vals, mask, pack_key(c.x, tz >> 1, ty >> 1, tx >> 1));
            if (o >= 0) down[(size_t)k * m + o] = i;
        }
        up_sorted[(size_t)k * n + j] = o;
    }
    if (__any_sync(0xffffffffu, o >= 0) && (threadIdx.x & 31) == 0) atomicOr(tile_mask + (j >> 7), 1u << k);
}
extern "C" size_t st_strided_maps_inv_workspace_bytes(int64_t n) { return st_inverse_plan_workspace_bytes(n); }

extern "C" int st_strided_maps_inv(const int32_t *coords, int64_t n, int64_t n_out, const uint64_t *out_keys, const int32_t *out_vals,
                                   int64_t out_capacity, int32_t *down, int32_t *row_index, int32_t *up_sorted,
                                   uint32_t *tile_mask, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return ST_OK;
    ST_REQUIRE(n < (1ll << 31), "sizes");
    Carver cv(workspace, workspace_bytes);
    uint32_t *keys = cv.take<uint32_t>(n), *keys2 = cv.take<uint32_t>(n);
    int32_t *vals = cv.take<int32_t>(n);
    size_t sb = inv_sort_bytes(n);
    void *sort_ws = cv.take<char>(sb);
    if (!cv.ok()) { set_error("st_strided_maps_inv: workspace too small"); return ST_ERR_WORKSPACE; }
    k_inv_keys<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const int4 *)coords, (int)n, keys, vals);
    ST_CHECK_LAUNCH();
    ST_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_ws, sb, keys, keys2, vals, row_index, (int)n, 0, 3, s));
    ST_CHECK_CUDA(cudaMemsetAsync(tile_mask, 0, cdiv(n, TC_M) * sizeof(uint32_t), s));
    ST_CHECK_CUDA(cudaMemsetAsync(down, 0xFF, (size_t)27 * n_out * sizeof(int32_t), s));
    dim3 grid((unsigned)cdiv(n, 256), 27);
    k_strided_maps_inv<<<grid, 256, 0, s>>>((const int4 *)coords, (int)n, (int)n_out, out_keys, out_vals, (uint32_t)(out_capacity - 1), row_index,
                                            down, up_sorted, tile_mask);
    ST_CHECK_LAUNCH();
    return ST_OK;
}

extern "C" int st_conv_gather_tc_inv(const float *in, int in_ld, const int32_t *up_sorted, const int32_t *row_index, const uint32_t *tile_mask,
                                     int64_t n_out, int ntaps, const float *wprep, int cin, int cout, const float *scale,
                                     const float *shift, float *out, int out_ld, int act, void *stream) {
    ST_REQUIRE(row_index != nullptr && tile_mask != nullptr, "plan of st_inverse_plan required");
    return conv_gather_tc_impl(in, in_ld, up_sorted, n_out, ntaps, wprep, cin, cout, scale, shift, nullptr, 0, nullptr, 0, nullptr, 0,
                               out, out_ld, act, row_index, tile_mask, stream);
}

// debug only (not part of include/st_b200.h)
extern "C" int st_debug_tc_trace(long long *out_host) {
    ST_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_tc_trace, sizeof(long long) * 8 * 128));
    return ST_OK;
}
extern "C" int st_debug_tc_set(int flags) {
    ST_CHECK_CUDA(cudaMemcpyToSymbol(g_tc_dbg, &flags, sizeof(int)));
    return ST_OK;
}
