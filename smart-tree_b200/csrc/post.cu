// Skeleton finishing on the device: node gather + prune + repair + smooth for every component in ONE
// launch (one CTA per connected component), writing one packed buffer the host copies back once.
//   gather  smart_tree/skeleton/path.py:118-129 (branch = medial points / radii of its path vertices)
//   prune   smart_tree/data_types/tree.py:94-121 (first skeleton only, tree.py:164-168)
//   repair  smart_tree/data_types/tree.py:73-92 + util/queries.py:89-133
//   smooth  smart_tree/data_types/tree.py:123-134 (zero-padded box filter on the radii)
#include <limits.h>

#include "common.cuh"

namespace st {
namespace {

constexpr int FLAG_KEEP = 1, FLAG_CONN = 2, FLAG_SMOOTH = 4;

struct FinishArgs {
    const float *pts, *radii;
    const int32_t *comp_off;
    int n_comp;
    const int32_t *path, *blen, *bpar, *cnb, *cnp;
    int prune_first;
    float min_radius, min_length;
    int repair, smooth_k;
    int32_t *depth;      // workspace [n]: tree depth of a branch below its first un-repaired ancestor
    int32_t *out;        // packed: header[4] | bmeta[B][4] | nodes[R][4] | smooth[R]
};

__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[31];
    const int r = x - v + (warp ? s_warp[warp - 1] : 0);
    __syncthreads();
    return r;
}

constexpr int FIN_CACHE = 4096;      // branches of a component whose flags / depths the serial pass keeps in shared memory

__global__ void __launch_bounds__(1024) k_finish(FinishArgs a) {
    const int c = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    __shared__ int s_warp[32];
    __shared__ int s_bbase, s_rbase, s_btot, s_rtot, s_maxdepth;
    if (tid == 0) {
        int bb = 0, rb = 0, bt = 0, rt = 0;
        for (int k = 0; k < a.n_comp; ++k) {
            const int nb = a.cnb[k], np = a.cnp[k];
            if (k < c) { bb += nb; rb += nb + np; }
            bt += nb; rt += nb + np;
        }
        s_bbase = bb; s_rbase = rb; s_btot = bt; s_rtot = rt; s_maxdepth = 0;
        if (c == 0) { a.out[0] = bt; a.out[1] = rt; a.out[2] = 0; a.out[3] = 0; }
    }
    __syncthreads();
    const int seg = a.comp_off[c];
    const int nb = a.cnb[c];
    int32_t *const bmeta = a.out + 4 + 4 * (size_t)s_bbase;
    float *const nodes_all = (float *)(a.out + 4 + 4 * (size_t)s_btot);
    float *const smooth_all = nodes_all + 4 * (size_t)s_rtot;
    const int32_t *const blen = a.blen + seg, *const bpar = a.bpar + seg, *const path = a.path + seg;
    int32_t *const depth = a.depth + seg;
    const int rbase = s_rbase;

    // ---- 1. spare-row index of every branch: rbase + (nodes of earlier branches) + (earlier branches)
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
        const int b = b0 + tid;
        const int v = b < nb ? blen[b] : 0;
        int total;
        const int ex = block_excl_scan(v, s_warp, total);
        if (b < nb) {
            bmeta[4 * b] = rbase + carry + ex + b;
            bmeta[4 * b + 1] = v;
            bmeta[4 * b + 2] = bpar[b];
            bmeta[4 * b + 3] = FLAG_KEEP;
        }
        carry += total;
    }
    __syncthreads();
    // ---- 2. gather the nodes (one warp per branch); branch length + end radii for the prune rule
    const bool prune = a.prune_first && c == 0;
    for (int b = warp; b < nb; b += nwarp) {
        const int row = bmeta[4 * b], len = bmeta[4 * b + 1];
        const int p0 = row - rbase - b;                  // position of the branch's first vertex in the path list
        double acc = 0.0;
        for (int i0 = 0; i0 < len; i0 += 32) {
            const int i = i0 + lane;
            float x = 0.f, y = 0.f, z = 0.f, r = 0.f;
            if (i < len) {
                const int v = seg + path[p0 + i];
                x = a.pts[3 * (size_t)v]; y = a.pts[3 * (size_t)v + 1]; z = a.pts[3 * (size_t)v + 2]; r = a.radii[v];
                float *o = nodes_all + 4 * (size_t)(row + 1 + i);
                o[0] = x; o[1] = y; o[2] = z; o[3] = r;
                smooth_all[row + 1 + i] = r;
                if (i == 0) {                            // spare row starts as a copy of the first node
                    o -= 4;
                    o[0] = x; o[1] = y; o[2] = z; o[3] = r;
                    smooth_all[row] = r;
                }
            }
            if (prune) {
                // segment i -> i+1: the neighbour's position comes from the next lane (or the next batch's lane 0)
                float nx = __shfl_down_sync(0xffffffffu, x, 1), ny = __shfl_down_sync(0xffffffffu, y, 1), nz = __shfl_down_sync(0xffffffffu, z, 1);
                if (lane == 31 && i + 1 < len) {
                    const int v2 = seg + path[p0 + i + 1];
                    nx = a.pts[3 * (size_t)v2]; ny = a.pts[3 * (size_t)v2 + 1]; nz = a.pts[3 * (size_t)v2 + 2];
                }
                if (i + 1 < len) {
                    const float dx = nx - x, dy = ny - y, dz = nz - z;
                    acc += (double)sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                }
            }
        }
        if (prune) {
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                const float length = (float)acc;
                const float r0 = a.radii[seg + path[p0]], r1 = a.radii[seg + path[p0 + (len > 0 ? len - 1 : 0)]];
                const bool ok = !(length < a.min_length) && !(fmaxf(r0, r1) < a.min_radius);
                depth[b] = ok ? 1 : 0;                   // provisional: the branch passes its own thresholds
            }
        }
    }
    __syncthreads();
    // ---- 3. keep flags (parents precede children) and repair depth, one thread, emission order.  The loop is a chain of
    //         dependent reads of the parent's flags and depth: from shared memory when the component's branch table fits
    //         (406 branches of the bench tree: ~0.2 ms of dependent L2 round trips otherwise)
    __shared__ int s_par[FIN_CACHE];
    __shared__ short s_dep[FIN_CACHE];
    __shared__ unsigned char s_flg[FIN_CACHE];
    const bool cached = nb <= FIN_CACHE;
    if (cached) {
        for (int b = tid; b < nb; b += blockDim.x) { s_par[b] = bmeta[4 * b + 2]; s_dep[b] = (short)(prune ? depth[b] : 0); }
        __syncthreads();
    }
    if (tid == 0) {
        int maxd = 0;
        for (int b = 0; b < nb; ++b) {
            const int par = cached ? s_par[b] : bmeta[4 * b + 2];
            const bool par_listed = par >= 0 && par < nb && par != b;
            const int pflags = (par_listed && par < b) ? (cached ? (int)s_flg[par] : bmeta[4 * par + 3]) : 0;
            int flags = FLAG_KEEP;
            if (prune) {
                // tree.py:105-119: the root (smallest id) always stays; others need a kept parent and pass the thresholds
                const bool par_kept = par_listed && par < b && (pflags & FLAG_KEEP);
                const int ok = cached ? (int)s_dep[b] : depth[b];
                if (b != 0 && !(par_kept && ok)) flags = 0;
            }
            int d = 0;
            if (a.repair && flags && par_listed && par < b && (pflags & FLAG_KEEP)) {
                flags |= FLAG_CONN;
                d = ((pflags & FLAG_CONN) ? (cached ? (int)s_dep[par] : depth[par]) : 0) + 1;
            }
            if (cached) { s_flg[b] = (unsigned char)flags; s_dep[b] = (short)d; }
            else { bmeta[4 * b + 3] = flags; depth[b] = d; }
            maxd = d > maxd ? d : maxd;
        }
        s_maxdepth = maxd;
    }
    __syncthreads();
    if (cached) {
        for (int b = tid; b < nb; b += blockDim.x) { bmeta[4 * b + 3] = s_flg[b]; depth[b] = s_dep[b]; }
        __syncthreads();
    }
    // ---- 4. repair, level by level: connection point = nearest-tube projection of the first node onto the
    //         parent's current polyline (spare row included iff the parent has been repaired)
    const int maxd = s_maxdepth;
    for (int lv = 1; lv <= maxd; ++lv) {
        for (int b = warp; b < nb; b += nwarp) {
            if (depth[b] != lv || !(bmeta[4 * b + 3] & FLAG_CONN)) continue;
            const int pb = bmeta[4 * b + 2];
            const int r0 = bmeta[4 * b] + 1;
            const float *nd = nodes_all;
            const float px = nd[4 * (size_t)r0], py = nd[4 * (size_t)r0 + 1], pz = nd[4 * (size_t)r0 + 2];
            const int prow = bmeta[4 * pb];
            const int t0 = (bmeta[4 * pb + 3] & FLAG_CONN) ? prow : prow + 1;
            const int t1 = prow + bmeta[4 * pb + 1];
            float best = FLT_MAX * 2.f, bx = 0, by = 0, bz = 0;     // +inf
            int bidx = INT_MAX;
#pragma unroll 4      // (the loads of four segments in flight: the polyline was written by this kernel and comes from L2)
            for (int m = t0 + lane; m < t1; m += 32) {
                const float4 A = __ldcg((const float4 *)(nd + 4 * (size_t)m)), B = __ldcg((const float4 *)(nd + 4 * (size_t)m + 4));
                float abx = B.x - A.x, aby = B.y - A.y, abz = B.z - A.z;
                float apx = px - A.x, apy = py - A.y, apz = pz - A.z;
                float t = fminf(fmaxf((apx * abx + apy * aby + apz * abz) / (abx * abx + aby * aby + abz * abz), 0.f), 1.f);
                float dx = A.x + t * abx - px, dy = A.y + t * aby - py, dz = A.z + t * abz - pz;
                float score = fabsf(sqrtf(dx * dx + dy * dy + dz * dz) - ((1.f - t) * A.w + t * B.w));
                if (score < best) { best = score; bidx = m; bx = dx; by = dy; bz = dz; }
            }
            for (int o = 16; o; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                float ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
                if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bx = ox; by = oy; bz = oz; }
            }
            if (lane == 0) {
                float *o = nodes_all + 4 * (size_t)(r0 - 1);
                o[0] = px + bx; o[1] = py + by; o[2] = pz + bz;      // radius already = first node's (tree.py:92)
            }
        }
        __syncthreads();
    }
    // ---- 5. smooth: box filter of width k over each kept branch longer than k nodes, window clipped to the branch
    if (a.smooth_k > 0) {
        const int h = a.smooth_k / 2;
        const float inv_scale = (float)a.smooth_k;
        for (int b = warp; b < nb; b += nwarp) {
            const int flags = bmeta[4 * b + 3];
            if (!(flags & FLAG_KEEP)) continue;
            const int first = bmeta[4 * b] + ((flags & FLAG_CONN) ? 0 : 1);
            const int last = bmeta[4 * b] + bmeta[4 * b + 1];
            const int cnt = last - first + 1;
            if (cnt <= a.smooth_k) continue;
            for (int i = first + lane; i <= last; i += 32) {
                const int lo = max(i - h, first), hi = min(i + h, last);
                float sacc = 0.f;
                for (int j = lo; j <= hi; ++j) sacc += nodes_all[4 * (size_t)j + 3];
                smooth_all[i] = sacc / inv_scale;
            }
            if (lane == 0) bmeta[4 * b + 3] = flags | FLAG_SMOOTH;
        }
    }
}

}  // namespace
}  // namespace st

using namespace st;

extern "C" size_t st_finish_skeletons_out_ints(int64_t n) {
    // header + worst case of n branches of one node each: 4 ints of metadata + 2 rows of 5 values per branch
    return 4 + (size_t)n * 4 + (size_t)2 * n * 5;
}

extern "C" int st_finish_skeletons(const float *medial_pts, const float *radii, const int32_t *comp_off, int32_t n_comp, int64_t n,
                                   const int32_t *path_vertices, const int32_t *branch_len, const int32_t *branch_parent,
                                   const int32_t *comp_n_branches, const int32_t *comp_n_path, int prune_first, float min_radius,
                                   float min_length, int repair, int smooth_kernel, int32_t *depth_workspace, int32_t *out,
                                   void *stream) {
    if (n_comp <= 0) return ST_OK;
    ST_REQUIRE(smooth_kernel == 0 || (smooth_kernel > 0 && (smooth_kernel & 1)), "smooth_kernel must be 0 or odd");
    ST_REQUIRE(depth_workspace && out, "workspace / out");
    ST_REQUIRE(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    FinishArgs a{medial_pts, radii, comp_off, n_comp, path_vertices, branch_len, branch_parent, comp_n_branches, comp_n_path,
                 prune_first, min_radius, min_length, repair, smooth_kernel, depth_workspace, out};
    k_finish<<<(unsigned)n_comp, 1024, 0, (cudaStream_t)stream>>>(a);
    ST_CHECK_LAUNCH();
    return ST_OK;
}
