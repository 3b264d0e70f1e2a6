// Shared helpers for libst_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>

#include "../../include/st_b200.h"

namespace st {

void set_error(const char *fmt, ...);

#define ST_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            st::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ST_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define ST_CHECK_LAUNCH()                                                                \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            st::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return ST_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define ST_REQUIRE(cond, msg)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            st::set_error("%s:%d requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
            return ST_ERR_ARG;                                                           \
        }                                                                                \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// carve typed pointers out of a caller-provided workspace
struct Carver {
    char *base;
    size_t off = 0, cap;
    Carver(void *p, size_t bytes) : base((char *)p), cap(bytes) {}
    template <typename T> T *take(size_t count) {
        size_t o = align_up(off);
        off = o + count * sizeof(T);
        return (T *)(base + o);
    }
    bool ok() const { return off <= cap; }
};

// ---- 64-bit packed voxel key: b<<48 | (z+1)<<32 | (y+1)<<16 | (x+1); sorts as (b,z,y,x)
// Fields are masked to their width; the legal range (checked once per coordinate set by st_hash_build) is
// 0 <= z,y,x <= ST_MAX_COORD (so that the -1 / +1 neighbours of a voxel still fit) and 0 <= b < ST_MAX_BATCH.
constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned ST_MAX_COORD = 65533u;
constexpr unsigned ST_MAX_BATCH = 32768u;

__host__ __device__ __forceinline__ uint64_t pack_key(int b, int z, int y, int x) {
    return ((uint64_t)((uint32_t)b & 0xFFFFu) << 48) | ((uint64_t)((uint32_t)(z + 1) & 0xFFFFu) << 32) |
           ((uint64_t)((uint32_t)(y + 1) & 0xFFFFu) << 16) | (uint64_t)((uint32_t)(x + 1) & 0xFFFFu);
}
__host__ __device__ __forceinline__ void unpack_key(uint64_t k, int &b, int &z, int &y, int &x) {
    b = (int)(k >> 48);
    z = (int)((k >> 32) & 0xFFFF) - 1;
    y = (int)((k >> 16) & 0xFFFF) - 1;
    x = (int)(k & 0xFFFF) - 1;
}
// Morton (Z-order) variant of the key: batch in the top 16 bits, then the bits of (z+1, y+1, x+1)
// interleaved.  Sorting by it keeps spatial neighbours close in memory (gather locality in L1/L2).
__host__ __device__ __forceinline__ uint64_t spread3(uint64_t v) {   // 16 bits -> every third bit
    v &= 0xFFFFull;
    v = (v | (v << 16)) & 0x0000FF0000FFull;
    v = (v | (v << 8)) & 0x00F00F00F00Full;
    v = (v | (v << 4)) & 0x0C30C30C30C3ull;
    v = (v | (v << 2)) & 0x249249249249ull;
    return v;
}
__host__ __device__ __forceinline__ uint64_t compact3(uint64_t v) {
    v &= 0x249249249249ull;
    v = (v | (v >> 2)) & 0x0C30C30C30C3ull;
    v = (v | (v >> 4)) & 0x00F00F00F00Full;
    v = (v | (v >> 8)) & 0x0000FF0000FFull;
    v = (v | (v >> 16)) & 0xFFFFull;
    return v;
}
__host__ __device__ __forceinline__ uint64_t morton_key(int b, int z, int y, int x) {
    return ((uint64_t)(uint32_t)b << 48) | (spread3((uint32_t)(z + 1)) << 2) | (spread3((uint32_t)(y + 1)) << 1) | spread3((uint32_t)(x + 1));
}
__host__ __device__ __forceinline__ void unpack_morton(uint64_t k, int &b, int &z, int &y, int &x) {
    b = (int)(k >> 48);
    z = (int)compact3(k >> 2) - 1;
    y = (int)compact3(k >> 1) - 1;
    x = (int)compact3(k) - 1;
}
__device__ __forceinline__ uint32_t hash_key(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}
__device__ __forceinline__ int hash_lookup(const uint64_t *__restrict__ keys,
                                           const int32_t *__restrict__ vals, uint32_t mask,
                                           uint64_t key) {
    uint32_t s = hash_key(key) & mask;
    while (true) {
        uint64_t k = __ldg(keys + s);
        if (k == key) return __ldg(vals + s);
        if (k == KEY_EMPTY) return -1;
        s = (s + 1) & mask;
    }
}

// exact fp32 squared distance, no FMA contraction: (dx*dx + dy*dy) + dz*dz
__device__ __forceinline__ float dist2_exact(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace st
