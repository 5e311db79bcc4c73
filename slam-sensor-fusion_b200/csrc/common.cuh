// common.cuh -- shared declarations of libssf_gpu (sm_100a only; no CPU fallback).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "ssf/ssf.h"

namespace ssf {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<uint64_t> g_queries;

#define SSF_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ssf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return SSF_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define SSF_TRY(expr)              \
    do {                           \
        int r__ = (expr);          \
        if (r__ != SSF_OK) return r__; \
    } while (0)

// count + check a kernel launch
#define SSF_LAUNCHED()                         \
    do {                                       \
        ssf::g_launches.fetch_add(1, std::memory_order_relaxed); \
        SSF_CUDA(cudaGetLastError());          \
    } while (0)

// Growable device buffer (never shrinks; freed with its owner).
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    int reserve(size_t n)
    {
        if (n <= cap) return SSF_OK;
        release();
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return SSF_ERR_NOMEM;
        }
        cap = want;
        return SSF_OK;
    }
};

template <class T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf()
    {
        if (p) cudaFreeHost(p);
    }
    int reserve(size_t n)
    {
        if (n <= cap) return SSF_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMallocHost(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return SSF_ERR_NOMEM;
        }
        cap = want;
        return SSF_OK;
    }
};

// ---------------------------------------------------------------------------------------
// Voxel-hash map index (built by map_build.cu, read by nn_search.cu).
//
// The map cloud is sorted by cell key = morton2(cy, cz) * NX + (cx + 1): all cells of one
// x-row are contiguous and ascending in x, rows follow a 2-D Morton curve over (y, z).
// The hash table holds one entry per cell that is occupied or has an occupied x-neighbour;
// its value is four offsets s0..s3 into the sorted cloud such that cell cx-1 = [s0,s1),
// cx = [s1,s2), cx+1 = [s2,s3) -- one probe yields a contiguous run of up to three cells.
// ---------------------------------------------------------------------------------------
struct MapView {
    const float4 *pts;  // sorted; .w carries the ORIGINAL index (bit pattern of an int)
    const float4 *nrm;  // normals in sorted order, or nullptr
    const unsigned long long *hkeys;
    const uint4 *hvals;
    uint32_t hmask;  // table size - 1 (power of two)
    uint32_t n_pts;  // finite target points
    float ox, oy, oz, inv_h;
    int nx, ny, nz;  // grid extent in cells; NX = nx + 2
    int own_lo, own_hi;  // map sharding: this rank owns queries whose cell column cx is in [own_lo, own_hi)
};

constexpr unsigned long long kEmptyKey = ~0ull;

__host__ __device__ inline uint32_t spread16(uint32_t v)
{
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// cx in [-1, nx], cy in [0, ny), cz in [0, nz)
__host__ __device__ inline unsigned long long cell_key(int cx, int cy, int cz, int nx)
{
    unsigned long long row = (unsigned long long)(spread16((uint32_t)cy) | (spread16((uint32_t)cz) << 1));
    return row * (unsigned long long)(nx + 2) + (unsigned long long)(cx + 1);
}

__host__ __device__ inline uint32_t hash_key(unsigned long long k)
{
    uint32_t x = (uint32_t)k ^ ((uint32_t)(k >> 32) * 0x9E3779B1u);
    x *= 0x85EBCA6Bu;
    x ^= x >> 15;
    x *= 0xC2B2AE35u;
    x ^= x >> 13;
    return x;
}

#ifdef __CUDACC__
// Monotone cell coordinate: every step (sub, mul, clamp, floor) is non-decreasing in v, so
// |a - b| <= r implies cell(a - r) <= cell(b) <= cell(a + r) -- the covering argument of the
// exact search does not depend on how the float operations round.
__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int n)
{
    float u = __fmul_rn(__fsub_rn(v, o), inv_h);
    u = fminf(fmaxf(u, -2.0f), (float)n + 1.0f);
    return (int)floorf(u);
}
// order-preserving float <-> int map, so float min/max can use integer atomics
__device__ __forceinline__ int float_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }
#endif

// ---- primitives.cu -----------------------------------------------------------------------
struct Scratch {  // reusable temporaries of one context
    DevBuf<uint32_t> scan_tmp[3];
    DevBuf<uint32_t> hist;
    DevBuf<unsigned long long> keys_alt;
    DevBuf<uint32_t> vals_alt;
};

// out[i] = sum_{j<i} in[j]; in == out allowed.  total (optional, device) receives the sum.
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *total_dev, Scratch &s,
                       cudaStream_t st);
// Stable LSD radix sort of (key, val) pairs on the low `bits` bits of the keys.  Sorted data
// ends up in keys/vals (ping-pong through scratch).
int radix_sort_pairs_u64(unsigned long long *keys, uint32_t *vals, size_t n, int bits, Scratch &s, cudaStream_t st);
// bbox[0..2] = min xyz, bbox[3..5] = max xyz over finite points (device array of 6 floats);
// n_finite (device) = number of finite points.
int bbox_finite(const float4 *pts, size_t n, float *bbox_dev, uint32_t *n_finite_dev, cudaStream_t st);

}  // namespace ssf
