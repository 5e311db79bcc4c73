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

#ifdef SSF_BOUNDS
// debug build only (make bounds): trap on any index outside its array
#define SSF_CHECK(cond) do { if (!(cond)) { printf("SSF_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define SSF_CHECK(cond) ((void)0)
#endif

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
// Voxel-grid map index (built by map_build.cu, read by nn_device.cuh).
//
// The map is binned into cubic cells of edge h (a few points per occupied cell; h does NOT
// depend on the rejection threshold -- the search walks as many rings of rows as the current
// best distance needs).  A ROW is the set of cells sharing (cy, cz); it is cut into BLOCKS of
// 32 consecutive x cells.  The directory holds one 8-byte entry per (block, row):
//     .x = occupancy mask of the 32 cells, .y = id of the block's first occupied cell
// and is laid out so that the 16 rows of a 4x4 (y, z) tile at the same x block share one
// 128-byte line:  index = (((cz>>2)*nty + (cy>>2))*nbx + bx)*16 + (cz&3)*4 + (cy&3).
// The cloud is sorted by key = index*32 + (cx&31), so the occupied cells of a block -- and any
// x run inside it -- are contiguous:  cells [a, b] of a block hold points
//     [cell_start[.y + popc(mask & below(a))], cell_start[.y + popc(mask & upto(b))]).
// One 8-byte load answers "is anything there" for up to 32 cells; no hashing, no probing.
// ---------------------------------------------------------------------------------------
struct MapView {
    const float4 *pts;  // sorted by key; .w carries the ORIGINAL index (bit pattern of an int)
    const float4 *pn;   // (point, normal) pairs in sorted order -- pn[2 i] == pts[i], pn[2 i + 1] = its normal -- or
                        // nullptr: the point-to-plane sums read both with ONE 256-bit load (ld_point_normal)
    const uint2 *dir;   // directory, ntz*nty*nbx*16 entries
    const uint32_t *cell_start;  // n_cells + 1 offsets into pts
    uint32_t n_pts;  // finite target points
    uint32_t n_cells;  // occupied cells (cell_start has n_cells + 1 entries)
    float ox, oy, oz, inv_h;
    float hq;        // (1 / inv_h) * (1 - 1e-6), rounded down: under-estimate of the cell edge
    int nx, ny, nz;  // grid extent in cells
    int nbx, nty;    // directory: x blocks per row, 4x4 tiles along y
    float bmin[3], bmax[3];  // bounding box of the finite points
    // reach mask (map_build.cu, build_reach_mask): one bit per cell, laid out like dir[].x; a CLEAR bit
    // says that no target point lies within sqrt(reach2) of ANY position binned into that cell, so a
    // search whose bound is below reach2 ends there.  nullptr when not built.
    const uint32_t *reach;
    float reach2;
    float cert_mu;           // margin of the search certificates (nn_device.cuh), a fraction of the cell edge
    float cert_step;         // certificates are written once the last pose update was below this (metres)
    // map sharding: this rank owns the queries whose shard column
    // floor((x - shard_ox) * shard_inv_h) lies in [own_lo, own_hi)
    float shard_ox, shard_inv_h;
    int own_lo, own_hi;
};

__host__ __device__ inline uint32_t dir_index(int nbx, int nty, int bx, int cy, int cz)
{
    return ((((uint32_t)(cz >> 2) * (uint32_t)nty + (uint32_t)(cy >> 2)) * (uint32_t)nbx + (uint32_t)bx) << 4) |
           (uint32_t)(((cz & 3) << 2) | (cy & 3));
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
// 32-byte (point, normal) record of MapView::pn with one 256-bit load (sm_100: LDG.E.ENL2.256): half the L1
// wavefronts of two scattered 128-bit gathers
__device__ __forceinline__ void ld_point_normal(const float4 *rec, float4 &p, float4 &n)
{
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w), "=f"(n.x), "=f"(n.y), "=f"(n.z), "=f"(n.w)
                 : "l"(rec));
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
// Segmented form: n a multiple of kSortTileSize; seg_of_tile[t] = (first sort tile of t's segment, tiles in it).
// Every segment is sorted on its own (all 32 key bits), in place.
constexpr int kSortTileSize = 2048;
// segs[i] = (first sort tile, tiles, base offset of the segment's slot range, unused), n_segs segments.
int seg_radix_sort_pairs_u32(uint32_t *keys, uint32_t *vals, size_t n, const uint2 *seg_of_tile, const uint4 *segs,
                             uint32_t n_segs, Scratch &s, cudaStream_t st);
// in-place exclusive scan of per-tile counts (entries_per_tile per sort tile), every segment on its
// own and starting at its base offset
int seg_scan_u32(uint32_t *counts, const uint4 *segs, uint32_t n_segs, uint32_t entries_per_tile, cudaStream_t st);
// bbox[0..2] = min xyz, bbox[3..5] = max xyz over finite points (device array of 6 floats);
// n_finite (device) = number of finite points.
int bbox_finite(const float4 *pts, size_t n, float *bbox_dev, uint32_t *n_finite_dev, cudaStream_t st);

}  // namespace ssf
