// voxel_grid.cu -- K2: voxel-grid downsample with pcl::VoxelGrid<PointXYZ> semantics.
//
// Replaces vg.setLeafSize(l, l, l); vg.filter(*cloud) at reference
// localization/src/global_map_frames_manager.cpp:143-146 (and the north-star's scan
// downsample).  PCL's algorithm (voxel_grid.hpp, restated in SURVEY.md Appendix B.3):
//   min/max over finite points -> min_b = floor(min * inv_leaf), div_b -> per point
//   idx = (floor(p * inv_leaf) - min_b) . (1, dx, dx*dy) -> sort by idx -> one centroid per
//   run, float running sum in run order divided by the count, output in ascending idx.
// PCL's sort is not stable, so its in-voxel order is undefined; the contract here (and in the
// oracle) is ascending original index, which a STABLE radix sort gives.  Each run is summed
// by one thread in order, so the float centroid is bit-identical to a sequential CPU loop.
#include <cfloat>
#include <climits>
#include <cmath>

#include "voxel_grid.cuh"

namespace ssf {

__global__ void __launch_bounds__(256)
    voxel_keys_kernel(const float4 *__restrict__ in, uint32_t n, float inv, int mbx, int mby, int mbz, int dvx, int dvy,
                      unsigned long long sentinel, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    unsigned long long k = sentinel;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        // static_cast<int>(std::floor(p.x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0]))
        const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)mbx);
        const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)mby);
        const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)mbz);
        const int idx = i0 + i1 * dvx + i2 * dvx * dvy;
        k = (unsigned long long)(uint32_t)idx;
    }
    keys[i] = k;
    vals[i] = i;
}

__global__ void __launch_bounds__(256)
    voxel_flags_kernel(const unsigned long long *__restrict__ keys, uint32_t n_finite, uint32_t *__restrict__ flags)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_finite) return;
    flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(128)
    voxel_centroid_kernel(const float4 *__restrict__ in, const unsigned long long *__restrict__ keys,
                          const uint32_t *__restrict__ vals, const uint32_t *__restrict__ flags,
                          const uint32_t *__restrict__ scan, uint32_t n_finite, float4 *__restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_finite || !flags[j]) return;
    const unsigned long long k = keys[j];
    float cx = 0.f, cy = 0.f, cz = 0.f;
    uint32_t e = j;
    while (e < n_finite && keys[e] == k) {
        const float4 p = in[vals[e]];
        cx = __fadd_rn(cx, p.x);
        cy = __fadd_rn(cy, p.y);
        cz = __fadd_rn(cz, p.z);
        ++e;
    }
    const float cnt = (float)(e - j);
    out[scan[j]] = make_float4(__fdiv_rn(cx, cnt), __fdiv_rn(cy, cnt), __fdiv_rn(cz, cnt), 1.0f);
}

static int bit_width_u64(unsigned long long v)
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

// ---- Open3D semantics (the Python node's pcd.voxel_down_sample, localization_node.py:47) -----------------
// open3d::geometry::PointCloud::VoxelDownSample [ext]: voxel_min_bound = min_bound - voxel_size / 2,
// index = floor((p - voxel_min_bound) / voxel_size) per axis, everything in double; one AccumulatedPoint per
// occupied voxel, points added in input order, centroid = sum / count in double.  Open3D returns the voxels
// in unordered_map iteration order (unspecified): the contract here is ascending (z, y, x) voxel index.
__global__ void __launch_bounds__(256)
    voxel_keys_o3d_kernel(const float4 *__restrict__ in, uint32_t n, double bx, double by, double bz, double v,
                          unsigned long long dvx, unsigned long long dvy, unsigned long long sentinel,
                          unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    unsigned long long k = sentinel;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const long long i0 = (long long)floor(__ddiv_rn(__dsub_rn((double)p.x, bx), v));
        const long long i1 = (long long)floor(__ddiv_rn(__dsub_rn((double)p.y, by), v));
        const long long i2 = (long long)floor(__ddiv_rn(__dsub_rn((double)p.z, bz), v));
        k = (unsigned long long)i0 + (unsigned long long)i1 * dvx + (unsigned long long)i2 * dvx * dvy;
    }
    keys[i] = k;
    vals[i] = i;
}

__global__ void __launch_bounds__(128)
    voxel_centroid_o3d_kernel(const float4 *__restrict__ in, const unsigned long long *__restrict__ keys,
                              const uint32_t *__restrict__ vals, const uint32_t *__restrict__ flags,
                              const uint32_t *__restrict__ scan, uint32_t n_finite, float4 *__restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_finite || !flags[j]) return;
    const unsigned long long k = keys[j];
    double cx = 0.0, cy = 0.0, cz = 0.0;
    uint32_t e = j;
    while (e < n_finite && keys[e] == k) {
        const float4 p = in[vals[e]];
        cx = __dadd_rn(cx, (double)p.x);
        cy = __dadd_rn(cy, (double)p.y);
        cz = __dadd_rn(cz, (double)p.z);
        ++e;
    }
    const double cnt = (double)(e - j);
    out[scan[j]] = make_float4((float)__ddiv_rn(cx, cnt), (float)__ddiv_rn(cy, cnt), (float)__ddiv_rn(cz, cnt), 1.0f);
}

int voxel_downsample_o3d_device(VoxelWork &w, size_t n, double voxel, Scratch &s, cudaStream_t st, uint32_t *n_out)
{
    *n_out = 0;
    if (n == 0) return SSF_OK;
    SSF_TRY(w.small.reserve(16));
    float *bbox_dev = w.small.p;
    uint32_t *cnt_dev = reinterpret_cast<uint32_t *>(w.small.p + 8);
    SSF_TRY(bbox_finite(w.in.p, n, bbox_dev, cnt_dev, st));
    float hb[6];
    uint32_t n_finite = 0;
    SSF_CUDA(cudaMemcpyAsync(hb, bbox_dev, sizeof(hb), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaMemcpyAsync(&n_finite, cnt_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    if (n_finite == 0) return SSF_OK;
    double b[3];
    unsigned long long d[3];
    for (int k = 0; k < 3; ++k) {
        b[k] = (double)hb[k] - voxel * 0.5;
        d[k] = (unsigned long long)floor(((double)hb[3 + k] - b[k]) / voxel) + 1ull;
    }
    const double cells = (double)d[0] * (double)d[1] * (double)d[2];
    if (!(cells < 9.0e18)) {
        set_error("voxel_down_sample: voxel size %g is too small for the cloud's extent", voxel);
        return SSF_ERR_INVALID;
    }
    const unsigned long long sentinel = d[0] * d[1] * d[2];
    SSF_TRY(w.keys.reserve(n));
    SSF_TRY(w.vals.reserve(n));
    SSF_TRY(w.flags.reserve(n));
    SSF_TRY(w.scan.reserve(n));
    voxel_keys_o3d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.in.p, (uint32_t)n, b[0], b[1], b[2], voxel, d[0], d[1],
                                                                       sentinel, w.keys.p, w.vals.p);
    SSF_LAUNCHED();
    SSF_TRY(radix_sort_pairs_u64(w.keys.p, w.vals.p, n, bit_width_u64(sentinel), s, st));
    voxel_flags_kernel<<<(n_finite + 255) / 256, 256, 0, st>>>(w.keys.p, n_finite, w.flags.p);
    SSF_LAUNCHED();
    SSF_TRY(exclusive_scan_u32(w.flags.p, w.scan.p, n_finite, cnt_dev + 1, s, st));
    voxel_centroid_o3d_kernel<<<(n_finite + 127) / 128, 128, 0, st>>>(w.in.p, w.keys.p, w.vals.p, w.flags.p, w.scan.p,
                                                                     n_finite, w.out.p);
    SSF_LAUNCHED();
    uint32_t cnt = 0;
    SSF_CUDA(cudaMemcpyAsync(&cnt, cnt_dev + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    *n_out = cnt;
    return SSF_OK;
}

int voxel_downsample_device(VoxelWork &w, size_t n, float leaf, Scratch &s, cudaStream_t st, uint32_t *n_out,
                            int *refused)
{
    *n_out = 0;
    *refused = 0;
    if (n == 0) return SSF_OK;
    SSF_TRY(w.small.reserve(16));
    float *bbox_dev = w.small.p;
    uint32_t *cnt_dev = reinterpret_cast<uint32_t *>(w.small.p + 8);
    SSF_TRY(bbox_finite(w.in.p, n, bbox_dev, cnt_dev, st));
    float hb[6];
    uint32_t n_finite = 0;
    SSF_CUDA(cudaMemcpyAsync(hb, bbox_dev, sizeof(hb), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaMemcpyAsync(&n_finite, cnt_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    if (n_finite == 0) return SSF_OK;
    const float inv = 1.0f / leaf;
    // overflow guard of PCL (voxel_grid.hpp): dx*dy*dz > INT32_MAX -> output = input
    long long d64[3];
    for (int k = 0; k < 3; ++k) d64[k] = (long long)((hb[3 + k] - hb[k]) * inv) + 1;
    if (d64[0] * d64[1] * d64[2] > (long long)INT32_MAX) {
        *refused = 1;
        SSF_CUDA(cudaMemcpyAsync(w.out.p, w.in.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        *n_out = (uint32_t)n;
        return SSF_OK;
    }
    int minb[3], divb[3];
    for (int k = 0; k < 3; ++k) {
        minb[k] = (int)floorf(hb[k] * inv);
        const int maxb = (int)floorf(hb[3 + k] * inv);
        divb[k] = maxb - minb[k] + 1;
    }
    const unsigned long long cells = (unsigned long long)divb[0] * divb[1] * divb[2];
    const unsigned long long sentinel = cells;  // above every valid index
    SSF_TRY(w.keys.reserve(n));
    SSF_TRY(w.vals.reserve(n));
    SSF_TRY(w.flags.reserve(n));
    SSF_TRY(w.scan.reserve(n));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    voxel_keys_kernel<<<blocks, 256, 0, st>>>(w.in.p, (uint32_t)n, inv, minb[0], minb[1], minb[2], divb[0], divb[1],
                                              sentinel, w.keys.p, w.vals.p);
    SSF_LAUNCHED();
    SSF_TRY(radix_sort_pairs_u64(w.keys.p, w.vals.p, n, bit_width_u64(sentinel), s, st));
    const unsigned blocks_f = (n_finite + 255) / 256;
    voxel_flags_kernel<<<blocks_f, 256, 0, st>>>(w.keys.p, n_finite, w.flags.p);
    SSF_LAUNCHED();
    SSF_TRY(exclusive_scan_u32(w.flags.p, w.scan.p, n_finite, cnt_dev + 1, s, st));
    voxel_centroid_kernel<<<(n_finite + 127) / 128, 128, 0, st>>>(w.in.p, w.keys.p, w.vals.p, w.flags.p, w.scan.p,
                                                                 n_finite, w.out.p);
    SSF_LAUNCHED();
    uint32_t cnt = 0;
    SSF_CUDA(cudaMemcpyAsync(&cnt, cnt_dev + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    *n_out = cnt;
    return SSF_OK;
}

}  // namespace ssf

// =========================================================================================
// batched form (scan downsample in front of the ICP loop)
// =========================================================================================
namespace ssf {

// vbox per scan: 6 ordered ints (min xyz, max xyz), [6] finite count, [7] pad
__global__ void vb_init_kernel(int *vbox, uint32_t n_scans)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    for (int k = 0; k < 3; ++k) {
        vbox[8 * s + k] = float_ordered(FLT_MAX);
        vbox[8 * s + 3 + k] = float_ordered(-FLT_MAX);
    }
    vbox[8 * s + 6] = 0;
    vbox[8 * s + 7] = 0;
}

// one block per sort tile (kSlotAlign consecutive slots of one scan), eight points per thread
__global__ void __launch_bounds__(256)
    vb_bbox_kernel(const float4 *__restrict__ raw, const uint32_t *__restrict__ tile_scan,
                   const uint32_t *__restrict__ meta, int *vbox)
{
    const uint32_t base = blockIdx.x * kSlotAlign;
    const uint32_t s = tile_scan[base / kTile];
    const uint32_t n = meta[5 * s + 1], pt_begin = meta[5 * s + 2];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < kSlotAlign / 256; ++i) {
        const uint32_t slot = base + (uint32_t)i * 256 + threadIdx.x;
        if (slot - pt_begin < n) {
            const float4 p = raw[slot];
            if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
                mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
                mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
                ++cnt;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    // one atomic set per block (a tile belongs to one scan), not per warp
    __shared__ float smn[8][3], smx[8][3];
    __shared__ uint32_t scnt[8];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { smn[warp][k] = mn[k]; smx[warp][k] = mx[k]; }
        scnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (int w = 0; w < 8; ++w) {
            c += scnt[w];
#pragma unroll
            for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], smn[w][k]); mx[k] = fmaxf(mx[k], smx[w][k]); }
        }
        if (c) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                atomicMin(&vbox[8 * s + k], float_ordered(mn[k]));
                atomicMax(&vbox[8 * s + 3 + k], float_ordered(mx[k]));
            }
            atomicAdd(reinterpret_cast<uint32_t *>(&vbox[8 * s + 6]), c);
        }
    }
}

// vgrid per scan: minb[3], divb[3], refused, pad
__global__ void vb_grid_kernel(const int *__restrict__ vbox, int32_t *__restrict__ vgrid, uint32_t n_scans, float inv)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    int32_t *g = vgrid + 8 * s;
    for (int k = 0; k < 8; ++k) g[k] = 0;
    if (vbox[8 * s + 6] == 0) { g[3] = g[4] = g[5] = 1; return; }
    long long d64[3];
    for (int k = 0; k < 3; ++k) {
        const float mn = ordered_float(vbox[8 * s + k]), mx = ordered_float(vbox[8 * s + 3 + k]);
        d64[k] = (long long)__fmul_rn(__fsub_rn(mx, mn), inv) + 1;
        g[k] = (int)floorf(__fmul_rn(mn, inv));
        g[3 + k] = (int)floorf(__fmul_rn(mx, inv)) - g[k] + 1;
    }
    if (d64[0] * d64[1] * d64[2] > (long long)INT32_MAX) g[6] = 1;  // PCL refuses: output = input
}

// per sort tile: (first sort tile of its scan, sort tiles of that scan)
__global__ void vb_seg_kernel(const uint32_t *__restrict__ tile_scan, const uint32_t *__restrict__ meta,
                              uint32_t n_sort_tiles, uint2 *__restrict__ seg)
{
    constexpr uint32_t kPer = kSlotAlign / kTile;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_sort_tiles) return;
    const uint32_t s = tile_scan[t * kPer];
    seg[t] = make_uint2(meta[5 * s + 3] / kPer, meta[5 * s + 4] / kPer);
}

constexpr uint32_t kDeadKey = 0xFFFFFFFFu;  // padding / dropped points: sort to the end of their scan

__global__ void __launch_bounds__(256)
    vb_keys_kernel(const float4 *__restrict__ raw, const uint32_t *__restrict__ tile_scan,
                   const uint32_t *__restrict__ meta, const int32_t *__restrict__ vgrid, float inv,
                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t n_slots)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const uint32_t s = tile_scan[slot / kTile];
    const uint32_t n = meta[5 * s + 1], pt_begin = meta[5 * s + 2];
    const uint32_t row = slot - pt_begin;
    uint32_t k = kDeadKey;
    if (row < n) {
        const int32_t *g = vgrid + 8 * s;
        const float4 p = raw[slot];
        if (g[6]) {
            k = row;  // refused: every point is its own voxel
        } else if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)g[0]);
            const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)g[1]);
            const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)g[2]);
            k = (uint32_t)(i0 + i1 * g[3] + i2 * g[3] * g[4]);  // < 2^31 (PCL's overflow guard)
        }
    }
    keys[slot] = k;
    vals[slot] = slot;
}

__global__ void vb_segs_kernel(const uint32_t *__restrict__ meta, uint32_t n_scans, uint4 *__restrict__ segs)
{
    constexpr uint32_t kPer = kSlotAlign / kTile;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    segs[s] = make_uint4(meta[5 * s + 3] / kPer, meta[5 * s + 4] / kPer, 0u, 0u);  // run ids restart at 0 per scan
}

__global__ void vb_sortsegs_kernel(const uint32_t *__restrict__ meta, uint32_t n_scans, uint4 *__restrict__ segs)
{
    constexpr uint32_t kPer = kSlotAlign / kTile;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    segs[s] = make_uint4(meta[5 * s + 3] / kPer, meta[5 * s + 4] / kPer, meta[5 * s + 2], 0u);  // sorted in place
}

// is sorted element j (key k) the first of its voxel run?
__device__ __forceinline__ bool run_head(const uint32_t *__restrict__ keys, uint32_t j, uint32_t k, uint32_t seg_begin)
{
    return k != kDeadKey && (j == seg_begin || keys[j - 1] != k);
}

// voxel runs that start in each sort tile
__global__ void __launch_bounds__(256)
    vb_runcount_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ tile_scan,
                       const uint32_t *__restrict__ meta, uint32_t *__restrict__ tile_runs,
                       uint32_t *__restrict__ tile_base)
{
    __shared__ uint32_t s_cnt[8];
    const uint32_t base = blockIdx.x * kSlotAlign;
    const uint32_t seg_begin = meta[5 * tile_scan[base / kTile] + 2];
    uint32_t c = 0;
    for (uint32_t l = threadIdx.x; l < (uint32_t)kSlotAlign; l += 256) {
        const uint32_t j = base + l;
        c += run_head(keys, j, keys[j], seg_begin) ? 1u : 0u;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        tile_runs[blockIdx.x] = t;
        tile_base[blockIdx.x] = t;  // scanned in place afterwards (no copy-engine work on the compute stream)
    }
}

// Centroids of the runs that start in one sort tile.  The tile's points are gathered into shared
// memory by all threads (independent loads), then each run head adds its run IN ORDER -- the same
// float chain as a sequential CPU loop -- continuing in global memory if the run crosses the tile end.
__global__ void __launch_bounds__(256)
    vb_centroid_kernel(const float4 *__restrict__ raw, const uint32_t *__restrict__ keys,
                       const uint32_t *__restrict__ vals, const uint32_t *__restrict__ tile_base,
                       const uint32_t *__restrict__ tile_scan, const uint32_t *__restrict__ meta,
                       float4 *__restrict__ src)
{
    constexpr int kPerThread = kSlotAlign / 256;
    // element l lives at l + (l >> 5): a thread owns kPerThread CONSECUTIVE elements, so without the
    // padding the lanes of a warp would sit kPerThread words apart -- an 8-way bank conflict on every
    // access of the head scan and the run sums
    constexpr int kPadded = kSlotAlign + kSlotAlign / 32;
    __shared__ uint32_t s_key[kPadded];
    __shared__ float s_x[kPadded], s_y[kPadded], s_z[kPadded];
    auto pad = [](uint32_t l) { return l + (l >> 5); };
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * kSlotAlign;
    const uint32_t s = tile_scan[base / kTile];
    const uint32_t seg_begin = meta[5 * s + 2], seg_end = seg_begin + meta[5 * s + 4] * kTile;
    {   // all loads of a thread issued back to back: keys and indices (coalesced), then the gathers
        uint32_t kk[kPerThread], vv[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            kk[i] = keys[base + i * 256 + threadIdx.x];
            vv[i] = vals[base + i * 256 + threadIdx.x];
        }
        float4 pp[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) pp[i] = kk[i] != kDeadKey ? raw[vv[i]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const uint32_t l = i * 256 + threadIdx.x;
            s_key[pad(l)] = kk[i];
            s_x[pad(l)] = pp[i].x; s_y[pad(l)] = pp[i].y; s_z[pad(l)] = pp[i].z;
        }
    }
    __syncthreads();
    // thread t owns the consecutive elements [t * kPerThread, (t + 1) * kPerThread): local run ids
    const uint32_t l0 = threadIdx.x * kPerThread;
    uint32_t heads = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const uint32_t l = l0 + i, k = s_key[pad(l)];
        const bool h = k != kDeadKey && (base + l == seg_begin || (l ? s_key[pad(l - 1)] : keys[base - 1]) != k);
        heads |= h ? (1u << i) : 0u;
        cnt += h ? 1u : 0u;
    }
    uint32_t incl = cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t rank = incl - cnt + tile_base[blockIdx.x];
    for (int w = 0; w < warp; ++w) rank += s_warp[w];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        if (!(heads & (1u << i))) continue;
        uint32_t l = l0 + i;
        const uint32_t k = s_key[pad(l)];
        float cx = 0.f, cy = 0.f, cz = 0.f;
        uint32_t n = 0;
        for (; l < (uint32_t)kSlotAlign && s_key[pad(l)] == k; ++l, ++n) {
            cx = __fadd_rn(cx, s_x[pad(l)]);
            cy = __fadd_rn(cy, s_y[pad(l)]);
            cz = __fadd_rn(cz, s_z[pad(l)]);
        }
        if (l == (uint32_t)kSlotAlign) {  // the run goes on in the next tile(s) of the same scan
            for (uint32_t e = base + kSlotAlign; e < seg_end && keys[e] == k; ++e, ++n) {
                const float4 p = raw[vals[e]];
                cx = __fadd_rn(cx, p.x);
                cy = __fadd_rn(cy, p.y);
                cz = __fadd_rn(cz, p.z);
            }
        }
        const float fn = (float)n;
        SSF_CHECK(seg_begin + rank < seg_end);
        src[seg_begin + rank] = make_float4(__fdiv_rn(cx, fn), __fdiv_rn(cy, fn), __fdiv_rn(cz, fn), 1.0f);
        ++rank;
    }
}

// centroids of scan s = runs that start in its tiles: last tile's base + count
__global__ void vb_counts_kernel(ScanState *st, const uint32_t *__restrict__ tile_base,
                                 const uint32_t *__restrict__ tile_runs, const uint4 *__restrict__ segs,
                                 uint32_t n_scans, int have_runs)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    uint32_t n = 0;
    if (have_runs && segs[s].y > 0) {
        const uint32_t last = segs[s].x + segs[s].y - 1;
        n = tile_base[last] + tile_runs[last];
    }
    st[s].n_pts = n;
}

int voxel_downsample_batch(BatchBuffers &b, const uint32_t *meta_dev, float leaf, Scratch &s, cudaStream_t st)
{
    const uint32_t n_scans = (uint32_t)b.n_scans, n_slots = (uint32_t)b.n_slots, tiles = (uint32_t)b.n_tiles;
    if (n_scans == 0) return SSF_OK;
    const uint32_t n_sort_tiles = n_slots / kSlotAlign;
    SSF_TRY(b.vbox.reserve((size_t)8 * n_scans));
    SSF_TRY(b.vgrid.reserve((size_t)8 * n_scans));
    SSF_TRY(b.vkeys.reserve((size_t)n_slots + 1));
    SSF_TRY(b.vvals.reserve((size_t)n_slots + 1));
    SSF_TRY(b.vseg.reserve((size_t)n_sort_tiles + 1));
    SSF_TRY(b.vsegs.reserve(n_scans));
    SSF_TRY(b.vruns.reserve((size_t)2 * n_sort_tiles + 2));
    int *vbox = reinterpret_cast<int *>(b.vbox.p);
    uint32_t *tile_runs = b.vruns.p, *tile_base = b.vruns.p + n_sort_tiles + 1;
    const float inv = 1.0f / leaf;
    const unsigned sb = (n_scans + 127) / 128;
    vb_init_kernel<<<sb, 128, 0, st>>>(vbox, n_scans);
    SSF_LAUNCHED();
    vb_segs_kernel<<<sb, 128, 0, st>>>(meta_dev, n_scans, b.vsegs.p);
    SSF_LAUNCHED();
    if (tiles > 0) {
        vb_bbox_kernel<<<n_sort_tiles, 256, 0, st>>>(b.raw.p, b.tile_scan.p, meta_dev, vbox);
        SSF_LAUNCHED();
        vb_grid_kernel<<<sb, 128, 0, st>>>(vbox, b.vgrid.p, n_scans, inv);
        SSF_LAUNCHED();
        vb_seg_kernel<<<(n_sort_tiles + 127) / 128, 128, 0, st>>>(b.tile_scan.p, meta_dev, n_sort_tiles, b.vseg.p);
        SSF_LAUNCHED();
        vb_keys_kernel<<<(n_slots + 255) / 256, 256, 0, st>>>(b.raw.p, b.tile_scan.p, meta_dev, b.vgrid.p, inv, b.vkeys.p,
                                                            b.vvals.p, n_slots);
        SSF_LAUNCHED();
        // the sort's own segment table carries each scan's first SLOT as base offset
        SSF_TRY(b.vsegs_sort.reserve(n_scans));
        vb_sortsegs_kernel<<<sb, 128, 0, st>>>(meta_dev, n_scans, b.vsegs_sort.p);
        SSF_LAUNCHED();
        SSF_TRY(seg_radix_sort_pairs_u32(b.vkeys.p, b.vvals.p, n_slots, b.vseg.p, b.vsegs_sort.p, n_scans, s, st));
        vb_runcount_kernel<<<n_sort_tiles, 256, 0, st>>>(b.vkeys.p, b.tile_scan.p, meta_dev, tile_runs, tile_base);
        SSF_LAUNCHED();
        SSF_TRY(seg_scan_u32(tile_base, b.vsegs.p, n_scans, 1, st));
        vb_centroid_kernel<<<n_sort_tiles, 256, 0, st>>>(b.raw.p, b.vkeys.p, b.vvals.p, tile_base, b.tile_scan.p, meta_dev,
                                                        b.src.p);
        SSF_LAUNCHED();
    }
    vb_counts_kernel<<<sb, 128, 0, st>>>(b.state.p, tile_base, tile_runs, b.vsegs.p, n_scans, tiles > 0 ? 1 : 0);
    SSF_LAUNCHED();
    return SSF_OK;
}

}  // namespace ssf
