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

__global__ void __launch_bounds__(kTile)
    vb_bbox_kernel(const float4 *__restrict__ raw, const uint32_t *__restrict__ tile_scan,
                   const uint32_t *__restrict__ meta, int *vbox)
{
    const uint32_t s = tile_scan[blockIdx.x];
    const uint32_t n = meta[5 * s + 1], pt_begin = meta[5 * s + 2], tile_begin = meta[5 * s + 3];
    const uint32_t row = (blockIdx.x - tile_begin) * kTile + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t cnt = 0;
    if (row < n) {
        const float4 p = raw[(size_t)pt_begin + row];
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            mn[0] = mx[0] = p.x; mn[1] = mx[1] = p.y; mn[2] = mx[2] = p.z;
            cnt = 1;
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
    __shared__ float smn[kTile / 32][3], smx[kTile / 32][3];
    __shared__ uint32_t scnt[kTile / 32];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { smn[warp][k] = mn[k]; smx[warp][k] = mx[k]; }
        scnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (int w = 0; w < kTile / 32; ++w) {
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

// flags[j] = 1 on the first element of every voxel run; flags[n_slots] = 0 (so the exclusive
// scan also yields the total)
__global__ void __launch_bounds__(256)
    vb_flags_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ tile_scan,
                    const uint32_t *__restrict__ meta, uint32_t n_slots, uint32_t *__restrict__ flags)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_slots) return;
    uint32_t f = 0;
    if (j < n_slots) {
        const uint32_t k = keys[j];
        if (k != kDeadKey) {
            const uint32_t s = tile_scan[j / kTile];
            f = (j == meta[5 * s + 2] || keys[j - 1] != k) ? 1u : 0u;
        }
    }
    flags[j] = f;
}

__global__ void __launch_bounds__(128)
    vb_centroid_kernel(const float4 *__restrict__ raw, const uint32_t *__restrict__ keys,
                       const uint32_t *__restrict__ vals, const uint32_t *__restrict__ flags,
                       const uint32_t *__restrict__ scan, uint32_t n_slots, const uint32_t *__restrict__ tile_scan,
                       const uint32_t *__restrict__ meta, float4 *__restrict__ src)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_slots || !flags[j]) return;
    const uint32_t k = keys[j];
    const uint32_t s = tile_scan[j / kTile];
    const uint32_t pt_begin = meta[5 * s + 2], seg_end = pt_begin + meta[5 * s + 4] * kTile;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    uint32_t e = j;
    while (e < seg_end && keys[e] == k) {
        const float4 p = raw[vals[e]];
        cx = __fadd_rn(cx, p.x);
        cy = __fadd_rn(cy, p.y);
        cz = __fadd_rn(cz, p.z);
        ++e;
    }
    const float cnt = (float)(e - j);
    const uint32_t out = pt_begin + (scan[j] - scan[pt_begin]);
    src[out] = make_float4(__fdiv_rn(cx, cnt), __fdiv_rn(cy, cnt), __fdiv_rn(cz, cnt), 1.0f);
}

// centroids of scan s = voxel runs that start inside its slot range
__global__ void vb_counts_kernel(ScanState *st, const uint32_t *__restrict__ scan, const uint32_t *__restrict__ meta,
                                 uint32_t n_scans, int have_runs)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    const uint32_t pt_begin = meta[5 * s + 2], seg_end = pt_begin + meta[5 * s + 4] * kTile;
    st[s].n_pts = have_runs ? scan[seg_end] - scan[pt_begin] : 0u;
}

int voxel_downsample_batch(BatchBuffers &b, const uint32_t *meta_dev, float leaf, Scratch &s, cudaStream_t st)
{
    const uint32_t n_scans = (uint32_t)b.n_scans, n_slots = (uint32_t)b.n_slots, tiles = (uint32_t)b.n_tiles;
    if (n_scans == 0) return SSF_OK;
    SSF_TRY(b.vbox.reserve((size_t)8 * n_scans));
    SSF_TRY(b.vgrid.reserve((size_t)8 * n_scans));
    SSF_TRY(b.vflags.reserve((size_t)n_slots + 1));
    SSF_TRY(b.vscan.reserve((size_t)n_slots + 1));
    SSF_TRY(b.vkeys.reserve((size_t)n_slots + 1));
    SSF_TRY(b.vvals.reserve((size_t)n_slots + 1));
    int *vbox = reinterpret_cast<int *>(b.vbox.p);
    const float inv = 1.0f / leaf;
    const unsigned sb = (n_scans + 127) / 128;
    vb_init_kernel<<<sb, 128, 0, st>>>(vbox, n_scans);
    SSF_LAUNCHED();
    if (tiles > 0) {
        const uint32_t n_sort_tiles = n_slots / kSlotAlign;
        SSF_TRY(b.vseg.reserve(n_sort_tiles));
        vb_bbox_kernel<<<tiles, kTile, 0, st>>>(b.raw.p, b.tile_scan.p, meta_dev, vbox);
        SSF_LAUNCHED();
        vb_grid_kernel<<<sb, 128, 0, st>>>(vbox, b.vgrid.p, n_scans, inv);
        SSF_LAUNCHED();
        vb_seg_kernel<<<(n_sort_tiles + 127) / 128, 128, 0, st>>>(b.tile_scan.p, meta_dev, n_sort_tiles, b.vseg.p);
        SSF_LAUNCHED();
        const unsigned fb = (n_slots + 256) / 256;  // n_slots + 1 elements
        vb_keys_kernel<<<fb, 256, 0, st>>>(b.raw.p, b.tile_scan.p, meta_dev, b.vgrid.p, inv, b.vkeys.p, b.vvals.p, n_slots);
        SSF_LAUNCHED();
        SSF_TRY(seg_radix_sort_pairs_u32(b.vkeys.p, b.vvals.p, n_slots, b.vseg.p, s, st));
        vb_flags_kernel<<<fb, 256, 0, st>>>(b.vkeys.p, b.tile_scan.p, meta_dev, n_slots, b.vflags.p);
        SSF_LAUNCHED();
        SSF_TRY(exclusive_scan_u32(b.vflags.p, b.vscan.p, (size_t)n_slots + 1, nullptr, s, st));
        vb_centroid_kernel<<<(n_slots + 127) / 128, 128, 0, st>>>(b.raw.p, b.vkeys.p, b.vvals.p, b.vflags.p, b.vscan.p,
                                                                 n_slots, b.tile_scan.p, meta_dev, b.src.p);
        SSF_LAUNCHED();
    }
    vb_counts_kernel<<<sb, 128, 0, st>>>(b.state.p, b.vscan.p, meta_dev, n_scans, tiles > 0 ? 1 : 0);
    SSF_LAUNCHED();
    return SSF_OK;
}

}  // namespace ssf
