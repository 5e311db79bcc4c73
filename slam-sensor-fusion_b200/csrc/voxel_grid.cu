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
