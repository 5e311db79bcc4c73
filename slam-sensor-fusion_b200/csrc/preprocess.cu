// preprocess.cu -- the cloud pre-processing that runs right before the registration
// (reference localization/include/localization/point_cloud_processing.hpp):
//   applyUniformSubsample (:55-74)        every point_step-th point, unchanged if size < step
//   removeFloor (:76-92)                   keep z > 0, order preserved
//   cropPointCloudThroughRadius (:31-53)   pcl::search::KdTree::radiusSearch around the pose's
//                                          translation: points with d2 < r^2 (FLANN: strict),
//                                          SORTED by ascending distance, ties by index
// Callers: localization_node.cpp:20, 211-213, 292, 296, 302.  Device versions: flag -> exclusive
// scan -> scatter (stable compaction), plus a stable radix sort on the d2 bit pattern for the crop.
#include <cfloat>
#include <cmath>

#include "common.cuh"
#include "preprocess.cuh"

namespace ssf {

__global__ void __launch_bounds__(256) subsample_kernel(const float4 *__restrict__ in, uint32_t n_out, uint32_t step,
                                                        float4 *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out[i] = in[(size_t)i * step];
}

__global__ void __launch_bounds__(256) floor_flags_kernel(const float4 *__restrict__ in, uint32_t n,
                                                          uint32_t *__restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = in[i].z > 0.f ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
    crop_flags_kernel(const float4 *__restrict__ in, uint32_t n, float cx, float cy, float cz, float r2,
                      uint32_t *__restrict__ flags, float *__restrict__ d2_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    // flann::L2_Simple between the query (centre) and the point, float, no FMA
    const float dx = __fsub_rn(cx, p.x), dy = __fsub_rn(cy, p.y), dz = __fsub_rn(cz, p.z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    d2_out[i] = d2;
    flags[i] = (isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && d2 < r2) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
    compact_kernel(const float4 *__restrict__ in, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ scan,
                   uint32_t n, float4 *__restrict__ out, int32_t *__restrict__ idx_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    out[scan[i]] = in[i];
    if (idx_out) idx_out[scan[i]] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
    crop_keys_kernel(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ scan, const float *__restrict__ d2,
                     uint32_t n, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    keys[scan[i]] = (unsigned long long)__float_as_uint(d2[i]);  // d2 >= 0: bit pattern orders like the value
    vals[scan[i]] = i;
}

__global__ void __launch_bounds__(256) gather_kernel(const float4 *__restrict__ in, const uint32_t *__restrict__ vals,
                                                     uint32_t n_out, float4 *__restrict__ out, int32_t *__restrict__ idx_out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    out[j] = in[vals[j]];
    if (idx_out) idx_out[j] = (int32_t)vals[j];
}

static inline unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

int subsample_device(PreprocWork &w, size_t n, size_t step, uint32_t *n_out, cudaStream_t st)
{
    if (step == 0) {
        set_error("applyUniformSubsample: point_step must be >= 1");
        return SSF_ERR_INVALID;
    }
    if (n < step) {  // hpp:58-61: cloud left unchanged
        if (n) SSF_CUDA(cudaMemcpyAsync(w.out.p, w.in.p, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        *n_out = (uint32_t)n;
        return SSF_OK;
    }
    const uint32_t m = (uint32_t)((n + step - 1) / step);
    if (m) {
        subsample_kernel<<<blocks_for(m), 256, 0, st>>>(w.in.p, m, (uint32_t)step, w.out.p);
        SSF_LAUNCHED();
    }
    *n_out = m;
    return SSF_OK;
}

static int compact_common(PreprocWork &w, size_t n, Scratch &s, cudaStream_t st, uint32_t *n_out)
{
    SSF_TRY(w.small.reserve(4));
    SSF_TRY(exclusive_scan_u32(w.flags.p, w.scan.p, n, w.small.p, s, st));
    SSF_CUDA(cudaMemcpyAsync(n_out, w.small.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    return SSF_OK;
}

int remove_floor_device(PreprocWork &w, size_t n, Scratch &s, uint32_t *n_out, cudaStream_t st)
{
    *n_out = 0;
    if (n == 0) return SSF_OK;
    SSF_TRY(w.flags.reserve(n));
    SSF_TRY(w.scan.reserve(n));
    floor_flags_kernel<<<blocks_for(n), 256, 0, st>>>(w.in.p, (uint32_t)n, w.flags.p);
    SSF_LAUNCHED();
    SSF_TRY(compact_common(w, n, s, st, n_out));
    compact_kernel<<<blocks_for(n), 256, 0, st>>>(w.in.p, w.flags.p, w.scan.p, (uint32_t)n, w.out.p, nullptr);
    SSF_LAUNCHED();
    return SSF_OK;
}

int crop_radius_device(PreprocWork &w, size_t n, const float center[3], double radius, Scratch &s, uint32_t *n_out,
                       cudaStream_t st)
{
    *n_out = 0;
    if (n == 0) return SSF_OK;
    SSF_TRY(w.flags.reserve(n));
    SSF_TRY(w.scan.reserve(n));
    SSF_TRY(w.d2.reserve(n));
    SSF_TRY(w.keys.reserve(n));
    SSF_TRY(w.vals.reserve(n));
    SSF_TRY(w.idx.reserve(n));
    const float r2 = (float)(radius * radius);  // pcl::KdTreeFLANN::radiusSearch passes float(radius * radius)
    crop_flags_kernel<<<blocks_for(n), 256, 0, st>>>(w.in.p, (uint32_t)n, center[0], center[1], center[2], r2, w.flags.p,
                                                    w.d2.p);
    SSF_LAUNCHED();
    SSF_TRY(compact_common(w, n, s, st, n_out));
    if (*n_out == 0) return SSF_OK;
    crop_keys_kernel<<<blocks_for(n), 256, 0, st>>>(w.flags.p, w.scan.p, w.d2.p, (uint32_t)n, w.keys.p, w.vals.p);
    SSF_LAUNCHED();
    SSF_TRY(radix_sort_pairs_u64(w.keys.p, w.vals.p, *n_out, 32, s, st));  // stable: equal d2 stay in index order
    gather_kernel<<<blocks_for(*n_out), 256, 0, st>>>(w.in.p, w.vals.p, *n_out, w.out.p, w.idx.p);
    SSF_LAUNCHED();
    return SSF_OK;
}

}  // namespace ssf
