// ingest.cuh -- what ingest.cu needs from the handles defined in abi.cu.
#pragma once
#include "common.cuh"

struct ssf_ctx;
struct ssf_icp;

namespace ssf {

// a context's stream / scratch / upload path, for translation units that do not see struct ssf_ctx
struct ssf_ctx_ref {
    ssf_ctx *c;
    explicit ssf_ctx_ref(ssf_ctx *c_) : c(c_) {}
    int use() const;                 // cudaSetDevice
    cudaStream_t stream() const;
    Scratch &scratch() const;
    // host cloud (floats at a byte stride) -> float4 on the device (stream-ordered)
    int upload_cloud(const float *xyz, size_t n, size_t stride_bytes, float4 *dst) const;
};

// setTargetPointCloud from points that are already in HBM (same device as the handle): copy + index build
int icp_set_target_device(ssf_icp *icp, const float4 *pts_dev, size_t n, const ssf_ctx_ref &from);

// records with float32 / float64 x, y, z at byte offsets -> float4 (w = 1)
int extract_xyz_device(const unsigned char *raw_dev, size_t n, size_t point_bytes, size_t ox, size_t oy, size_t oz,
                       bool f64, bool swap, float4 *out, cudaStream_t st);

}  // namespace ssf
