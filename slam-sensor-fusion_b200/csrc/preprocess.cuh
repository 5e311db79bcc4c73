// preprocess.cuh -- device versions of point_cloud_processing.hpp (see preprocess.cu).
#pragma once
#include "common.cuh"

namespace ssf {

struct PreprocWork {
    DevBuf<float4> in, out;
    DevBuf<uint32_t> flags, scan, vals, small;
    DevBuf<float> d2;
    DevBuf<unsigned long long> keys;
    DevBuf<int32_t> idx;  // crop: original index of every output point
};

int subsample_device(PreprocWork &w, size_t n, size_t step, uint32_t *n_out, cudaStream_t st);
int remove_floor_device(PreprocWork &w, size_t n, Scratch &s, uint32_t *n_out, cudaStream_t st);
int crop_radius_device(PreprocWork &w, size_t n, const float center[3], double radius, Scratch &s, uint32_t *n_out,
                       cudaStream_t st);

}  // namespace ssf
