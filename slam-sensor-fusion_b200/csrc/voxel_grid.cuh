// voxel_grid.cuh -- K2: pcl::VoxelGrid semantics on the device.
#pragma once
#include "common.cuh"

namespace ssf {

struct VoxelWork {
    DevBuf<float4> in, out;
    DevBuf<unsigned long long> keys;
    DevBuf<uint32_t> vals, flags, scan;
    DevBuf<float> small;
};

// w.in holds n points; w.out receives *n_out centroids in ascending voxel index (w = 1).
// *refused = 1 when PCL's index-overflow guard fires: w.out = w.in, *n_out = n.
int voxel_downsample_device(VoxelWork &w, size_t n, float leaf, Scratch &s, cudaStream_t st, uint32_t *n_out,
                            int *refused);

}  // namespace ssf
