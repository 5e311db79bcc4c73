// voxel_grid.cuh -- K2: pcl::VoxelGrid semantics on the device.
#pragma once
#include "common.cuh"
#include "icp.cuh"

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

// Open3D semantics (double arithmetic, origin min_bound - voxel / 2): see voxel_grid.cu
int voxel_downsample_o3d_device(VoxelWork &w, size_t n, double voxel, Scratch &s, cudaStream_t st, uint32_t *n_out);

// Batched form: every scan of the batch (raw points in b.raw, tile-aligned slots, raw counts in
// meta[5*s + 1]) is downsampled with ONE segmented sort; centroids land in b.src at the scan's
// slots and ScanState::n_pts is set to the centroid count.  No host synchronisation.
int voxel_downsample_batch(BatchBuffers &b, const uint32_t *meta_dev, float leaf, Scratch &s, cudaStream_t st);

}  // namespace ssf
