// map_index.cuh -- owner of the HBM-resident voxel-hash map (see MapView in common.cuh).
#pragma once
#include <climits>

#include "common.cuh"

namespace ssf {

struct MapIndex {
    DevBuf<float4> raw;      // target cloud in ORIGINAL order (w unused); kept for re-indexing
    DevBuf<float4> raw_nrm;  // normals in original order (optional)
    DevBuf<int32_t> global_index;  // map sharding: global index of every raw point (optional)
    bool has_global_index = false;
    bool sharded = false;          // origin / ownership given by the caller (ssf_shard_info)
    float shard_origin[3] = {0, 0, 0};
    int own_lo = INT32_MIN, own_hi = INT32_MAX;
    DevBuf<float4> pts;      // sorted by cell key, w = original index
    DevBuf<float4> nrm;      // sorted normals
    DevBuf<unsigned long long> keys;
    DevBuf<uint32_t> vals;
    DevBuf<uint32_t> flags;
    DevBuf<uint32_t> cell_id;
    DevBuf<unsigned long long> cell_keys;
    DevBuf<uint32_t> cell_start;
    DevBuf<unsigned long long> hkeys;
    DevBuf<uint4> hvals;
    DevBuf<float> small;  // bbox (6 floats) + counters
    MapView view{};
    size_t n_raw = 0;       // points given to set_target
    bool has_normals = false;
    float cell_size = 0.f;
    uint32_t n_cells = 0, n_entries = 0, table_size = 0;
    float bbox[6] = {0, 0, 0, 0, 0, 0};
};

// raw (and raw_nrm when has_normals) must already hold n_raw points on the device.
int build_map_index(MapIndex &m, float cell_size, Scratch &s, cudaStream_t st);

}  // namespace ssf
