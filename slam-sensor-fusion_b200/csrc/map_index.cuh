// map_index.cuh -- owner of the HBM-resident voxel-grid map (see MapView in common.cuh).
#pragma once
#include <climits>

#include "common.cuh"

namespace ssf {

struct MapIndex {
    DevBuf<float4> raw;      // target cloud in ORIGINAL order (w unused)
    DevBuf<float4> raw_nrm;  // normals in original order (optional)
    DevBuf<int32_t> global_index;  // map sharding: global index of every raw point (optional)
    bool has_global_index = false;
    bool sharded = false;          // ownership columns given by the caller (ssf_shard_info)
    float shard_origin[3] = {0, 0, 0};
    float shard_cell = 0.f;
    int own_lo = INT32_MIN, own_hi = INT32_MAX;
    DevBuf<float4> pts;      // sorted by directory key, w = original index
    DevBuf<float4> nrm;      // sorted (point, normal) records, 32 bytes each (MapView::pn)
    DevBuf<unsigned long long> keys;
    DevBuf<uint32_t> vals;
    DevBuf<uint32_t> flags;
    DevBuf<uint32_t> cell_id;
    DevBuf<uint32_t> cell_start;
    DevBuf<uint2> dir;
    DevBuf<float> small;  // bbox (6 floats) + counters
    DevBuf<uint32_t> reach;     // reach mask of view.reach, built for reach_limit
    float reach_limit = -1.f;   // the rejection threshold the mask was built for (< 0: none)
    MapView view{};
    size_t n_raw = 0;       // points given to set_target
    bool has_normals = false;
    float cell_size = 0.f;
    uint32_t n_cells = 0, n_dir = 0;
    int build_passes = 0;   // sort passes the cell-size search needed
    float bbox[6] = {0, 0, 0, 0, 0, 0};
};

// raw (and raw_nrm when has_normals) must already hold n_raw points on the device.
// cell_size > 0 fixes the cell edge; cell_size <= 0 picks it from the measured occupancy
// (target: a few points per occupied cell).
int build_map_index(MapIndex &m, float cell_size, Scratch &s, cudaStream_t st);

// Make m.view.reach valid for searches with the rejection threshold `limit` (squared distance) and the
// map's certificate margin: (re)built when the threshold changes, dropped (nullptr) when the reach
// spans too many cells to pay off.  Stream-ordered; never call inside a graph capture.
int ensure_reach_mask(MapIndex &m, float limit, cudaStream_t st);

}  // namespace ssf
