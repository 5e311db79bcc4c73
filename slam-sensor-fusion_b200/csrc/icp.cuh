// icp.cuh -- device-side state of a batch of scans and the host entry points of the solver.
#pragma once
#include <vector>

#include "common.cuh"
#include "map_index.cuh"

namespace ssf {

constexpr int kTile = 512;     // queries per block of the search kernels; scans are tile-aligned
constexpr int kThreads = 128;  // threads of a fused search block
constexpr int kWideThreads = 512;  // ... for small batches (latency): more threads per tile
constexpr int kAccum = 32;  // doubles per partial row
constexpr int kSlotAlign = kSortTileSize;  // a scan's slot range is a whole number of sort tiles
static_assert(kSlotAlign % kTile == 0, "slot alignment must be a multiple of the search tile");

// One scan of a batch (device memory).
struct ScanState {
    float T[16];       // composed transform, column-major
    float T_step[16];  // REFERENCE: last step, applied to P by ref_step_kernel
    float T_init[16];
    float last_error;  // ICPPointToPoint::last_error_
    float error;       // error reported in the result
    float pend_error;  // REFERENCE: error of the pass that decided to re-search
    int iterations, done, converged, aborted, n_searches, k_last, need_search, have_step;
    double fitness, rmse;  // O3D
    float last_step;       // size of the last pose update (metres; rotations weighted by a 30 m lever arm)
    int comm_error;        // map sharding: a peer's rows did not arrive in time (solve_xchg_kernel)
    uint32_t pt_begin;     // first slot in the packed arrays (multiple of kTile)
    uint32_t n_pts;        // source points (after optional voxel downsample)
    uint32_t tile_begin, tile_cap;
};

struct BatchBuffers {
    // packed, tile-aligned per scan
    DevBuf<float4> src;        // source points (sensor frame)
    DevBuf<float4> P;          // REFERENCE: transformed source, advanced in place
    DevBuf<float4> Q;          // REFERENCE: matched target point per row
    DevBuf<int32_t> corr;      // correspondence (original target index) or -1
    DevBuf<uint2> cert;        // search certificates: (radius | issuing iteration, neighbour position), nn_device.cuh
    DevBuf<float> pose_hist;   // [scan][kCertHist][16]: pose used by search launch i (certificates refer to it)
    DevBuf<uint32_t> tile_scan;
    DevBuf<uint4> active;        // tiles that hold points (search kernel work list): (tile, scan, first slot, points)
    DevBuf<uint32_t> counters;   // [0 .. P]: tiles listed for search launch i (unsharded: [0] serves every launch),
                                 // [P + 1 .. 2P + 1]: tile fetch counter of launch i; P = num_iterations + 1
    DevBuf<float4> tile_box;     // map sharding: bounding box of every tile's source points (ShardView)
    DevBuf<unsigned long long> search_stats;  // per search launch: (queries answered, queries walked)
    int search_stats_len = 0;
    DevBuf<double> partials;   // [tile][kAccum]
    DevBuf<double> sums;       // [scan][kAccum]: per-scan totals (all-reduced across ranks when sharded)
    DevBuf<ScanState> state;
    DevBuf<ssf_icp_result> results;
    DevBuf<float> trace_err;       // [scan][num_iterations]
    DevBuf<int32_t> trace_search;  // [scan][num_iterations]
    // voxel-downsample stage
    DevBuf<float4> raw;            // raw scans when source_voxel_leaf > 0
    DevBuf<uint32_t> vkeys;
    DevBuf<uint32_t> vvals;
    DevBuf<uint2> vseg;            // per sort tile: (first tile of its scan, tiles in it)
    DevBuf<uint4> vsegs;           // per scan: (first sort tile, sort tiles, 0, 0): run ids restart per scan
    DevBuf<uint4> vsegs_sort;      // per scan: (first sort tile, sort tiles, first slot, 0)
    DevBuf<uint32_t> vruns;        // per sort tile: voxel runs starting in it, then their exclusive scan
    DevBuf<float> vbox;            // per scan: min xyz, max xyz (ordered ints during reduce)
    DevBuf<int32_t> vgrid;         // per scan: minb[3], divb[3], refused, pad
    size_t n_scans = 0, n_tiles = 0, n_slots = 0;
    size_t max_tiles = 0;          // capacity in search tiles (fixed at creation)
    // CUDA graph of the alignment loop, replayed while the launch shape stays the same
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned long long> graph_key;
    uint64_t graph_kernels = 0;
    ~BatchBuffers()
    {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
    }
    int trace_len = 0;
};

// optional event bracket around the search kernels (roofline measurement, see ssf.h)
struct SearchTimer {
    bool enabled = false;
    std::vector<cudaEvent_t> pool;   // created on demand, reused
    size_t used = 0;                 // events handed out since the last collect (pairs)
    int begin(cudaStream_t st);
    int end(cudaStream_t st);
};

// all-reduce hook (map sharding): sum `count` doubles at `buf` across ranks, enqueued on `stream`
typedef int (*AllreduceFn)(void *user, double *buf, size_t count, void *stream);

// Map sharding, in-kernel exchange: every rank owns one buffer (rows[2][world][max_scans][32] doubles
// followed by flags[2][world] u64) and holds peer pointers to all of them (CUDA IPC over NVLink).
// The row-sum kernel STORES this rank's per-scan rows into every rank's buffer and then publishes an
// epoch in each rank's flag; the solve kernel waits for all ranks' flags and adds the rows in rank
// order -- an all-gather + ordered reduction inside the two kernels, no host hook, no NCCL call.
struct XchView {
    void *const *peers = nullptr;  // device array [world] of buffer base pointers (own entry included)
    int rank = 0, world = 0;
    uint32_t max_scans = 0;
    unsigned long long *epoch = nullptr;  // device: epoch of the current run's first pass (advanced by the run itself)
    unsigned long long timeout_ns = 20000000000ull;  // give up waiting for a peer's rows after this long
};

// Map sharding: ownership test of whole tiles (icp_kernels.cu, tile_may_own).  enabled == 0: every tile is listed.
struct ShardView {
    int enabled = 0;
    float ox = 0.f, inv_h = 1.f;
    int own_lo = 0, own_hi = 0;
    const float4 *tile_box = nullptr;  // [2 * tile]: min / max corner of the tile's source points, sensor frame
};

// all-reduce behind the C ABI (NCCL, loaded at run time: nccl_link.cu): sum `count` doubles in place on `stream`
typedef int (*NcclAllreduceFn)(void *user, double *buf, size_t count, cudaStream_t stream);

struct IcpConfig {
    float max_corr;
    int num_iterations;
    float acc_err;
    float eps;
    int mode;
    int reduce;
    AllreduceFn allreduce = nullptr;
    void *allreduce_user = nullptr;
    XchView xch;                      // world > 0: exchange in-kernel instead of through the hook
    NcclAllreduceFn nccl_allreduce = nullptr;  // ncclAllReduce on the library's stream (graph-capturable)
    void *nccl_user = nullptr;
};

// Enqueue the whole alignment of every scan in the batch on `st` (no host sync inside).
// T_init: n_scans x 16 floats (column-major), device-readable (pinned host memory is fine).
int run_batch(const MapView &map, const IcpConfig &cfg, BatchBuffers &b, const float *T_init, cudaStream_t st,
              SearchTimer *timer);
// standalone search over n already-transformed queries (device pointers)
int nn_search_device(const MapView &map, const float4 *queries, size_t n, float limit, int32_t *idx, float *d2,
                     cudaStream_t st);
// initialise ScanState from T_init (device array n_scans x 16) -- also resets counters
int init_states(BatchBuffers &b, const float *T_init_dev, cudaStream_t st);

}  // namespace ssf
