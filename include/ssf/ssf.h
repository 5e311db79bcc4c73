/*
 * ssf.h -- C ABI of libssf_gpu.so: scan-to-map registration on one B200 (sm_100a).
 *
 * This is the drop-in boundary for the registration hot path of
 * viniciusvidal2/slam-sensor-fusion.  Every entry point names the reference interface it
 * replaces (paths relative to the reference repository root).  Plain C: opaque handles,
 * plain pointers and sizes, int status codes; no C++ / torch / PCL / Eigen types.
 *
 * Conventions
 *   - Points are float triples at a caller-given byte stride (16 for pcl::PointXYZ /
 *     float4, 12 for packed xyz).  All pointers are HOST pointers unless a function says
 *     "device".  The library copies at set time: the caller may free or overwrite its
 *     buffer as soon as the call returns, exactly like the reference's deep copies
 *     (localization/src/icp_point_to_point.cpp:44-55).
 *   - 4x4 transforms are 16 floats, COLUMN-major, like Eigen::Matrix4f.
 *   - Every function returns SSF_OK (0) or a negative ssf_status; ssf_last_error() gives
 *     the message of the calling thread's last failure.  There is no CPU fallback: with no
 *     usable CUDA device ssf_ctx_create fails with SSF_ERR_CUDA.
 *   - A handle is not thread-safe (the reference object is used from one single-threaded
 *     executor, localization/src/main.cpp:18); distinct handles are independent.
 */
#ifndef SSF_SSF_H
#define SSF_SSF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssf_ctx ssf_ctx;     /* one CUDA device: stream, scratch memory            */
typedef struct ssf_icp ssf_icp;     /* one ICPPointToPoint object: map in HBM + parameters */
typedef struct ssf_batch ssf_batch; /* a set of scans resident in HBM, aligned together    */
typedef struct ssf_map ssf_map;     /* the whole map cloud resident in HBM (source of target crops) */

typedef enum {
    SSF_OK = 0,
    SSF_ERR_INVALID = -1, /* bad argument                                   */
    SSF_ERR_CUDA = -2,    /* CUDA runtime failure (message has the detail)  */
    SSF_ERR_NOMEM = -3,   /* host or device allocation failed               */
    SSF_ERR_STATE = -4,   /* call order: no target / no source set          */
    SSF_ERR_COMM = -5     /* multi-GPU exchange failure                     */
} ssf_status;

/* Solver run by ssf_icp_align. */
typedef enum {
    /* ICPPointToPoint::calculateAlignment as written (icp_point_to_point.cpp:185-254):
     * one search before the loop, lazy re-search, shrinking source, Kabsch/SVD step.      */
    SSF_MODE_REFERENCE = 0,
    /* Gauss-Newton, re-search every iteration, 6x6 normal equations + Cholesky on device. */
    SSF_MODE_GN_P2P = 1,
    SSF_MODE_GN_P2PLANE = 2, /* needs target normals */
    /* Control flow of open3d registration_icp(PointToPoint) used by the Python node
     * (localization_python/localization_python/localization_node.py:233-237).             */
    SSF_MODE_O3D_P2P = 3
} ssf_mode;

/* How the per-iteration sums are formed in SSF_MODE_REFERENCE. */
typedef enum {
    /* float, sequential, in source-row order -- the reference's own loops
     * (icp_point_to_point.cpp:117-121, 126-134, 164-167); results are bit-identical to the
     * CPU restatement, including which iterations re-search.                              */
    SSF_REDUCE_STRICT = 0,
    /* double, parallel tree; fastest, differs from STRICT by float summation noise.       */
    SSF_REDUCE_FAST = 1
} ssf_reduce;

/* Replaces the constructor arguments and setters of ICPPointToPoint
 * (icp_point_to_point.h:49-80, icp_point_to_point.cpp:3-42). */
typedef struct {
    float max_correspondence_dist; /* compared with the SQUARED distance (cpp:70) */
    int32_t num_iterations;
    float acceptable_mean_error;
    float transformation_epsilon;
    int32_t mode;   /* ssf_mode   */
    int32_t reduce; /* ssf_reduce */
    int32_t debug;  /* setDebugMode: print the per-iteration trace of cpp:172-183,237-246 */
    float source_voxel_leaf; /* > 0: voxel-grid downsample each source scan first (north-star step 1) */
} ssf_icp_params;

/* Replaces struct ICPResult (icp_point_to_point.h:28-39) plus diagnostics. */
typedef struct {
    float transformation[16]; /* column-major; = initial transform when aborted           */
    float error;              /* 1e6 when aborted (ICPResult default)                      */
    int32_t iterations;
    int32_t has_converged;
    int32_t n_searches;       /* correspondence searches run                               */
    int32_t k_final;          /* correspondences in the last search                        */
    int32_t aborted;          /* 1: < 10 correspondences on the first search (cpp:196-200); 2: multi-GPU exchange failed */
    float fitness;            /* k_final / n_source                                        */
    int32_t n_source;         /* source points after optional voxel downsample             */
    float device_ms;          /* device time of the call (CUDA events)                     */
} ssf_icp_result;

const char *ssf_last_error(void);
const char *ssf_version(void);

/* ---- context ------------------------------------------------------------------------- */
int ssf_ctx_create(int device_ordinal, ssf_ctx **out);
void ssf_ctx_destroy(ssf_ctx *ctx);
int ssf_ctx_synchronize(ssf_ctx *ctx);
/* cudaStream_t of the context (as void*), for callers that time with their own events. */
void *ssf_ctx_stream(ssf_ctx *ctx);

/* ---- registration object: ICPPointToPoint -------------------------------------------- */
/* ICPPointToPoint::ICPPointToPoint (icp_point_to_point.cpp:3-12). */
int ssf_icp_create(ssf_ctx *ctx, const ssf_icp_params *params, ssf_icp **out);
void ssf_icp_destroy(ssf_icp *icp);
/* setMaxCorrespondenceDist / setNumIterations / setTransformationEpsilon /
 * setAcceptableMeanError / setDebugMode (icp_point_to_point.cpp:14-42). */
int ssf_icp_set_params(ssf_icp *icp, const ssf_icp_params *params);
int ssf_icp_get_params(const ssf_icp *icp, ssf_icp_params *out);
/* setTargetPointCloud (icp_point_to_point.cpp:49-55): copy the map to HBM and build the
 * voxel-grid index that replaces kdtree_.setInputCloud.  normals (optional, same count)
 * are required by SSF_MODE_GN_P2PLANE.  The map stays resident until the next call. */
int ssf_icp_set_target(ssf_icp *icp, const float *xyz, size_t n, size_t stride_bytes, const float *normals,
                       size_t normals_stride_bytes);
/* setSourcePointCloud (icp_point_to_point.cpp:44-47). */
int ssf_icp_set_source(ssf_icp *icp, const float *xyz, size_t n, size_t stride_bytes);
/* setInitialTransformation (icp_point_to_point.cpp:34-37). */
int ssf_icp_set_initial(ssf_icp *icp, const float T_colmajor[16]);
/* calculateAlignment (icp_point_to_point.cpp:185-254). */
int ssf_icp_align(ssf_icp *icp, ssf_icp_result *out);
/* Correspondence of every ORIGINAL source row after the last align: target index or -1
 * (the std::vector filled at icp_point_to_point.cpp:60-75, kept per row).  n = n_source. */
int ssf_icp_get_correspondences(ssf_icp *icp, int32_t *idx_out, size_t n);
/* Per-pass trace of the last align (REFERENCE mode): error measured at the top of pass i
 * (printStepDebug, cpp:172-183) and whether pass i re-searched.  NaN / 0 for passes not run. */
int ssf_icp_get_trace(ssf_icp *icp, float *iter_err, int32_t *iter_searched, size_t n);
size_t ssf_icp_target_size(const ssf_icp *icp);

/* kdtree_.nearestKSearch(p, 1, ..) + the d2 < max_sqdist test for n query points
 * (icp_point_to_point.cpp:64-70) against the handle's target.  idx[i] = index of the
 * nearest target point (lowest index on equal distance), d2[i] its squared distance; -1 /
 * FLT_MAX when no target point has d2 < max_sqdist.  Parity and benchmark entry. */
int ssf_nn_search(ssf_icp *icp, const float *queries, size_t n, size_t stride_bytes, float max_sqdist, int32_t *idx,
                  float *d2);

/* Throughput of the search alone ("NN queries/sec" of BASELINE.json): uploads the n queries once,
 * runs the search kernel `reps` times back to back and returns the average device time of one
 * pass in *ms_per_pass (CUDA events).  Results of the last pass go to idx / d2 when non-NULL. */
int ssf_nn_search_bench(ssf_icp *icp, const float *queries, size_t n, size_t stride_bytes, float max_sqdist, int reps,
                        float *ms_per_pass, int32_t *idx, float *d2);

/* pcl::VoxelGrid<PointXYZ> with setLeafSize(leaf, leaf, leaf)
 * (localization/src/global_map_frames_manager.cpp:143-146).  out must hold n float4
 * (16-byte stride, w = 1).  *refused = 1 when PCL's index-overflow guard fires, in which
 * case the input is returned unchanged like PCL does. */
int ssf_voxel_downsample(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, float leaf, float *out,
                         size_t *n_out, int *refused);

/* open3d PointCloud.voxel_down_sample(voxel_size) of the Python node (localization_python/localization_python/
 * localization_node.py:47): voxel origin = min_bound - voxel_size / 2, index = floor((p - origin) / voxel_size),
 * centroid = sum / count, all in double on the float32 coordinates; output in ascending (z, y, x) voxel index
 * (Open3D's own order is that of an unordered_map), rounded to float32.  out holds n float4. */
int ssf_voxel_downsample_o3d(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, double voxel_size, float *out,
                             size_t *n_out);

/* ---- cloud pre-processing (localization/include/localization/point_cloud_processing.hpp) ---- */
/* out buffers hold n float4 (16-byte stride, w = 1); *n_out receives the number written. */
/* applyUniformSubsample (hpp:55-74): points 0, step, 2*step, ...; unchanged when n < step. */
int ssf_cloud_subsample(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, size_t point_step, float *out,
                        size_t *n_out);
/* removeFloor (hpp:76-92): keep points with z > 0, order preserved. */
int ssf_cloud_remove_floor(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, float *out, size_t *n_out);
/* cropPointCloudThroughRadius (hpp:31-53): points with squared distance to `center` (the pose's
 * translation T(0..2,3)) < float(radius*radius), ordered by ascending distance, ties by index --
 * the order pcl::search::KdTree::radiusSearch returns.  indices_out (optional, n ints) receives
 * the original index of every output point. */
int ssf_cloud_crop_radius(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, const float center[3],
                          double radius, float *out, size_t *n_out, int32_t *indices_out);

/* ---- map ingestion and the resident map (rows N3 / N2 of the survey) ------------------------------ */
/* pcl::io::loadPCDFile<PointXYZ> for the recorder's tiles (mapping/src/map_data_save_node.cpp:71-80; read at
 * localization/src/global_map_frames_manager.cpp:101,129): DATA binary or ascii, x / y / z float32 or float64
 * among any other fields.  xyz_out == NULL: only *n_points is set (size query); else n x 3 packed floats. */
int ssf_pcd_read(const char *path, float *xyz_out, size_t cap_points, size_t *n_points);
/* pcl::io::savePCDFileBinary of a PointXYZ cloud (global_map_frames_manager.cpp:148): FIELDS x y z, 12 B/point. */
int ssf_pcd_write_binary(const char *path, const float *xyz, size_t n, size_t stride_bytes);
/* pcl::fromROSMsg (localization_node.cpp:290-291, map_data_save_node.cpp:66) for the float32 x / y / z fields
 * of a sensor_msgs/PointCloud2 byte buffer: n_points = width * height records of point_step bytes with the
 * fields at the given offsets -> n x float4 (w = 1), extracted on the device. */
int ssf_cloud_from_pointcloud2(ssf_ctx *ctx, const unsigned char *data, size_t n_points, size_t point_step,
                               size_t off_x, size_t off_y, size_t off_z, int is_bigendian, float *xyz_out);
/* The map cloud kept in HBM.  ssf_map_create uploads it once; ssf_map_from_pcd_folder is
 * GlobalMapFramesManager::getMapCloud (global_map_frames_manager.cpp:93-151): <folder>/<map_name>.pcd if it
 * exists (loaded as it is), else every *.pcd of the folder in readdir order, concatenated, pcl::VoxelGrid
 * (voxel_size) on the device, saved as <map_name>.pcd when save != 0 -- tiles go pinned host -> HBM, the
 * merged cloud never returns to the host. */
int ssf_map_create(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, ssf_map **out);
int ssf_map_from_pcd_folder(ssf_ctx *ctx, const char *data_folder, const char *map_name, float voxel_size, int save,
                            ssf_map **out);
void ssf_map_destroy(ssf_map *map);
size_t ssf_map_size(const ssf_map *map);
double ssf_map_ingest_ms(const ssf_map *map); /* stream ms of the last tiles -> HBM -> voxel filter (includes waits for file reads) */
double ssf_map_merge_ms(const ssf_map *map);  /* ... of which the voxel filter over the concatenated cloud (device ms) */
int ssf_map_download(ssf_map *map, float *xyz_out /* n x float4 */, size_t cap_points);
/* applyUniformSubsample(map_cloud_, step) (localization_node.cpp:20), in place in HBM. */
int ssf_map_subsample(ssf_map *map, size_t point_step);
/* cropPointCloudThroughRadius(map_T_sensor_, radius, map_cloud_, cropped) (localization_node.cpp:302) on the
 * resident map: same points, same order (ascending distance, ties by index) as ssf_cloud_crop_radius, with no
 * upload.  xyz_out (n x float4) / indices_out optional (NULL: count only). */
int ssf_map_crop_radius(ssf_map *map, const float center[3], double radius, float *xyz_out, size_t cap_points,
                        size_t *n_out, int32_t *indices_out);
/* The re-crop of localization_node.cpp:300-305 -- crop, then icp_->setTargetPointCloud(cropped) -- entirely in
 * HBM: a window change of the resident map, no 16 B/point round trip through the host. */
int ssf_map_crop_to_target(ssf_map *map, ssf_icp *icp, const float center[3], double radius, size_t *n_out);

/* ---- brute-force pose-grid alignment (localization/src/brute_force_alignment.cpp) ------------- */
/* setXYZStep / setXYZRange / setRotationStep / setRotationRange / setMeanErrorThreshold
 * (brute_force_alignment.cpp:12-42; node values at localization_node.cpp:38-43). */
typedef struct {
    float x_step, y_step, z_step;
    float x_range, y_range, z_range;
    float yaw_step, yaw_range;
    float mean_error_threshold;
} ssf_bfa_params;
/* Number of candidate poses of createTestTransformSequences (brute_force_alignment.cpp:148-180). */
size_t ssf_bfa_pose_count(const ssf_bfa_params *params);
/* BruteForceAlignment::alignClouds (brute_force_alignment.cpp:65-136) against the handle's target:
 * every candidate prev * T(x,y,z,yaw) is scored by the mean squared distance of the n source
 * points to their nearest target point (unbounded search).  *success = 1 and T_best = the FIRST
 * candidate in the reference's loop order whose score is below the threshold; otherwise
 * *success = 0 and T_best = the best-scoring candidate (the reference's next starting pose).
 * scores_out (optional, ssf_bfa_pose_count floats): every candidate's score, loop order. */
int ssf_bfa_align(ssf_icp *icp, const float *src_xyz, size_t n, size_t stride_bytes, const float T_prev_colmajor[16],
                  const ssf_bfa_params *params, float T_best_colmajor[16], float *best_score, int *success,
                  float *scores_out);

/* ---- batches: offline reprocessing of scan sequences (BASELINE.json config 4) --------- */
/* Every scan of a batch is aligned against the handle's target with the handle's
 * parameters; scans are independent (the per-scan loop of
 * localization/src/localization_node.cpp:335-338 over a recorded sequence). */
int ssf_batch_create(ssf_icp *icp, size_t max_scans, size_t max_total_points, ssf_batch **out);
void ssf_batch_destroy(ssf_batch *b);
/* Copy n_scans scans to HBM.  xyz: concatenated points of all scans; n_pts[s] their sizes. */
int ssf_batch_upload(ssf_batch *b, const float *xyz, const size_t *n_pts, size_t n_scans, size_t stride_bytes);
/* Same, but returns without waiting for the copy: the caller's buffer (pinned memory for a true
 * overlap) must stay untouched until ssf_batch_results of this batch returns.  Uploads run on a
 * per-batch copy stream, so the upload of one batch overlaps the alignment of another. */
int ssf_batch_upload_async(ssf_batch *b, const float *xyz, const size_t *n_pts, size_t n_scans, size_t stride_bytes);
/* Initial transforms, n_scans x 16 floats column-major. */
int ssf_batch_set_initial(ssf_batch *b, const float *T_colmajor);
/* Run the whole batch on the device; asynchronous on the context stream. */
int ssf_batch_run(ssf_batch *b);
/* Wait and copy the n_scans results back. */
int ssf_batch_results(ssf_batch *b, ssf_icp_result *out, size_t n_scans);
/* Search statistics of the last run (GN / O3D modes): per search launch, the queries it answered and
 * how many of them needed a walk of the index (the others were confirmed by their search certificate).
 * At most cap entries are written; *n_launches = launches of the run. */
int ssf_batch_search_stats(ssf_batch *b, uint64_t *answered, uint64_t *walked, size_t cap, size_t *n_launches);
/* upload + set_initial + run + results in one call (host buffers in, host results out). */
int ssf_icp_align_batch(ssf_icp *icp, const float *xyz, const size_t *n_pts, size_t n_scans, size_t stride_bytes,
                        const float *T_colmajor, ssf_icp_result *out);

/* ---- map sharding across GPUs (BASELINE.json configs 3 and 5) --------------------------- */
/* One process per GPU holds one spatial shard of the map.  All ranks agree on a global grid
 * (origin = global minimum corner, cell_size >= sqrt(max_correspondence_dist) * 1.01) and on a
 * partition of its cell COLUMNS (x index) into ranges [own_lo, own_hi).  A rank's shard must
 * contain every map point whose column lies in [own_lo - 1, own_hi + 1) (one-cell halo), so the
 * exact neighbour of every query it owns is local.  Scans are replicated; a rank searches only
 * the queries whose column it owns, and the per-scan sums (n_scans x 32 doubles) are all-reduced
 * once per iteration through the caller's hook before every rank runs the identical solve.
 * global_index (optional): index of each shard point in the unsharded cloud, used for tie-breaks
 * and reported correspondences.  GN and O3D modes only. */
typedef struct {
    float origin[3];
    float cell_size;
    int32_t own_lo, own_hi;
} ssf_shard_info;
int ssf_icp_set_target_shard(ssf_icp *icp, const float *xyz, size_t n, size_t stride_bytes, const float *normals,
                             size_t normals_stride_bytes, const int32_t *global_index, const ssf_shard_info *info);
/* Sum `count` doubles at DEVICE pointer `buf` across ranks, in place, ordered on `cuda_stream`.
 * Return 0 on success.  (bench.py passes torch.distributed.all_reduce over NCCL.) */
typedef int (*ssf_allreduce_fn)(void *user, double *buf, size_t count, void *cuda_stream);
int ssf_icp_set_allreduce(ssf_icp *icp, ssf_allreduce_fn fn, void *user);
/* The same sum through NCCL, with no callback: ncclAllReduce (n_scans x 32 doubles) on the library's own
 * stream, inside the CUDA graph of the loop.  libnccl.so.2 is bound at run time.  Rank 0 calls
 * ssf_nccl_unique_id, the caller carries the 128 bytes to every rank (any transport), and every rank
 * calls ssf_icp_nccl_init -- collectively, like ncclCommInitRank.  Takes precedence over the hook. */
#define SSF_NCCL_ID_BYTES 128
int ssf_nccl_unique_id(unsigned char id_out[SSF_NCCL_ID_BYTES]);
int ssf_icp_nccl_init(ssf_icp *icp, const unsigned char id[SSF_NCCL_ID_BYTES], int rank, int world);
int ssf_icp_nccl_close(ssf_icp *icp);
/* In-kernel exchange instead of the hook (one process per GPU, all on one NVLink box): every rank
 * creates its exchange buffer and gets a handle blob (CUDA IPC handle + identity), the caller gathers the blobs of
 * all ranks (any transport: torch.distributed, MPI, a file) and hands the rank-ordered array
 * (world x SSF_XCH_HANDLE_BYTES) to ssf_icp_exchange_open.  From then on ONE kernel per iteration sums a scan's partial rows, stores the row
 * straight into every rank's buffer (peer stores over NVLink, per-scan release flag), waits for the same
 * scan's rows of the other ranks, adds them in rank order and solves -- no host call and no NCCL
 * collective per iteration; the loop stays one CUDA graph.  All ranks must run the same sequence of alignments with the same
 * scans; max_scans bounds the scans per batch.  world <= 32. */
#define SSF_XCH_HANDLE_BYTES 128 /* CUDA IPC handle + device UUID + (max_scans, world, rank) */
int ssf_icp_exchange_create(ssf_icp *icp, int rank, int world, size_t max_scans,
                            unsigned char handle_out[SSF_XCH_HANDLE_BYTES]);
/* SSF_ERR_COMM when two ranks share a device or the ranks' (world, max_scans) differ.  During a run a rank
 * waits at most 20 s for a peer's sums, and rejects sums of another batch shape (scans, iterations, mode):
 * the scan's result then carries aborted == 2 and ssf_batch_results returns SSF_ERR_COMM. */
int ssf_icp_exchange_open(ssf_icp *icp, const unsigned char *handles /* world x SSF_XCH_HANDLE_BYTES, rank order */);
int ssf_icp_exchange_close(ssf_icp *icp);

/* ---- profiling hook ------------------------------------------------------------------- */
/* When enabled, every launch of the NN-search kernels (K3) on this context is bracketed by a
 * pair of CUDA events on the context stream.  ssf_ctx_search_time waits for the stream, adds
 * the elapsed times up and clears the list: *ms = total, *launches = how many. */
int ssf_ctx_time_searches(ssf_ctx *ctx, int enable);
int ssf_ctx_search_time(ssf_ctx *ctx, double *ms, uint64_t *launches);
/* Per-launch elapsed times (ms) of the recorded launches, in launch order, without clearing the
 * list: at most cap values are written, *launches = how many were recorded. */
int ssf_ctx_search_times(ssf_ctx *ctx, float *ms_out, uint64_t cap, uint64_t *launches);

/* ---- counters ------------------------------------------------------------------------- */
/* Kernels launched by this library since process start (all contexts). */
uint64_t ssf_kernel_launches(void);
/* NN queries issued by search kernels since process start. */
uint64_t ssf_nn_queries(void);

#ifdef __cplusplus
}
#endif
#endif /* SSF_SSF_H */
