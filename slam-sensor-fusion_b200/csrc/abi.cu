// abi.cu -- the extern "C" surface declared in include/ssf/ssf.h.
//
// Host-side orchestration only: handles own HBM buffers, copy caller data in (deep copy at
// set time, like reference localization/src/icp_point_to_point.cpp:44-55), enqueue the
// kernels of map_build.cu / icp_kernels.cu / voxel_grid.cu on the context stream and copy
// the small results out.  There is no CPU implementation of any step behind these calls.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "bfa.cuh"
#include "icp.cuh"
#include "ingest.cuh"
#include "map_index.cuh"
#include "nccl_link.cuh"
#include "preprocess.cuh"
#include "voxel_grid.cuh"

namespace ssf {

std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_queries{0};
static thread_local char t_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

}  // namespace ssf

using namespace ssf;

struct ssf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Scratch scratch;
    DevBuf<unsigned char> stage;  // raw bytes of caller clouds before packing
    SearchTimer timer;
};

struct ssf_batch {
    ssf_icp *icp = nullptr;
    BatchBuffers buf;
    size_t max_scans = 0, max_points = 0;
    PinnedBuf<float> T_init_pinned;  // initial transforms: pinned host memory the device reads directly (see set_initial)
    DevBuf<uint32_t> meta_dev;  // per scan: raw_begin, n_raw, pt_begin, tile_begin, tile_cap
    std::vector<uint32_t> meta_host;
    PinnedBuf<uint32_t> meta_pinned;     // what the device copy reads (pageable sources make cudaMemcpyAsync synchronise)
    PinnedBuf<ssf_icp_result> res_pinned;
    std::vector<uint32_t> n_raw;
    size_t total_points = 0;
    bool uploaded = false, initial_set = false, ran = false;
    bool ever_ran = false;  // ran_ev has been recorded at least once
    bool uploaded_raw = false;  // scans went to buf.raw (voxel stage in front of the loop)
    float last_ms = 0.f;
    // copies run on their own stream so the upload of one batch overlaps the alignment of another
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t uploaded_ev = nullptr, ran_ev = nullptr, ev0 = nullptr, ev1 = nullptr;
    DevBuf<unsigned char> stage;  // raw bytes of the caller's scans before packing
    cudaEvent_t dbg[4] = {nullptr, nullptr, nullptr, nullptr};  // SSF_DEBUG_TIMING: H2D begin/end, pack end
};

struct ssf_icp {
    ssf_ctx *ctx = nullptr;
    ssf_icp_params prm{};
    MapIndex map;
    bool has_target = false;
    float T_init[16];
    ssf_batch *single = nullptr;  // the one-scan batch behind set_source / align
    bool has_source = false;
    size_t n_source = 0;
    bool in_shard_call = false;
    AllreduceFn allreduce = nullptr;  // map sharding hook
    void *allreduce_user = nullptr;
    // map sharding, in-kernel exchange (XchView in icp.cuh)
    struct {
        int rank = 0, world = 0;
        size_t max_scans = 0, bytes = 0;
        void *local = nullptr;             // this rank's buffer (cudaMalloc)
        std::vector<void *> peers;         // every rank's buffer as seen from this device
        std::vector<bool> opened;          // peers[r] came from cudaIpcOpenMemHandle
        DevBuf<void *> peers_dev;
        DevBuf<unsigned long long> epoch_dev;  // epoch of the next run's first pass (flags start at 0); advanced on the device
        bool ready = false;
    } xch;
    NcclLink *nccl = nullptr;  // map sharding through ncclAllReduce (nccl_link.cu)
    DevBuf<float4> q_dev;  // ssf_nn_search temporaries
    DevBuf<int32_t> q_idx;
    DevBuf<float> q_d2;
};

#define SSF_ARG(cond, msg)                 \
    do {                                   \
        if (!(cond)) {                     \
            ssf::set_error("%s", msg);     \
            return SSF_ERR_INVALID;        \
        }                                  \
    } while (0)

static int use_device(const ssf_ctx *ctx)
{
    SSF_CUDA(cudaSetDevice(ctx->device));
    return SSF_OK;
}

// ---- packing kernels ------------------------------------------------------------------------
// caller cloud (floats at a byte stride) -> float4
__global__ void __launch_bounds__(256) pack_points_kernel(const unsigned char *__restrict__ raw, size_t n,
                                                          size_t stride_bytes, float4 *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = reinterpret_cast<const float *>(raw + i * stride_bytes);
    out[i] = make_float4(p[0], p[1], p[2], 1.0f);
}

// concatenated scans -> tile-aligned per-scan slots; one block column per scan
__global__ void __launch_bounds__(256)
    pack_scans_kernel(const unsigned char *__restrict__ raw, size_t stride_bytes, const uint32_t *__restrict__ meta,
                      float4 *__restrict__ out)
{
    const uint32_t s = blockIdx.y;
    const uint32_t raw_begin = meta[5 * s + 0], n = meta[5 * s + 1], pt_begin = meta[5 * s + 2];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *p = reinterpret_cast<const float *>(raw + (size_t)(raw_begin + i) * stride_bytes);
        out[(size_t)pt_begin + i] = make_float4(p[0], p[1], p[2], 1.0f);
    }
}

__global__ void layout_kernel(ScanState *st, const uint32_t *__restrict__ meta, uint32_t n_scans,
                              uint32_t *__restrict__ tile_scan)
{
    const uint32_t s = blockIdx.x;
    if (s >= n_scans) return;
    const uint32_t n = meta[5 * s + 1], pt_begin = meta[5 * s + 2], tile_begin = meta[5 * s + 3],
                   tile_cap = meta[5 * s + 4];
    if (threadIdx.x == 0) {
        st[s].pt_begin = pt_begin;
        st[s].n_pts = n;
        st[s].tile_begin = tile_begin;
        st[s].tile_cap = tile_cap;
    }
    for (uint32_t t = threadIdx.x; t < tile_cap; t += blockDim.x) tile_scan[tile_begin + t] = s;
}

// ---- misc -------------------------------------------------------------------------------------
extern "C" const char *ssf_last_error(void) { return t_err; }
extern "C" const char *ssf_version(void) { return "ssf-gpu 0.1 (sm_100a)"; }
extern "C" uint64_t ssf_kernel_launches(void) { return g_launches.load(); }
extern "C" uint64_t ssf_nn_queries(void) { return g_queries.load(); }

// ---- context ----------------------------------------------------------------------------------
extern "C" int ssf_ctx_create(int device_ordinal, ssf_ctx **out)
{
    SSF_ARG(out, "ssf_ctx_create: out == NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no usable CUDA device (%s); libssf_gpu has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return SSF_ERR_CUDA;
    }
    SSF_ARG(device_ordinal >= 0 && device_ordinal < count, "ssf_ctx_create: device ordinal out of range");
    ssf_ctx *c = new (std::nothrow) ssf_ctx;
    if (!c) return SSF_ERR_NOMEM;
    c->device = device_ordinal;
    SSF_CUDA(cudaSetDevice(device_ordinal));
    SSF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SSF_CUDA(cudaEventCreate(&c->ev0));
    SSF_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    return SSF_OK;
}

extern "C" void ssf_ctx_destroy(ssf_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->timer.pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int ssf_ctx_synchronize(ssf_ctx *ctx)
{
    SSF_ARG(ctx, "ssf_ctx_synchronize: ctx == NULL");
    SSF_TRY(use_device(ctx));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSF_OK;
}

extern "C" void *ssf_ctx_stream(ssf_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int ssf_ctx_time_searches(ssf_ctx *ctx, int enable)
{
    SSF_ARG(ctx, "ssf_ctx_time_searches: ctx == NULL");
    SSF_TRY(use_device(ctx));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->timer.enabled = enable != 0;
    ctx->timer.used = 0;
    return SSF_OK;
}

extern "C" int ssf_ctx_search_time(ssf_ctx *ctx, double *ms, uint64_t *launches)
{
    SSF_ARG(ctx && ms && launches, "ssf_ctx_search_time: NULL argument");
    SSF_TRY(use_device(ctx));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < ctx->timer.used; i += 2) {
        float t = 0.f;
        SSF_CUDA(cudaEventElapsedTime(&t, ctx->timer.pool[i], ctx->timer.pool[i + 1]));
        tot += t;
    }
    *ms = tot;
    *launches = ctx->timer.used / 2;
    ctx->timer.used = 0;
    return SSF_OK;
}

extern "C" int ssf_ctx_search_times(ssf_ctx *ctx, float *ms_out, uint64_t cap, uint64_t *launches)
{
    SSF_ARG(ctx && launches && (ms_out || cap == 0), "ssf_ctx_search_times: NULL argument");
    SSF_TRY(use_device(ctx));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t n = 0;
    for (size_t i = 0; i + 1 < ctx->timer.used; i += 2, ++n) {
        if (n >= cap) continue;
        SSF_CUDA(cudaEventElapsedTime(&ms_out[n], ctx->timer.pool[i], ctx->timer.pool[i + 1]));
    }
    *launches = n;
    return SSF_OK;
}

// ---- helpers ------------------------------------------------------------------------------------
static int check_params(const ssf_icp_params *p)
{
    SSF_ARG(p, "params == NULL");
    SSF_ARG(p->num_iterations >= 0 && p->num_iterations <= 100000, "num_iterations out of range");
    SSF_ARG(p->mode >= SSF_MODE_REFERENCE && p->mode <= SSF_MODE_O3D_P2P, "unknown mode");
    SSF_ARG(p->reduce == SSF_REDUCE_STRICT || p->reduce == SSF_REDUCE_FAST, "unknown reduce");
    SSF_ARG(!(p->max_correspondence_dist != p->max_correspondence_dist), "max_correspondence_dist is NaN");
    SSF_ARG(!(p->source_voxel_leaf < 0.f), "source_voxel_leaf < 0");
    return SSF_OK;
}

// copy n points (stride) from the host into ctx->stage and pack them as float4 into dst
static int upload_cloud(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, float4 *dst)
{
    if (n == 0) return SSF_OK;
    SSF_ARG(xyz, "cloud pointer == NULL");
    SSF_ARG(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride_bytes must be a multiple of 4 and >= 12");
    const size_t bytes = (n - 1) * stride_bytes + 12;
    SSF_TRY(ctx->stage.reserve(bytes));
    SSF_CUDA(cudaMemcpyAsync(ctx->stage.p, xyz, bytes, cudaMemcpyHostToDevice, ctx->stream));
    pack_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->stage.p, n, stride_bytes, dst);
    SSF_LAUNCHED();
    return SSF_OK;
}

// ---- access for ingest.cu ---------------------------------------------------------------------------
int ssf::ssf_ctx_ref::use() const { return use_device(c); }
cudaStream_t ssf::ssf_ctx_ref::stream() const { return c->stream; }
ssf::Scratch &ssf::ssf_ctx_ref::scratch() const { return c->scratch; }
int ssf::ssf_ctx_ref::upload_cloud(const float *xyz, size_t n, size_t stride_bytes, float4 *dst) const
{
    return ::upload_cloud(c, xyz, n, stride_bytes, dst);
}

// ---- registration object ------------------------------------------------------------------------
extern "C" int ssf_icp_create(ssf_ctx *ctx, const ssf_icp_params *params, ssf_icp **out)
{
    SSF_ARG(ctx && out, "ssf_icp_create: NULL argument");
    *out = nullptr;
    SSF_TRY(check_params(params));
    ssf_icp *h = new (std::nothrow) ssf_icp;
    if (!h) return SSF_ERR_NOMEM;
    h->ctx = ctx;
    h->prm = *params;
    for (int i = 0; i < 16; ++i) h->T_init[i] = (i % 5 == 0) ? 1.f : 0.f;  // Matrix4f::Identity (cpp:11)
    *out = h;
    return SSF_OK;
}

static void exchange_release(ssf_icp *icp);

extern "C" void ssf_icp_destroy(ssf_icp *icp)
{
    if (!icp) return;
    cudaSetDevice(icp->ctx->device);
    cudaStreamSynchronize(icp->ctx->stream);
    if (icp->single) ssf_batch_destroy(icp->single);
    exchange_release(icp);
    nccl_link_destroy(icp->nccl);
    delete icp;
}

extern "C" int ssf_icp_set_params(ssf_icp *icp, const ssf_icp_params *params)
{
    SSF_ARG(icp, "ssf_icp_set_params: icp == NULL");
    SSF_TRY(check_params(params));
    icp->prm = *params;
    return SSF_OK;
}

extern "C" int ssf_icp_get_params(const ssf_icp *icp, ssf_icp_params *out)
{
    SSF_ARG(icp && out, "ssf_icp_get_params: NULL argument");
    *out = icp->prm;
    return SSF_OK;
}

extern "C" size_t ssf_icp_target_size(const ssf_icp *icp) { return icp ? icp->map.n_raw : 0; }

extern "C" int ssf_icp_set_target(ssf_icp *icp, const float *xyz, size_t n, size_t stride_bytes, const float *normals,
                                  size_t normals_stride_bytes)
{
    SSF_ARG(icp, "ssf_icp_set_target: icp == NULL");
    SSF_ARG(n == 0 || xyz, "ssf_icp_set_target: xyz == NULL");
    SSF_ARG(n < ((size_t)1 << 31), "ssf_icp_set_target: more than 2^31 - 1 points");
    ssf_ctx *ctx = icp->ctx;
    SSF_TRY(use_device(ctx));
    icp->has_target = false;
    MapIndex &m = icp->map;
    if (!icp->in_shard_call) {
        m.sharded = false;
        m.has_global_index = false;
        m.own_lo = INT32_MIN;
        m.own_hi = INT32_MAX;
    }
    m.n_raw = n;
    m.has_normals = normals != nullptr && n > 0;
    SSF_TRY(m.raw.reserve(n ? n : 1));
    SSF_TRY(upload_cloud(ctx, xyz, n, stride_bytes, m.raw.p));
    if (m.has_normals) {
        SSF_TRY(m.raw_nrm.reserve(n));
        SSF_CUDA(cudaStreamSynchronize(ctx->stream));  // stage buffer is reused
        SSF_TRY(upload_cloud(ctx, normals, n, normals_stride_bytes, m.raw_nrm.p));
    }
    // the cell edge is chosen from the cloud's own density (map_build.cu), not from the threshold
    SSF_TRY(build_map_index(m, 0.f, ctx->scratch, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    icp->has_target = true;
    return SSF_OK;
}

int ssf::icp_set_target_device(ssf_icp *icp, const float4 *pts_dev, size_t n, const ssf_ctx_ref &from)
{
    SSF_ARG(icp && (pts_dev || n == 0), "set target from device: NULL argument");
    SSF_ARG(n < ((size_t)1 << 31), "set target from device: more than 2^31 - 1 points");
    ssf_ctx *ctx = icp->ctx;
    SSF_ARG(from.c->device == ctx->device, "the resident map and the registration object are on different devices");
    SSF_TRY(use_device(ctx));
    icp->has_target = false;
    MapIndex &m = icp->map;
    m.sharded = false;
    m.has_global_index = false;
    m.own_lo = INT32_MIN;
    m.own_hi = INT32_MAX;
    m.n_raw = n;
    m.has_normals = false;
    SSF_TRY(m.raw.reserve(n ? n : 1));
    if (from.c->stream != ctx->stream) SSF_CUDA(cudaStreamSynchronize(from.c->stream));  // the crop is complete
    if (n) SSF_CUDA(cudaMemcpyAsync(m.raw.p, pts_dev, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    SSF_TRY(build_map_index(m, 0.f, ctx->scratch, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    icp->has_target = true;
    return SSF_OK;
}

extern "C" int ssf_icp_set_target_shard(ssf_icp *icp, const float *xyz, size_t n, size_t stride_bytes,
                                        const float *normals, size_t normals_stride_bytes, const int32_t *global_index,
                                        const ssf_shard_info *info)
{
    SSF_ARG(icp && info, "ssf_icp_set_target_shard: NULL argument");
    SSF_ARG(info->own_lo <= info->own_hi, "ssf_icp_set_target_shard: own_lo > own_hi");
    SSF_ARG(info->cell_size > 0.f, "ssf_icp_set_target_shard: cell_size must be > 0");
    MapIndex &m = icp->map;
    SSF_TRY(use_device(icp->ctx));
    m.sharded = true;
    memcpy(m.shard_origin, info->origin, sizeof(m.shard_origin));
    m.own_lo = info->own_lo;
    m.own_hi = info->own_hi;
    m.has_global_index = global_index != nullptr && n > 0;
    if (m.has_global_index) {
        SSF_TRY(m.global_index.reserve(n));
        SSF_CUDA(cudaMemcpyAsync(m.global_index.p, global_index, n * sizeof(int32_t), cudaMemcpyHostToDevice,
                                 icp->ctx->stream));
        SSF_CUDA(cudaStreamSynchronize(icp->ctx->stream));
    }
    m.shard_cell = info->cell_size;
    icp->in_shard_call = true;
    int rc = ssf_icp_set_target(icp, xyz, n, stride_bytes, normals, normals_stride_bytes);
    icp->in_shard_call = false;
    if (rc != SSF_OK) m.sharded = false;
    return rc;
}

extern "C" int ssf_icp_set_allreduce(ssf_icp *icp, ssf_allreduce_fn fn, void *user)
{
    SSF_ARG(icp, "ssf_icp_set_allreduce: icp == NULL");
    icp->allreduce = fn;
    icp->allreduce_user = user;
    return SSF_OK;
}

static void exchange_release(ssf_icp *icp)
{
    auto &x = icp->xch;
    for (size_t r = 0; r < x.peers.size(); ++r)
        if (x.opened[r] && x.peers[r]) cudaIpcCloseMemHandle(x.peers[r]);
    x.peers.clear();
    x.opened.clear();
    if (x.local) cudaFree(x.local);
    x.local = nullptr;
    x.ready = false;
    x.world = 0;
}

extern "C" int ssf_icp_exchange_create(ssf_icp *icp, int rank, int world, size_t max_scans,
                                       unsigned char handle_out[SSF_XCH_HANDLE_BYTES])
{
    SSF_ARG(icp && handle_out, "ssf_icp_exchange_create: NULL argument");
    SSF_ARG(world >= 1 && world <= 32 && rank >= 0 && rank < world, "ssf_icp_exchange_create: bad rank / world (<= 32)");
    SSF_ARG(max_scans >= 1 && max_scans < (1u << 24), "ssf_icp_exchange_create: max_scans out of range");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SSF_TRY(use_device(icp->ctx));
    exchange_release(icp);
    auto &x = icp->xch;
    x.rank = rank;
    x.world = world;
    x.max_scans = max_scans;
    x.bytes = (size_t)2 * world * max_scans * kAccum * sizeof(double) + (size_t)4 * world * max_scans * sizeof(unsigned long long);
    SSF_CUDA(cudaMalloc(&x.local, x.bytes));
    SSF_CUDA(cudaMemset(x.local, 0, x.bytes));
    SSF_TRY(x.epoch_dev.reserve(1));
    const unsigned long long one = 1;
    SSF_CUDA(cudaMemcpy(x.epoch_dev.p, &one, sizeof(one), cudaMemcpyHostToDevice));
    SSF_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    SSF_CUDA(cudaIpcGetMemHandle(&h, x.local));
    // blob: [0..63] IPC handle, [64..79] device UUID, [80..87] max_scans, [88..91] world, [92..95] rank
    cudaDeviceProp prop;
    SSF_CUDA(cudaGetDeviceProperties(&prop, icp->ctx->device));
    memset(handle_out, 0, SSF_XCH_HANDLE_BYTES);
    memcpy(handle_out, &h, 64);
    memcpy(handle_out + 64, &prop.uuid, 16);
    const unsigned long long ms64 = max_scans;
    const int32_t w32 = world, r32 = rank;
    memcpy(handle_out + 80, &ms64, 8);
    memcpy(handle_out + 88, &w32, 4);
    memcpy(handle_out + 92, &r32, 4);
    return SSF_OK;
}

extern "C" int ssf_icp_exchange_open(ssf_icp *icp, const unsigned char *handles)
{
    SSF_ARG(icp && handles, "ssf_icp_exchange_open: NULL argument");
    auto &x = icp->xch;
    if (!x.local || x.world < 1) {
        set_error("ssf_icp_exchange_open: call ssf_icp_exchange_create first");
        return SSF_ERR_STATE;
    }
    SSF_TRY(use_device(icp->ctx));
    // every rank must sit on its own device (two ranks on one GPU would be mutually waiting launches)
    // and must have created the exchange with the same shape
    for (int r = 0; r < x.world; ++r) {
        const unsigned char *hr = handles + (size_t)SSF_XCH_HANDLE_BYTES * r;
        unsigned long long ms64 = 0;
        int32_t w32 = 0, r32 = -1;
        memcpy(&ms64, hr + 80, 8);
        memcpy(&w32, hr + 88, 4);
        memcpy(&r32, hr + 92, 4);
        if (ms64 != x.max_scans || w32 != x.world || r32 != r) {
            set_error("ssf_icp_exchange_open: handle %d was created as rank %d of %d with max_scans %llu (here: world %d, "
                      "max_scans %zu)", r, (int)r32, (int)w32, ms64, x.world, x.max_scans);
            return SSF_ERR_COMM;
        }
        for (int q = 0; q < r; ++q)
            if (memcmp(hr + 64, handles + (size_t)SSF_XCH_HANDLE_BYTES * q + 64, 16) == 0) {
                set_error("ssf_icp_exchange_open: ranks %d and %d are on the same GPU; the in-kernel exchange needs one "
                          "device per rank", q, r);
                return SSF_ERR_COMM;
            }
    }
    x.peers.assign(x.world, nullptr);
    x.opened.assign(x.world, false);
    for (int r = 0; r < x.world; ++r) {
        if (r == x.rank) {
            x.peers[r] = x.local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)SSF_XCH_HANDLE_BYTES * r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&x.peers[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return SSF_ERR_COMM;
        }
        x.opened[r] = true;
    }
    SSF_TRY(x.peers_dev.reserve(x.world));
    SSF_CUDA(cudaMemcpy(x.peers_dev.p, x.peers.data(), x.world * sizeof(void *), cudaMemcpyHostToDevice));
    x.ready = true;
    return SSF_OK;
}

extern "C" int ssf_icp_exchange_close(ssf_icp *icp)
{
    SSF_ARG(icp, "ssf_icp_exchange_close: icp == NULL");
    SSF_TRY(use_device(icp->ctx));
    cudaStreamSynchronize(icp->ctx->stream);
    exchange_release(icp);
    return SSF_OK;
}

extern "C" int ssf_nccl_unique_id(unsigned char id_out[SSF_NCCL_ID_BYTES])
{
    SSF_ARG(id_out, "ssf_nccl_unique_id: NULL argument");
    return nccl_link_unique_id(id_out);
}

extern "C" int ssf_icp_nccl_init(ssf_icp *icp, const unsigned char id[SSF_NCCL_ID_BYTES], int rank, int world)
{
    SSF_ARG(icp && id, "ssf_icp_nccl_init: NULL argument");
    SSF_ARG(world >= 1 && rank >= 0 && rank < world, "ssf_icp_nccl_init: bad rank / world");
    SSF_TRY(use_device(icp->ctx));
    nccl_link_destroy(icp->nccl);
    icp->nccl = nullptr;
    return nccl_link_create(id, rank, world, &icp->nccl);
}

extern "C" int ssf_icp_nccl_close(ssf_icp *icp)
{
    SSF_ARG(icp, "ssf_icp_nccl_close: icp == NULL");
    SSF_TRY(use_device(icp->ctx));
    cudaStreamSynchronize(icp->ctx->stream);
    nccl_link_destroy(icp->nccl);
    icp->nccl = nullptr;
    return SSF_OK;
}

extern "C" int ssf_icp_set_initial(ssf_icp *icp, const float T_colmajor[16])
{
    SSF_ARG(icp && T_colmajor, "ssf_icp_set_initial: NULL argument");
    memcpy(icp->T_init, T_colmajor, sizeof(icp->T_init));
    return SSF_OK;
}

extern "C" int ssf_icp_set_source(ssf_icp *icp, const float *xyz, size_t n, size_t stride_bytes)
{
    SSF_ARG(icp, "ssf_icp_set_source: icp == NULL");
    SSF_ARG(n == 0 || xyz, "ssf_icp_set_source: xyz == NULL");
    SSF_TRY(use_device(icp->ctx));
    if (!icp->single || icp->single->max_points < n) {
        if (icp->single) ssf_batch_destroy(icp->single);
        icp->single = nullptr;
        SSF_TRY(ssf_batch_create(icp, 1, n + n / 4 + 1024, &icp->single));
    }
    const size_t cnt = n;
    SSF_TRY(ssf_batch_upload(icp->single, xyz, &cnt, 1, stride_bytes ? stride_bytes : 16));
    icp->has_source = true;
    icp->n_source = n;
    return SSF_OK;
}

// setDebugMode(true): the text printStepDebug (cpp:172-183) and the tail of calculateAlignment
// (cpp:237-246) write to stdout, from the per-pass error trace of the finished alignment.
static void debug_print(ssf_icp *icp, const ssf_icp_result &out)
{
    const ssf_icp_params &p = icp->prm;
    const size_t n = (size_t)icp->single->buf.trace_len;
    std::vector<float> err(n ? n : 1);
    if (n == 0 || ssf_icp_get_trace(icp, err.data(), nullptr, n) != SSF_OK) return;
    const bool ref = p.mode == SSF_MODE_REFERENCE;
    float last_error = 3.402823466e+38f;  // cpp:205
    for (size_t i = 0; i < n; ++i) {
        if (err[i] != err[i]) break;
        printf("[ICP INFO] Iteration %zu - Error: %g\n", i, err[i]);
        if (err[i] < p.acceptable_mean_error) printf("[ICP INFO] Acceptable error reached. Stopping iterations.\n");
        // REFERENCE: the test of cpp:179 against the previous pass's error; GN: the pose update of this
        // pass was below the epsilon (the stop rule of those modes)
        const bool eps_hit = ref ? fabsf(last_error - err[i]) < p.transformation_epsilon
                                 : (out.has_converged && (int)i + 1 == out.iterations && !(err[i] < p.acceptable_mean_error));
        if (eps_hit) printf("[ICP INFO] Transformation epsilon reached. Stopping iterations.\n");
        last_error = err[i];
    }
    if (out.iterations == p.num_iterations)
        printf("[ICP INFO] We reached the maximum number of iterations. Returning best transform found.\n");
    printf("[ICP INFO] Total iterations taken: %d\n[ICP INFO] Final error: %g\n", out.iterations, out.error);
    // operator<< of Eigen::Matrix4f (default IOFormat): %g coefficients right-aligned to the widest one
    char cell[16][32];
    int wid = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            const int len = snprintf(cell[r * 4 + c], sizeof(cell[0]), "%g", out.transformation[c * 4 + r]);
            if (len > wid) wid = len;
        }
    printf("[ICP INFO] Final transformation matrix: \n");
    for (int r = 0; r < 4; ++r)
        printf("%*s %*s %*s %*s\n", wid, cell[r * 4], wid, cell[r * 4 + 1], wid, cell[r * 4 + 2], wid, cell[r * 4 + 3]);
    fflush(stdout);
}

extern "C" int ssf_icp_align(ssf_icp *icp, ssf_icp_result *out)
{
    SSF_ARG(icp && out, "ssf_icp_align: NULL argument");
    if (!icp->has_target) {
        set_error("ssf_icp_align: no target set (call ssf_icp_set_target first)");
        return SSF_ERR_STATE;
    }
    if (!icp->has_source) {
        set_error("ssf_icp_align: no source set (call ssf_icp_set_source first)");
        return SSF_ERR_STATE;
    }
    SSF_TRY(ssf_batch_set_initial(icp->single, icp->T_init));
    SSF_TRY(ssf_batch_run(icp->single));
    SSF_TRY(ssf_batch_results(icp->single, out, 1));
    if (out->aborted == 1) fprintf(stderr, "[ICP ERROR] Not enough valid correspondences found. Aborting.\n");  // cpp:198
    if (icp->prm.debug && out->aborted == 0) debug_print(icp, *out);
    return SSF_OK;
}

extern "C" int ssf_icp_get_correspondences(ssf_icp *icp, int32_t *idx_out, size_t n)
{
    SSF_ARG(icp && (idx_out || n == 0), "ssf_icp_get_correspondences: NULL argument");
    if (!icp->single || !icp->single->ran) {
        set_error("ssf_icp_get_correspondences: no alignment has run");
        return SSF_ERR_STATE;
    }
    SSF_ARG(n <= icp->single->n_raw[0], "ssf_icp_get_correspondences: n larger than the source");
    SSF_TRY(use_device(icp->ctx));
    if (n == 0) return SSF_OK;
    SSF_CUDA(cudaMemcpyAsync(idx_out, icp->single->buf.corr.p + icp->single->meta_host[2], n * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, icp->ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(icp->ctx->stream));
    return SSF_OK;
}

extern "C" int ssf_icp_get_trace(ssf_icp *icp, float *iter_err, int32_t *iter_searched, size_t n)
{
    SSF_ARG(icp, "ssf_icp_get_trace: icp == NULL");
    if (!icp->single || !icp->single->ran) {
        set_error("ssf_icp_get_trace: no alignment has run");
        return SSF_ERR_STATE;
    }
    SSF_TRY(use_device(icp->ctx));
    const size_t have = (size_t)icp->single->buf.trace_len;
    for (size_t i = 0; i < n; ++i) {
        if (iter_err) iter_err[i] = NAN;
        if (iter_searched) iter_searched[i] = 0;
    }
    const size_t m = n < have ? n : have;
    if (m == 0) return SSF_OK;
    if (iter_err)
        SSF_CUDA(cudaMemcpyAsync(iter_err, icp->single->buf.trace_err.p, m * sizeof(float), cudaMemcpyDeviceToHost,
                                 icp->ctx->stream));
    if (iter_searched)
        SSF_CUDA(cudaMemcpyAsync(iter_searched, icp->single->buf.trace_search.p, m * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, icp->ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(icp->ctx->stream));
    return SSF_OK;
}

extern "C" int ssf_nn_search(ssf_icp *icp, const float *queries, size_t n, size_t stride_bytes, float max_sqdist,
                             int32_t *idx, float *d2)
{
    SSF_ARG(icp, "ssf_nn_search: icp == NULL");
    SSF_ARG(n == 0 || (queries && idx && d2), "ssf_nn_search: NULL argument");
    if (!icp->has_target) {
        set_error("ssf_nn_search: no target set");
        return SSF_ERR_STATE;
    }
    if (n == 0) return SSF_OK;
    ssf_ctx *ctx = icp->ctx;
    SSF_TRY(use_device(ctx));
    SSF_TRY(icp->q_dev.reserve(n));
    SSF_TRY(icp->q_idx.reserve(n));
    SSF_TRY(icp->q_d2.reserve(n));
    SSF_TRY(upload_cloud(ctx, queries, n, stride_bytes, icp->q_dev.p));
    SSF_TRY(nn_search_device(icp->map.view, icp->q_dev.p, n, max_sqdist, icp->q_idx.p, icp->q_d2.p, ctx->stream));
    SSF_CUDA(cudaMemcpyAsync(idx, icp->q_idx.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SSF_CUDA(cudaMemcpyAsync(d2, icp->q_d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    return SSF_OK;
}

extern "C" int ssf_nn_search_bench(ssf_icp *icp, const float *queries, size_t n, size_t stride_bytes, float max_sqdist,
                                   int reps, float *ms_per_pass, int32_t *idx, float *d2)
{
    SSF_ARG(icp && queries && ms_per_pass && n > 0 && reps > 0, "ssf_nn_search_bench: bad argument");
    if (!icp->has_target) {
        set_error("ssf_nn_search_bench: no target set");
        return SSF_ERR_STATE;
    }
    ssf_ctx *ctx = icp->ctx;
    SSF_TRY(use_device(ctx));
    SSF_TRY(icp->q_dev.reserve(n));
    SSF_TRY(icp->q_idx.reserve(n));
    SSF_TRY(icp->q_d2.reserve(n));
    SSF_TRY(upload_cloud(ctx, queries, n, stride_bytes, icp->q_dev.p));
    SSF_TRY(nn_search_device(icp->map.view, icp->q_dev.p, n, max_sqdist, icp->q_idx.p, icp->q_d2.p, ctx->stream));  // warm-up
    SSF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int r = 0; r < reps; ++r)
        SSF_TRY(nn_search_device(icp->map.view, icp->q_dev.p, n, max_sqdist, icp->q_idx.p, icp->q_d2.p, ctx->stream));
    SSF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (idx) SSF_CUDA(cudaMemcpyAsync(idx, icp->q_idx.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (d2) SSF_CUDA(cudaMemcpyAsync(d2, icp->q_d2.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    SSF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *ms_per_pass = ms / (float)reps;
    return SSF_OK;
}

// ---- voxel grid ---------------------------------------------------------------------------------
extern "C" int ssf_voxel_downsample(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, float leaf,
                                    float *out, size_t *n_out, int *refused)
{
    SSF_ARG(ctx && n_out, "ssf_voxel_downsample: NULL argument");
    SSF_ARG(n == 0 || (xyz && out), "ssf_voxel_downsample: NULL cloud");
    SSF_ARG(leaf > 0.f && std::isfinite(leaf), "ssf_voxel_downsample: leaf must be > 0");
    SSF_ARG(n < ((size_t)1 << 31), "ssf_voxel_downsample: more than 2^31 - 1 points");
    *n_out = 0;
    if (refused) *refused = 0;
    if (n == 0) return SSF_OK;
    SSF_TRY(use_device(ctx));
    VoxelWork w;
    SSF_TRY(w.in.reserve(n));
    SSF_TRY(w.out.reserve(n));
    SSF_TRY(upload_cloud(ctx, xyz, n, stride_bytes, w.in.p));
    uint32_t cnt = 0;
    int ref = 0;
    SSF_TRY(voxel_downsample_device(w, n, leaf, ctx->scratch, ctx->stream, &cnt, &ref));
    SSF_CUDA(cudaMemcpyAsync(out, w.out.p, (size_t)cnt * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_out = cnt;
    if (refused) *refused = ref;
    return SSF_OK;
}

extern "C" int ssf_voxel_downsample_o3d(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, double voxel_size,
                                        float *out, size_t *n_out)
{
    SSF_ARG(ctx && n_out, "ssf_voxel_downsample_o3d: NULL argument");
    SSF_ARG(n == 0 || (xyz && out), "ssf_voxel_downsample_o3d: NULL cloud");
    SSF_ARG(voxel_size > 0.0 && std::isfinite(voxel_size), "ssf_voxel_downsample_o3d: voxel_size must be > 0");
    SSF_ARG(n < ((size_t)1 << 31), "ssf_voxel_downsample_o3d: more than 2^31 - 1 points");
    *n_out = 0;
    if (n == 0) return SSF_OK;
    SSF_TRY(use_device(ctx));
    VoxelWork w;
    SSF_TRY(w.in.reserve(n));
    SSF_TRY(w.out.reserve(n));
    SSF_TRY(upload_cloud(ctx, xyz, n, stride_bytes, w.in.p));
    uint32_t cnt = 0;
    SSF_TRY(voxel_downsample_o3d_device(w, n, voxel_size, ctx->scratch, ctx->stream, &cnt));
    SSF_CUDA(cudaMemcpyAsync(out, w.out.p, (size_t)cnt * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_out = cnt;
    return SSF_OK;
}

// ---- cloud pre-processing ---------------------------------------------------------------------------
static int preproc_io(ssf_ctx *ctx, PreprocWork &w, const float *xyz, size_t n, size_t stride_bytes)
{
    SSF_ARG(n < ((size_t)1 << 31), "cloud larger than 2^31 - 1 points");
    SSF_TRY(use_device(ctx));
    SSF_TRY(w.in.reserve(n ? n : 1));
    SSF_TRY(w.out.reserve(n ? n : 1));
    return upload_cloud(ctx, xyz, n, stride_bytes, w.in.p);
}

static int preproc_out(ssf_ctx *ctx, PreprocWork &w, uint32_t cnt, float *out, size_t *n_out, int32_t *idx)
{
    if (cnt) SSF_CUDA(cudaMemcpyAsync(out, w.out.p, (size_t)cnt * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (cnt && idx) SSF_CUDA(cudaMemcpyAsync(idx, w.idx.p, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_out = cnt;
    return SSF_OK;
}

extern "C" int ssf_cloud_subsample(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, size_t point_step,
                                   float *out, size_t *n_out)
{
    SSF_ARG(ctx && n_out && (n == 0 || (xyz && out)), "ssf_cloud_subsample: NULL argument");
    *n_out = 0;
    PreprocWork w;
    SSF_TRY(preproc_io(ctx, w, xyz, n, stride_bytes));
    uint32_t cnt = 0;
    SSF_TRY(subsample_device(w, n, point_step, &cnt, ctx->stream));
    return preproc_out(ctx, w, cnt, out, n_out, nullptr);
}

extern "C" int ssf_cloud_remove_floor(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, float *out,
                                      size_t *n_out)
{
    SSF_ARG(ctx && n_out && (n == 0 || (xyz && out)), "ssf_cloud_remove_floor: NULL argument");
    *n_out = 0;
    PreprocWork w;
    SSF_TRY(preproc_io(ctx, w, xyz, n, stride_bytes));
    uint32_t cnt = 0;
    SSF_TRY(remove_floor_device(w, n, ctx->scratch, &cnt, ctx->stream));
    return preproc_out(ctx, w, cnt, out, n_out, nullptr);
}

extern "C" int ssf_cloud_crop_radius(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes,
                                     const float center[3], double radius, float *out, size_t *n_out,
                                     int32_t *indices_out)
{
    SSF_ARG(ctx && n_out && center && (n == 0 || (xyz && out)), "ssf_cloud_crop_radius: NULL argument");
    SSF_ARG(radius >= 0.0, "ssf_cloud_crop_radius: radius < 0");
    *n_out = 0;
    PreprocWork w;
    SSF_TRY(preproc_io(ctx, w, xyz, n, stride_bytes));
    uint32_t cnt = 0;
    SSF_TRY(crop_radius_device(w, n, center, radius, ctx->scratch, &cnt, ctx->stream));
    return preproc_out(ctx, w, cnt, out, n_out, indices_out);
}

// ---- brute-force alignment --------------------------------------------------------------------------
extern "C" size_t ssf_bfa_pose_count(const ssf_bfa_params *params)
{
    if (!params) return 0;
    const float id[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::vector<float> poses;
    if (bfa_poses_host(id, *params, poses) != SSF_OK) return 0;
    return poses.size() / 16;
}

extern "C" int ssf_bfa_align(ssf_icp *icp, const float *src_xyz, size_t n, size_t stride_bytes,
                             const float T_prev_colmajor[16], const ssf_bfa_params *params, float T_best_colmajor[16],
                             float *best_score, int *success, float *scores_out)
{
    SSF_ARG(icp && T_prev_colmajor && params && T_best_colmajor && best_score && success,
            "ssf_bfa_align: NULL argument");
    SSF_ARG(n >= 1 && src_xyz, "ssf_bfa_align: empty source cloud");
    if (!icp->has_target) {
        set_error("ssf_bfa_align: no target set");
        return SSF_ERR_STATE;
    }
    ssf_ctx *ctx = icp->ctx;
    SSF_TRY(use_device(ctx));
    std::vector<float> poses, scores;
    SSF_TRY(bfa_poses_host(T_prev_colmajor, *params, poses));
    BfaWork w;
    SSF_TRY(w.src.reserve(n));
    SSF_TRY(upload_cloud(ctx, src_xyz, n, stride_bytes, w.src.p));
    SSF_TRY(bfa_scores_device(icp->map.view, w, n, poses, scores, ctx->stream));
    // the decisions of cpp:107-134, taken in loop order on the bit-exact scores
    float best = 3.402823466e+38f;
    size_t best_k = 0, hit_k = 0;
    bool hit = false;
    for (size_t k = 0; k < scores.size(); ++k) {
        if (scores[k] < best) { best = scores[k]; best_k = k; }
        if (scores[k] < params->mean_error_threshold) { hit = true; hit_k = k; break; }
    }
    if (scores.empty()) {
        for (int i = 0; i < 16; ++i) T_best_colmajor[i] = (i % 5 == 0) ? 1.f : 0.f;  // cpp:68
    } else {
        memcpy(T_best_colmajor, poses.data() + 16 * (hit ? hit_k : best_k), 16 * sizeof(float));
    }
    *best_score = best;
    *success = hit ? 1 : 0;
    if (scores_out) memcpy(scores_out, scores.data(), scores.size() * sizeof(float));
    return SSF_OK;
}

// ---- batches ------------------------------------------------------------------------------------
extern "C" int ssf_batch_create(ssf_icp *icp, size_t max_scans, size_t max_total_points, ssf_batch **out)
{
    SSF_ARG(icp && out, "ssf_batch_create: NULL argument");
    SSF_ARG(max_scans >= 1 && max_scans < (1u << 24), "ssf_batch_create: max_scans out of range");
    SSF_ARG(max_total_points + max_scans * kSlotAlign < ((size_t)1 << 31), "ssf_batch_create: too many points");
    *out = nullptr;
    SSF_TRY(use_device(icp->ctx));
    ssf_batch *b = new (std::nothrow) ssf_batch;
    if (!b) return SSF_ERR_NOMEM;
    b->icp = icp;
    b->max_scans = max_scans;
    b->max_points = max_total_points;
    const size_t slots = max_total_points + max_scans * kSlotAlign;  // alignment padding
    const size_t tiles = slots / kTile + 1;
    int rc = SSF_OK;
    auto chk = [&](int r) { if (rc == SSF_OK) rc = r; };
    chk(b->buf.src.reserve(slots));
    chk(b->buf.corr.reserve(slots));
    chk(b->buf.cert.reserve(slots));
    chk(b->buf.pose_hist.reserve(max_scans * 32 * 16));
    chk(b->buf.tile_scan.reserve(tiles));
    b->buf.max_tiles = tiles;
    chk(b->buf.partials.reserve(tiles * kAccum));
    chk(b->buf.state.reserve(max_scans));
    chk(b->buf.sums.reserve(max_scans * kAccum));
    chk(b->buf.results.reserve(max_scans));
    chk(b->T_init_pinned.reserve(max_scans * 16));
    chk(b->meta_dev.reserve(max_scans * 5));
    if (rc == SSF_OK && (cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
                         cudaEventCreateWithFlags(&b->uploaded_ev, cudaEventDisableTiming) != cudaSuccess ||
                         cudaEventCreateWithFlags(&b->ran_ev, cudaEventDisableTiming) != cudaSuccess ||
                         cudaEventCreate(&b->ev0) != cudaSuccess || cudaEventCreate(&b->ev1) != cudaSuccess)) {
        set_error("ssf_batch_create: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = SSF_ERR_CUDA;
    }
    if (rc != SSF_OK) {
        ssf_batch_destroy(b);
        return rc;
    }
    *out = b;
    return SSF_OK;
}

extern "C" void ssf_batch_destroy(ssf_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->icp->ctx->device);
    cudaStreamSynchronize(b->icp->ctx->stream);
    if (b->copy_stream) cudaStreamSynchronize(b->copy_stream);
    if (b->uploaded_ev) cudaEventDestroy(b->uploaded_ev);
    if (b->ran_ev) cudaEventDestroy(b->ran_ev);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    delete b;
}

static int batch_upload_impl(ssf_batch *b, const float *xyz, const size_t *n_pts, size_t n_scans, size_t stride_bytes,
                             bool wait)
{
    SSF_ARG(b && n_pts, "ssf_batch_upload: NULL argument");
    SSF_ARG(n_scans >= 1 && n_scans <= b->max_scans, "ssf_batch_upload: n_scans exceeds the batch capacity");
    SSF_ARG(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride_bytes must be a multiple of 4 and >= 12");
    ssf_ctx *ctx = b->icp->ctx;
    SSF_TRY(use_device(ctx));
    size_t total = 0;
    for (size_t s = 0; s < n_scans; ++s) total += n_pts[s];
    SSF_ARG(total <= b->max_points, "ssf_batch_upload: points exceed the batch capacity");
    SSF_ARG(total == 0 || xyz, "ssf_batch_upload: xyz == NULL");
    cudaStream_t cs = b->copy_stream;
    // a previous alignment of THIS batch may still read its buffers
    if (b->ran) SSF_CUDA(cudaStreamWaitEvent(cs, b->ran_ev, 0));
    b->meta_host.assign(5 * n_scans, 0);
    b->n_raw.assign(n_scans, 0);
    size_t raw = 0, tile = 0;
    uint32_t max_n = 0;
    for (size_t s = 0; s < n_scans; ++s) {
        const uint32_t n = (uint32_t)n_pts[s];
        const uint32_t cap = (n + kSlotAlign - 1) / kSlotAlign * (kSlotAlign / kTile);  // search tiles, whole sort tiles
        b->meta_host[5 * s + 0] = (uint32_t)raw;
        b->meta_host[5 * s + 1] = n;
        b->meta_host[5 * s + 2] = (uint32_t)(tile * kTile);
        b->meta_host[5 * s + 3] = (uint32_t)tile;
        b->meta_host[5 * s + 4] = cap;
        b->n_raw[s] = n;
        raw += n;
        tile += cap;
        if (n > max_n) max_n = n;
    }
    b->buf.n_scans = n_scans;
    b->buf.n_tiles = tile;
    b->buf.n_slots = tile * kTile;
    b->total_points = total;
    // pinned copy of the table: a pageable source would make the "async" copy synchronise the stream
    SSF_TRY(b->meta_pinned.reserve(b->meta_host.size()));
    if (b->uploaded) SSF_CUDA(cudaEventSynchronize(b->uploaded_ev));  // the previous upload may still read it
    memcpy(b->meta_pinned.p, b->meta_host.data(), b->meta_host.size() * sizeof(uint32_t));
    SSF_CUDA(cudaMemcpyAsync(b->meta_dev.p, b->meta_pinned.p, b->meta_host.size() * sizeof(uint32_t),
                             cudaMemcpyHostToDevice, cs));
    if (total > 0) {
        const size_t bytes = (total - 1) * stride_bytes + 12;
        SSF_TRY(b->stage.reserve(bytes));
        static const bool dbg = getenv("SSF_DEBUG_TIMING") != nullptr;
        if (dbg && !b->dbg[0]) for (auto &e : b->dbg) cudaEventCreate(&e);
        if (dbg) cudaEventRecord(b->dbg[0], cs);
        SSF_CUDA(cudaMemcpyAsync(b->stage.p, xyz, bytes, cudaMemcpyHostToDevice, cs));
        if (dbg) cudaEventRecord(b->dbg[1], cs);
        unsigned bx = (max_n + 255) / 256;
        if (bx > 64) bx = 64;
        if (bx == 0) bx = 1;
        float4 *dst = b->buf.src.p;
        if (b->icp->prm.source_voxel_leaf > 0.f) {
            SSF_TRY(b->buf.raw.reserve(b->buf.src.cap));
            dst = b->buf.raw.p;
        }
        pack_scans_kernel<<<dim3(bx, (unsigned)n_scans), 256, 0, cs>>>(b->stage.p, stride_bytes, b->meta_dev.p, dst);
        SSF_LAUNCHED();
    }
    b->uploaded_raw = b->icp->prm.source_voxel_leaf > 0.f;
    layout_kernel<<<(unsigned)n_scans, 128, 0, cs>>>(b->buf.state.p, b->meta_dev.p, (uint32_t)n_scans, b->buf.tile_scan.p);
    SSF_LAUNCHED();
    if (b->dbg[2]) cudaEventRecord(b->dbg[2], cs);
    SSF_CUDA(cudaEventRecord(b->uploaded_ev, cs));
    if (wait) SSF_CUDA(cudaStreamSynchronize(cs));  // the caller's buffer is free again
    b->uploaded = true;
    b->initial_set = false;
    b->ran = false;
    return SSF_OK;
}

extern "C" int ssf_batch_upload(ssf_batch *b, const float *xyz, const size_t *n_pts, size_t n_scans,
                                size_t stride_bytes)
{
    return batch_upload_impl(b, xyz, n_pts, n_scans, stride_bytes, true);
}

extern "C" int ssf_batch_upload_async(ssf_batch *b, const float *xyz, const size_t *n_pts, size_t n_scans,
                                      size_t stride_bytes)
{
    return batch_upload_impl(b, xyz, n_pts, n_scans, stride_bytes, false);
}

extern "C" int ssf_batch_set_initial(ssf_batch *b, const float *T_colmajor)
{
    SSF_ARG(b && T_colmajor, "ssf_batch_set_initial: NULL argument");
    if (!b->uploaded) {
        set_error("ssf_batch_set_initial: upload scans first");
        return SSF_ERR_STATE;
    }
    ssf_ctx *ctx = b->icp->ctx;
    SSF_TRY(use_device(ctx));
    // 64 B per scan.  NOT a copy-engine transfer: a small H2D copy queued here would sit behind the
    // other batch's large scan upload on the copy engine and hold this batch's alignment back for
    // the whole of it.  The transforms go to pinned host memory instead, which init_states_kernel
    // reads directly (zero-copy, a few KB).
    if (b->ever_ran) SSF_CUDA(cudaEventSynchronize(b->ran_ev));  // a previous alignment may still read them
    memcpy(b->T_init_pinned.p, T_colmajor, b->buf.n_scans * 16 * sizeof(float));
    b->initial_set = true;
    return SSF_OK;
}

extern "C" int ssf_batch_run(ssf_batch *b)
{
    SSF_ARG(b, "ssf_batch_run: b == NULL");
    ssf_icp *icp = b->icp;
    if (!icp->has_target) {
        set_error("ssf_batch_run: no target set");
        return SSF_ERR_STATE;
    }
    if (!b->uploaded || !b->initial_set) {
        set_error("ssf_batch_run: upload scans and set initial transforms first");
        return SSF_ERR_STATE;
    }
    ssf_ctx *ctx = icp->ctx;
    SSF_TRY(use_device(ctx));
    const ssf_icp_params &p = icp->prm;
    BatchBuffers &buf = b->buf;
    // per-pass error trace: always in REFERENCE mode (ssf_icp_get_trace), in the other modes only for
    // the debug print of a single alignment
    const int trace_len = p.mode == SSF_MODE_REFERENCE ? p.num_iterations
                                                       : (p.debug && buf.n_scans == 1 ? p.num_iterations + 1 : 0);
    buf.trace_len = trace_len;
    SSF_TRY(buf.trace_err.reserve(buf.n_scans * (size_t)(trace_len ? trace_len : 1)));
    SSF_TRY(buf.trace_search.reserve(buf.n_scans * (size_t)(trace_len ? trace_len : 1)));
    if (p.mode == SSF_MODE_REFERENCE) {
        SSF_TRY(buf.P.reserve(buf.src.cap));
        SSF_TRY(buf.Q.reserve(buf.src.cap));
    }
    if ((p.source_voxel_leaf > 0.f) != b->uploaded_raw) {
        set_error("ssf_batch_run: source_voxel_leaf changed after the scans were uploaded; upload again");
        return SSF_ERR_STATE;
    }
    SSF_CUDA(cudaStreamWaitEvent(ctx->stream, b->uploaded_ev, 0));
    SSF_CUDA(cudaEventRecord(b->ev0, ctx->stream));
    if (p.source_voxel_leaf > 0.f && b->total_points > 0)
        SSF_TRY(voxel_downsample_batch(buf, b->meta_dev.p, p.source_voxel_leaf, ctx->scratch, ctx->stream));
    IcpConfig cfg{p.max_correspondence_dist, p.num_iterations, p.acceptable_mean_error, p.transformation_epsilon,
                  p.mode, p.reduce};
    if (icp->map.sharded) {
        cfg.allreduce = icp->allreduce;
        cfg.allreduce_user = icp->allreduce_user;
        if (icp->xch.ready) {
            if (buf.n_scans > icp->xch.max_scans) {
                set_error("ssf_batch_run: %zu scans exceed the exchange capacity (%zu)", buf.n_scans, icp->xch.max_scans);
                return SSF_ERR_INVALID;
            }
            cfg.xch.peers = icp->xch.peers_dev.p;
            cfg.xch.rank = icp->xch.rank;
            cfg.xch.world = icp->xch.world;
            cfg.xch.max_scans = (uint32_t)icp->xch.max_scans;
            cfg.xch.epoch = icp->xch.epoch_dev.p;
        } else if (icp->nccl) {
            cfg.nccl_allreduce = nccl_link_allreduce;
            cfg.nccl_user = icp->nccl;
        }
    }
    SSF_TRY(ensure_reach_mask(icp->map, cfg.max_corr, ctx->stream));
    SSF_TRY(run_batch(icp->map.view, cfg, buf, b->T_init_pinned.p, ctx->stream, &ctx->timer));
    SSF_CUDA(cudaEventRecord(b->ev1, ctx->stream));
    SSF_CUDA(cudaEventRecord(b->ran_ev, ctx->stream));
    b->ran = true;
    b->ever_ran = true;
    return SSF_OK;
}

extern "C" int ssf_batch_results(ssf_batch *b, ssf_icp_result *out, size_t n_scans)
{
    SSF_ARG(b && out, "ssf_batch_results: NULL argument");
    if (!b->ran) {
        set_error("ssf_batch_results: the batch has not run");
        return SSF_ERR_STATE;
    }
    SSF_ARG(n_scans <= b->buf.n_scans, "ssf_batch_results: n_scans larger than the batch");
    ssf_ctx *ctx = b->icp->ctx;
    SSF_TRY(use_device(ctx));
    // D2H on the batch's copy stream: waits for THIS batch's alignment only, so another batch may
    // keep the compute stream busy meanwhile
    SSF_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ran_ev, 0));
    SSF_TRY(b->res_pinned.reserve(n_scans ? n_scans : 1));
    SSF_CUDA(cudaMemcpyAsync(b->res_pinned.p, b->buf.results.p, n_scans * sizeof(ssf_icp_result), cudaMemcpyDeviceToHost,
                             b->copy_stream));
    SSF_CUDA(cudaStreamSynchronize(b->copy_stream));
    memcpy(out, b->res_pinned.p, n_scans * sizeof(ssf_icp_result));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess) {
        cudaGetLastError();
        ms = 0.f;
    }
    b->last_ms = ms;
    if (b->dbg[2]) {  // timeline relative to the start of this batch's H2D copy
        float t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        cudaEventElapsedTime(&t1, b->dbg[0], b->dbg[1]);
        cudaEventElapsedTime(&t2, b->dbg[0], b->dbg[2]);
        cudaEventElapsedTime(&t3, b->dbg[0], b->ev0);
        cudaEventElapsedTime(&t4, b->dbg[0], b->ev1);
        fprintf(stderr, "[ssf timing] batch %p: H2D end %.3f  pack end %.3f  run begin %.3f  run end %.3f ms\n", (void *)b, t1,
                t2, t3, t4);
    }
    for (size_t s = 0; s < n_scans; ++s) out[s].device_ms = ms;
    for (size_t s = 0; s < n_scans; ++s)
        if (out[s].aborted == 2) {
            set_error("scan %zu: the per-scan sums of a peer rank did not arrive (timeout, or the ranks ran different "
                      "batch shapes / iteration counts / modes)", s);
            return SSF_ERR_COMM;
        }
    return SSF_OK;
}

extern "C" int ssf_batch_search_stats(ssf_batch *b, uint64_t *answered, uint64_t *walked, size_t cap, size_t *n_launches)
{
    SSF_ARG(b && n_launches, "ssf_batch_search_stats: NULL argument");
    if (!b->ran) {
        set_error("ssf_batch_search_stats: the batch has not run");
        return SSF_ERR_STATE;
    }
    ssf_ctx *ctx = b->icp->ctx;
    SSF_TRY(use_device(ctx));
    const size_t n = (size_t)b->buf.search_stats_len;
    *n_launches = n;
    const size_t m = n < cap ? n : cap;
    if (m == 0) return SSF_OK;
    std::vector<unsigned long long> h(2 * m);
    SSF_CUDA(cudaStreamSynchronize(ctx->stream));
    SSF_CUDA(cudaMemcpy(h.data(), b->buf.search_stats.p, 2 * m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < m; ++i) {
        if (answered) answered[i] = h[2 * i];
        if (walked) walked[i] = h[2 * i + 1];
    }
    return SSF_OK;
}

extern "C" int ssf_icp_align_batch(ssf_icp *icp, const float *xyz, const size_t *n_pts, size_t n_scans,
                                   size_t stride_bytes, const float *T_colmajor, ssf_icp_result *out)
{
    SSF_ARG(icp && n_pts && T_colmajor && out, "ssf_icp_align_batch: NULL argument");
    SSF_ARG(n_scans >= 1, "ssf_icp_align_batch: n_scans == 0");
    size_t total = 0;
    for (size_t s = 0; s < n_scans; ++s) total += n_pts[s];
    ssf_batch *b = nullptr;
    SSF_TRY(ssf_batch_create(icp, n_scans, total + 1, &b));
    int rc = ssf_batch_upload(b, xyz, n_pts, n_scans, stride_bytes);
    if (rc == SSF_OK) rc = ssf_batch_set_initial(b, T_colmajor);
    if (rc == SSF_OK) rc = ssf_batch_run(b);
    if (rc == SSF_OK) rc = ssf_batch_results(b, out, n_scans);
    ssf_batch_destroy(b);
    return rc;
}
