// nccl_link.cu -- the per-iteration all-reduce of the map-sharded path through NCCL, behind the C ABI.
//
// North-star: "each GPU reducing its partial J^T J / J^T r and one 27-float NCCL allreduce over NVLink
// per iteration" -- here n_scans x 32 doubles, enqueued with ncclAllReduce on the library's own stream
// (so it is part of the captured CUDA graph of the loop) and followed by the identical solve on every
// rank.  C and C++ callers get the collective without Python: they only move the 128-byte unique id
// from rank 0 to the other ranks (any transport).
//
// libnccl is bound at run time (dlopen), not at link time: libssf_gpu.so stays loadable on a box
// without NCCL (single-GPU use), and inside a process that already carries an NCCL (PyTorch's bundled
// one) the same library instance is reused instead of a second copy being loaded.
#include <dlfcn.h>

#include <mutex>

#include "nccl_link.cuh"

namespace ssf {

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess == 0
constexpr int kNcclFloat64 = 8, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

struct Api {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
Api g_api;
std::once_flag g_once;

void load_api()
{
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);  // an NCCL already in the process (PyTorch's)
        if (g_api.lib) break;
    }
    for (const char *n : names) {
        if (g_api.lib) break;
        g_api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    }
    if (!g_api.lib) return;
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))dlsym(g_api.lib, "ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))dlsym(g_api.lib, "ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(g_api.lib, "ncclCommDestroy");
    g_api.AllReduce = (decltype(g_api.AllReduce))dlsym(g_api.lib, "ncclAllReduce");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(g_api.lib, "ncclGetErrorString");
    g_api.ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllReduce;
}

int need_api()
{
    std::call_once(g_once, load_api);
    if (!g_api.ok) {
        set_error("libnccl.so.2 could not be loaded (%s)", g_api.lib ? "symbols missing" : dlerror());
        return SSF_ERR_COMM;
    }
    return SSF_OK;
}

int fail(const char *what, ncclResult_t r)
{
    set_error("%s failed: %s", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "NCCL error");
    return SSF_ERR_COMM;
}
}  // namespace

struct NcclLink {
    ncclComm_t comm = nullptr;
};

int nccl_link_unique_id(unsigned char id_out[128])
{
    SSF_TRY(need_api());
    ncclUniqueId id;
    const ncclResult_t r = g_api.GetUniqueId(&id);
    if (r != 0) return fail("ncclGetUniqueId", r);
    memcpy(id_out, id.internal, 128);
    return SSF_OK;
}

int nccl_link_create(const unsigned char id_in[128], int rank, int world, NcclLink **out)
{
    SSF_TRY(need_api());
    ncclUniqueId id;
    memcpy(id.internal, id_in, 128);
    NcclLink *l = new NcclLink;
    const ncclResult_t r = g_api.CommInitRank(&l->comm, world, id, rank);
    if (r != 0) {
        delete l;
        return fail("ncclCommInitRank", r);
    }
    *out = l;
    return SSF_OK;
}

void nccl_link_destroy(NcclLink *l)
{
    if (!l) return;
    if (l->comm && g_api.ok) g_api.CommDestroy(l->comm);
    delete l;
}

int nccl_link_allreduce(void *link, double *buf, size_t count, cudaStream_t stream)
{
    NcclLink *l = static_cast<NcclLink *>(link);
    const ncclResult_t r = g_api.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, l->comm, stream);
    if (r != 0) return fail("ncclAllReduce", r);
    return SSF_OK;
}

}  // namespace ssf
