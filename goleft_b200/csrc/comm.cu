// comm.cu — the one collective on this path: an NCCL all-gather over NVLink 5 / NVSwitch that
// assembles the depthwed n-sites x n-samples matrix (and the indexcov cohort) from per-GPU shards.
// One process per GPU; the caller distributes the unique id (e.g. through torch.distributed or MPI).
//
// NCCL is bound at run time with dlopen("libnccl.so.2") so the library has no link-time NCCL
// dependency and shares whichever NCCL the host process already loaded (torch bundles its own).
#include "gl_common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;     // ncclSuccess == 0
enum { kNcclInt8 = 0 };       // ncclInt8 / ncclChar in nccl.h's ncclDataType_t

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int nccl_load(gl_ctx* ctx) {
    if (g_nccl.handle) return GL_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return gl_fail(ctx, GL_ENCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString)
        return gl_fail(ctx, GL_ENCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return GL_OK;
}

#define GL_NCCL(ctx, expr)                                                                          \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != 0) return gl_fail((ctx), GL_ENCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

}  // namespace

extern "C" {

int gl_comm_unique_id(uint8_t id128[128]) {
    if (!id128) return gl_fail(nullptr, GL_EINVAL, "null id");
    GL_CHECK(nccl_load(nullptr));
    ncclUniqueId id;
    GL_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return GL_OK;
}

int gl_comm_init(gl_ctx* ctx, const uint8_t id128[128], int rank, int world) {
    GL_CHECK(gl_use(ctx));
    if (!id128 || world < 1 || rank < 0 || rank >= world) return gl_fail(ctx, GL_EINVAL, "gl_comm_init: bad argument");
    if (ctx->nccl) return gl_fail(ctx, GL_ESTATE, "gl_comm_init: communicator already initialised");
    GL_CHECK(nccl_load(ctx));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    GL_NCCL(ctx, g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl = comm;
    ctx->rank = rank;
    ctx->world = world;
    return GL_OK;
}

int gl_comm_destroy(gl_ctx* ctx) {
    if (!ctx || !ctx->nccl) return GL_OK;
    cudaSetDevice(ctx->device);
    g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl));
    ctx->nccl = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return GL_OK;
}

int gl_allgather_device(gl_ctx* ctx, const void* d_send, void* d_recv, int64_t bytes) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->nccl) return gl_fail(ctx, GL_ESTATE, "gl_allgather_device: call gl_comm_init first");
    if (bytes < 0 || !d_send || !d_recv) return gl_fail(ctx, GL_EINVAL, "gl_allgather_device: bad argument");
    GL_NCCL(ctx, g_nccl.AllGather(d_send, d_recv, (size_t)bytes, kNcclInt8, static_cast<ncclComm_t>(ctx->nccl), ctx->stream));
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

// The same, ASYNCHRONOUS on the ctx's communication stream: the gather starts when everything queued on the compute stream
// so far has finished (so the block an earlier kernel wrote is complete) and runs beside later kernels — the depthwed
// matrix is gathered chunk by chunk while the next chunk is still being aggregated.  gl_comm_wait() joins.
int gl_allgather_device_async(gl_ctx* ctx, const void* d_send, void* d_recv, int64_t bytes) {
    GL_CHECK(gl_use(ctx));
    if (!ctx->nccl) return gl_fail(ctx, GL_ESTATE, "gl_allgather_device_async: call gl_comm_init first");
    if (bytes < 0 || !d_send || !d_recv) return gl_fail(ctx, GL_EINVAL, "gl_allgather_device_async: bad argument");
    if (!ctx->comm_stream) {
        GL_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        GL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    }
    GL_CUDA(ctx, cudaEventRecord(ctx->ev_comm, ctx->stream));
    GL_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm, 0));
    GL_NCCL(ctx, g_nccl.AllGather(d_send, d_recv, (size_t)bytes, kNcclInt8, static_cast<ncclComm_t>(ctx->nccl), ctx->comm_stream));
    return GL_OK;
}

// ---- peer memory between the per-GPU processes of one box (NVLink / NVSwitch): a device allocation of one rank mapped
// into another rank's address space, so that a kernel there can store into it directly (gl_depthwed_aggregate_i32_p2p).
int gl_ipc_export(gl_ctx* ctx, void* d_ptr, uint8_t handle64[64]) {
    GL_CHECK(gl_use(ctx));
    if (!d_ptr || !handle64) return gl_fail(ctx, GL_EINVAL, "gl_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    GL_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return GL_OK;
}

int gl_ipc_open(gl_ctx* ctx, const uint8_t handle64[64], void** d_peer_ptr) {
    GL_CHECK(gl_use(ctx));
    if (!handle64 || !d_peer_ptr) return gl_fail(ctx, GL_EINVAL, "gl_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    *d_peer_ptr = nullptr;
    GL_CUDA(ctx, cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GL_OK;
}

int gl_ipc_close(gl_ctx* ctx, void* d_peer_ptr) {
    GL_CHECK(gl_use(ctx));
    if (d_peer_ptr) GL_CUDA(ctx, cudaIpcCloseMemHandle(d_peer_ptr));
    return GL_OK;
}

int gl_comm_wait(gl_ctx* ctx) {
    GL_CHECK(gl_use(ctx));
    if (ctx->comm_stream) GL_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    GL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

}  // extern "C"
