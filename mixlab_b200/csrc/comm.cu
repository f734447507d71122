// comm.cu -- optional shared-source mode: one ingest GPU broadcasts a source line or frame to the
// sessions on the other GPUs over NVLink (NCCL).  NEW: the reference has no counterpart (one receiver per
// mountpoint, src/source.rs:93-95); sessions otherwise share nothing and the tick path uses no collective.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library must keep loading on boxes without NCCL
// and must not drag a second NCCL into a process that already has one (torch bundles its own).
#include <dlfcn.h>
#include <string.h>

#if __has_include(<nccl.h>)
#include <nccl.h>
#define MXL_HAVE_NCCL_HEADER 1
#else
#define MXL_HAVE_NCCL_HEADER 0
#endif

#include "common.h"

namespace mxl {
namespace {

#if MXL_HAVE_NCCL_HEADER
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl& nccl()
{
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* override_path = getenv("MXL_NCCL_LIB");
    for (const char* name : {override_path, "libnccl.so.2", "libnccl.so"}) {
        if (!name) continue;
        n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (n.handle) break;
    }
    if (!n.handle) return n;
    n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.handle, "ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.handle, "ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.handle, "ncclCommDestroy");
    n.Broadcast = (decltype(n.Broadcast))dlsym(n.handle, "ncclBroadcast");
    n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.handle, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.Broadcast && n.GetErrorString;
    return n;
}

#define MXL_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) MXL_FAIL(MXL_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(_r)); \
    } while (0)
#endif

int need_nccl()
{
#if MXL_HAVE_NCCL_HEADER
    if (!nccl().ok) MXL_FAIL(MXL_ERR_INVALID, "NCCL is not available (libnccl.so.2 not found; set MXL_NCCL_LIB)");
    return MXL_OK;
#else
    MXL_FAIL(MXL_ERR_INVALID, "built without nccl.h: the shared-source broadcast is unavailable");
#endif
}

int broadcast_bytes(mxl_ctx* ctx, void* dev, size_t bytes, int root)
{
#if MXL_HAVE_NCCL_HEADER
    if (!ctx || !ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    if (!ctx->comm) MXL_FAIL(MXL_ERR_INVALID, "mxl_ctx_comm_init has not been called on this context");
    if (root < 0 || root >= ctx->comm_world) MXL_FAIL(MXL_ERR_INVALID, "root %d of %d ranks", root, ctx->comm_world);
    if (bytes == 0) return MXL_OK;
    MXL_TRY(ctx->activate());
    MXL_TRY(ctx->compute_begin());
    MXL_NCCL(nccl().Broadcast(dev, dev, bytes, ncclUint8, root, (ncclComm_t)ctx->comm, ctx->stream));
    return ctx->compute_end();
#else
    (void)ctx; (void)dev; (void)bytes; (void)root;
    return need_nccl();
#endif
}

}  // namespace
}  // namespace mxl

using namespace mxl;

extern "C" {

int mxl_comm_unique_id(uint8_t id_out[MXL_COMM_ID_BYTES])
{
    if (!id_out) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    MXL_TRY(need_nccl());
#if MXL_HAVE_NCCL_HEADER
    static_assert(sizeof(ncclUniqueId) <= MXL_COMM_ID_BYTES, "ncclUniqueId larger than MXL_COMM_ID_BYTES");
    ncclUniqueId id;
    MXL_NCCL(nccl().GetUniqueId(&id));
    memset(id_out, 0, MXL_COMM_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
#endif
    return MXL_OK;
}

int mxl_ctx_comm_init(mxl_ctx* ctx, const uint8_t id[MXL_COMM_ID_BYTES], int rank, int world)
{
    if (!ctx || !id) MXL_FAIL(MXL_ERR_INVALID, "NULL argument");
    if (!ctx->has_device()) MXL_FAIL(MXL_ERR_NO_DEVICE, "no CUDA device bound to this context");
    if (world < 1 || rank < 0 || rank >= world) MXL_FAIL(MXL_ERR_INVALID, "rank %d of %d", rank, world);
    if (ctx->comm) MXL_FAIL(MXL_ERR_INVALID, "context already has a communicator");
    MXL_TRY(need_nccl());
#if MXL_HAVE_NCCL_HEADER
    MXL_TRY(ctx->activate());
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm = nullptr;
    MXL_NCCL(nccl().CommInitRank(&comm, world, uid, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
#endif
    return MXL_OK;
}

int mxl_ctx_comm_destroy(mxl_ctx* ctx)
{
    if (!ctx || !ctx->comm) return MXL_OK;
#if MXL_HAVE_NCCL_HEADER
    ctx->activate();
    cudaStreamSynchronize(ctx->stream);
    nccl().CommDestroy((ncclComm_t)ctx->comm);
#endif
    ctx->comm = nullptr;
    ctx->comm_world = 0;
    return MXL_OK;
}

int mxl_line_broadcast(mxl_line* line, int root)
{
    if (!line) MXL_FAIL(MXL_ERR_INVALID, "NULL line");
    if (line->type == MXL_LINE_VIDEO) MXL_FAIL(MXL_ERR_LINE_TYPE, "mxl_line_broadcast: audio lines only; broadcast the frames of a video line");
    return broadcast_bytes(line->ctx, line->dev, (size_t)line->len() * sizeof(float), root);
}

int mxl_frame_broadcast(mxl_frame* frame, int root)
{
    if (!frame) MXL_FAIL(MXL_ERR_INVALID, "NULL frame");
    return broadcast_bytes(frame->ctx, frame->dev, (size_t)frame->layout.size, root);
}

}  // extern "C"
