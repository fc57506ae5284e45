// NCCL is resolved at run time (dlopen) so that the library links against nothing but the CUDA
// runtime and — when the host process already carries an NCCL (e.g. the one bundled with torch) —
// shares that copy instead of loading a second one.
//
// The hot path has exactly one exchange step: all-reduce(min) of the per-replicate pooled minimum
// p-value maxT[m] across site shards (the reference computes that minimum over ALL informative
// sites of the alignment, HC.java:3239-3243).
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "pb_internal.h"

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclMin = 3 };

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api(std::string& err) {
    // contexts of one process may be driven from different host threads (pb_group does): resolve the table exactly once
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (api.h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
        }
    });
    if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
        err = "NCCL not available: dlopen(libnccl.so.2) failed or symbols missing";
        return nullptr;
    }
    return &api;
}

int pbn_unique_id(void* id128, std::string& err) {
    NcclApi* a = nccl_api(err);
    if (!a) return PB_ERR_NCCL;
    ncclUniqueId id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != 0) { err = std::string("ncclGetUniqueId: ") + (a->GetErrorString ? a->GetErrorString(r) : "error"); return PB_ERR_NCCL; }
    memcpy(id128, &id, 128);
    return PB_OK;
}

int pbn_init(pb_ctx* ctx, const void* id128, int rank, int world) {
    NcclApi* a = nccl_api(ctx->err);
    if (!a) return PB_ERR_NCCL;
    if (ctx->nccl_comm) { a->CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ctx->rank = rank; ctx->world = world;
    if (world == 1) return PB_OK;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = a->CommInitRank(&comm, world, id, rank);
    if (r != 0) { ctx->err = std::string("ncclCommInitRank: ") + (a->GetErrorString ? a->GetErrorString(r) : "error"); return PB_ERR_NCCL; }
    ctx->nccl_comm = comm;
    return PB_OK;
}

int pbn_allreduce_min_f64(pb_ctx* ctx, double* buf, int64_t n) {
    if (ctx->world <= 1) return PB_OK;
    NcclApi* a = nccl_api(ctx->err);
    if (!a || !ctx->nccl_comm) { ctx->err = "all-reduce requested but pb_comm_init was not called"; return PB_ERR_NCCL; }
    // p-values are non-negative doubles: their bit patterns order like unsigned integers (the kernels already atomicMin them that
    // way), and an integer min is eligible for NCCL's in-switch (NVLS) reduction, which a double min is not
    ncclResult_t r = a->AllReduce(buf, buf, (size_t)n, ncclUint64, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) { ctx->err = std::string("ncclAllReduce: ") + (a->GetErrorString ? a->GetErrorString(r) : "error"); return PB_ERR_NCCL; }
    return PB_OK;
}

void pbn_destroy(pb_ctx* ctx) {
    if (!ctx->nccl_comm) return;
    std::string err;
    NcclApi* a = nccl_api(err);
    if (a) a->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}
