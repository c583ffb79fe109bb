// demcz_comm.cu — multi-GPU exchange for the sharded sampler: one context (process) per GPU,
// chains partitioned contiguously, archive replicated.  Every K generations each rank has written
// its N_local new rows into the next archive segment; an in-place NCCL all-gather over
// NVLink/NVSwitch completes the segment on every rank (row order = global chain order = the
// single-GPU order).  Mirrors the barrier + shared-Z semantics of the reference's sync mode
// (src/demcz.jl:129-137) without its data race (src/demcz.jl:89-90 under :137).
//
// NCCL is resolved with dlopen at demcz_comm_init time: the library has no link-time NCCL
// dependency and loads on machines without it (single-GPU use).
#include "demcz_internal.h"
#include <dlfcn.h>
#include <cstring>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6,
       ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};

NcclApi& api()
{
    static NcclApi a;
    if (a.handle || !a.error.empty()) return a;
    const char* names[] = {getenv("DEMCZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) { a.error = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return a; }
#define SYM(field, name)                                                                 \
    *(void**)(&a.field) = dlsym(a.handle, name);                                         \
    if (!a.field) { a.error = std::string("missing NCCL symbol ") + name; return a; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return a;
}

int nccl_fail(demcz_ctx* c, const char* what, int rc)
{
    return demcz_fail(c, DEMCZ_ERR_NCCL, "%s: %s", what, api().GetErrorString ? api().GetErrorString(rc) : "?");
}

}  // namespace

extern "C" int demcz_comm_unique_id(uint8_t id[128])
{
    NcclApi& a = api();
    if (!a.error.empty()) return demcz_fail(nullptr, DEMCZ_ERR_NCCL, "%s", a.error.c_str());
    ncclUniqueId u;
    int rc = a.GetUniqueId(&u);
    if (rc != ncclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
    memcpy(id, u.internal, 128);
    return DEMCZ_OK;
}

extern "C" int demcz_comm_init(demcz_ctx* c, const uint8_t id[128])
{
    if (!c) return DEMCZ_ERR_INVALID;
    if (c->sticky) return c->sticky;
    if (c->o.world == 1) return DEMCZ_OK;
    NcclApi& a = api();
    if (!a.error.empty()) return demcz_fail(c, DEMCZ_ERR_NCCL, "%s", a.error.c_str());
    if (cudaSetDevice(c->device) != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "cudaSetDevice");
    // contiguous equal shards are required by the in-place all-gather
    if (c->o.N_global != (int64_t)c->o.N * c->o.world || c->o.chain_offset != (int64_t)c->o.N * c->o.rank)
        return demcz_fail(c, DEMCZ_ERR_INVALID, "comm_init: need N_global = N*world and chain_offset = N*rank");
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclComm_t comm = nullptr;
    int rc = a.CommInitRank(&comm, c->o.world, u, c->o.rank);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclCommInitRank", rc);
    c->comm = comm;
    return DEMCZ_OK;
}

void demcz_comm_destroy_internal(demcz_ctx* c)
{
    if (c->comm) { api().CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
}

int demcz_comm_allgather_rows(demcz_ctx* c, int64_t phys_row0)
{
    if (c->o.world == 1) return DEMCZ_OK;
    if (phys_row0 + c->o.N_global > c->zcap)
        return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "multi-GPU append segment wraps the archive ring; raise archive_capacity");
    float* seg = c->dZ + (size_t)phys_row0 * c->ZS;
    const size_t count = (size_t)c->o.N * c->ZS;                 // floats per rank
    int rc = api().AllGather(seg + (size_t)c->o.rank * count, seg, count, ncclFloat32, (ncclComm_t)c->comm, c->stream);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllGather", rc);
    return DEMCZ_OK;
}

int demcz_comm_allreduce_f64(demcz_ctx* c, double* dev, int n)
{
    if (c->o.world == 1) return DEMCZ_OK;
    int rc = api().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)c->comm, c->stream);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllReduce", rc);
    return DEMCZ_OK;
}

int demcz_comm_allreduce_u64(demcz_ctx* c, unsigned long long* dev, int n)
{
    if (c->o.world == 1) return DEMCZ_OK;
    int rc = api().AllReduce(dev, dev, (size_t)n, ncclUint64, ncclSum, (ncclComm_t)c->comm, c->stream);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllReduce", rc);
    return DEMCZ_OK;
}
