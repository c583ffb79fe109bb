// demcz_comm.cu — multi-GPU exchange for the sharded sampler: one context (process) per GPU,
// chains partitioned contiguously, archive replicated.  Every K generations each rank has written
// its N_local new rows into the next archive segment; an in-place NCCL all-gather over
// NVLink/NVSwitch completes the segment on every rank (row order = global chain order = the
// single-GPU order).  Mirrors the barrier + shared-Z semantics of the reference's sync mode
// (src/demcz.jl:129-137) without its data race (src/demcz.jl:89-90 under :137).
//
// NCCL is resolved with dlopen at demcz_comm_init time: the library has no link-time NCCL
// dependency and loads on machines without it (single-GPU use).
//
// Preferred transport once the peers' archives are opened (demcz_comm_open_peers): our own push kernel.
// Every rank stores its N_local new rows (the 16-byte pieces that hold data, not the slot padding) straight
// into the same slots of every peer's archive replica over NVLink/NVSwitch — the minimum bytes that must cross
// the links (N_local * 16*ceil(d/4) bytes per peer), no staging buffer, no unpack pass — and a one-word NCCL
// flag word per rank in every peer's allocation publishes the segment.  In overlap mode the push runs on a
// high-priority side stream under the next K-block's step kernel.
#include "demcz_internal.h"
#include "demcz_sync.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <unistd.h>
#include <cstring>
#include <algorithm>

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
    int (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;
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
    SYM(CommSplit, "ncclCommSplit")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return a;
}

// stream memory operations, resolved from the driver at run time (no link-time libcuda dependency)
typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct MemOps { StreamValue32Fn wait = nullptr, write = nullptr; bool tried = false; };
MemOps& memops()
{
    static MemOps m;
    if (m.tried) return m;
    m.tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        m.wait = reinterpret_cast<StreamValue32Fn>(p);
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        m.write = reinterpret_cast<StreamValue32Fn>(p);
    cudaGetLastError();
    return m;
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
    // side stream + its own communicator for the overlapped exchange (NCCL operations of one communicator must
    // not be issued from two streams)
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi) != cudaSuccess)
        return demcz_fail(c, DEMCZ_ERR_CUDA, "comm_init: cudaStreamCreateWithPriority");
    ncclComm_t comm2 = nullptr;
    rc = a.CommSplit(comm, 0, c->o.rank, &comm2, nullptr);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclCommSplit", rc);
    c->comm2 = comm2;
    for (int i = 0; i < 2; ++i) {
        if (cudaEventCreateWithFlags(&c->ev_app[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_vis[i], cudaEventDisableTiming) != cudaSuccess)
            return demcz_fail(c, DEMCZ_ERR_CUDA, "comm_init: cudaEventCreate");
    }
    if (cudaMalloc((void**)&c->d_barrier, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(c->d_barrier, 0, 8 * sizeof(unsigned long long)) != cudaSuccess)
        return demcz_fail(c, DEMCZ_ERR_CUDA, "comm_init: barrier word");
    // NCCL connects lazily at the first collective of a communicator (hundreds of ms): pay for it here, not inside
    // the first Rhat / adaptγ reduction of a run
    rc = a.AllReduce(c->d_barrier + 2, c->d_barrier + 2, 1, ncclUint64, ncclSum, comm, c->stream);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllReduce (warm-up)", rc);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "comm_init: warm-up sync");
    rc = a.AllReduce(c->d_barrier + 3, c->d_barrier + 3, 1, ncclUint64, ncclSum, comm2, c->comm_stream);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllReduce (warm-up, side)", rc);
    if (cudaStreamSynchronize(c->comm_stream) != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "comm_init: warm-up sync");
    return DEMCZ_OK;
}

void demcz_comm_destroy_internal(demcz_ctx* c)
{
    for (int r = 0; r < 8; ++r)
        if (c->peerZ[r] && c->peer_ipc[r]) cudaIpcCloseMemHandle(c->peerZ[r]);
    if (c->comm2) { api().CommDestroy((ncclComm_t)c->comm2); c->comm2 = nullptr; }
    if (c->comm) { api().CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
    for (int i = 0; i < 2; ++i) {
        if (c->ev_app[i]) cudaEventDestroy(c->ev_app[i]);
        if (c->ev_vis[i]) cudaEventDestroy(c->ev_vis[i]);
    }
    for (int i = 1; i < 8; ++i) if (c->xs[i]) cudaStreamDestroy(c->xs[i]);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->d_barrier) cudaFree(c->d_barrier);
}

// ------------------------------------------------------------------------------------------------
// peer-to-peer archive replicas
// ------------------------------------------------------------------------------------------------
namespace {
struct PeerInfo {                       // 128 bytes on the wire
    int64_t pid;
    int32_t device, ZS;
    uint64_t ptr;
    int64_t zcap;
    cudaIpcMemHandle_t handle;          // 64 bytes
    uint8_t pad[128 - 8 - 8 - 8 - 8 - 64];
};
static_assert(sizeof(PeerInfo) == 128, "PeerInfo must be 128 bytes");

struct PeerPtrs { float* p[8]; };

// Rows [row0, row0 + nrows) of the local archive -> the same slots of every peer replica; only the C data pieces
// (16 bytes each) of a slot cross the links.  A small persistent grid (a few CTAs are enough to saturate NVLink
// with 16-byte stores) so that it shares the SMs with the step kernel it runs under.  When the last CTA is done —
// all stores fenced at system scope — it publishes the segment: peer r's flag word [my rank] = seq.
__global__ void __launch_bounds__(256)
push_rows_kernel(const float* __restrict__ Zl, PeerPtrs peers, int npeers, long long row0, int nrows, int ZS, int C,
                 unsigned long long flag_off_floats, int my_rank, unsigned long long seq, unsigned int* done_counter)
{
    constexpr int U = 8;                         // independent 16-byte loads in flight per thread
    const long long total = (long long)nrows * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * U) {
        float4 v[U];
        size_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long t = t0 + u * stride;
            const long long row = t / C;
            const int piece = (int)(t - row * C);
            off[u] = (size_t)(row0 + row) * ZS + piece * 4;
            if (t < total) v[u] = __ldcg(reinterpret_cast<const float4*>(Zl + off[u]));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t0 + u * stride < total)
                for (int r = 0; r < npeers; ++r) *reinterpret_cast<float4*>(peers.p[r] + off[u]) = v[u];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
        __threadfence();                         // acquire side of the ticket: the other CTAs' (fenced) stores happen-before the flag stores below
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *done_counter = 0u;
        if ((int)threadIdx.x < npeers) {
            unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.p[threadIdx.x] + flag_off_floats) + my_rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
        }
    }
}

// Copy-engine transport: the rows went out as one peer copy per rank on the same stream; stream order makes them
// complete (and visible at the destination) before this kernel starts, which then only publishes the segment.
__global__ void publish_flags_kernel(PeerPtrs peers, int npeers, unsigned long long flag_off_floats, int my_rank,
                                     unsigned long long seq)
{
    if ((int)threadIdx.x < npeers) {
        __threadfence_system();
        unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.p[threadIdx.x] + flag_off_floats) + my_rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    }
}

// Stream-ordered wait until every peer has published `need` segments (end of a run; the gathering kernels poll the
// same words in their own prologue, so no such kernel sits between two K-blocks).
__global__ void wait_flags_kernel(const unsigned long long* flags, int world, int my_rank, unsigned long long need)
{
    demcz::wait_peer_segments(flags, world, my_rank, need);
}
}  // namespace

extern "C" int demcz_comm_peer_info(demcz_ctx* c, uint8_t info[128])
{
    if (!c || !info) return DEMCZ_ERR_INVALID;
    if (c->sticky) return c->sticky;
    if (!c->archive_set) return demcz_fail(c, DEMCZ_ERR_INVALID, "comm_peer_info: call demcz_set_archive first");
    if (cudaSetDevice(c->device) != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "cudaSetDevice");
    PeerInfo pi;
    memset(&pi, 0, sizeof pi);
    pi.pid = (int64_t)getpid();
    pi.device = c->device; pi.ZS = c->ZS; pi.ptr = (uint64_t)(uintptr_t)c->dZ; pi.zcap = c->zcap;
    if (cudaIpcGetMemHandle(&pi.handle, c->dZ) != cudaSuccess) {
        cudaGetLastError();              // same-process peers do not need the handle
        memset(&pi.handle, 0, sizeof pi.handle);
    }
    memcpy(info, &pi, sizeof pi);
    return DEMCZ_OK;
}

extern "C" int demcz_comm_open_peers(demcz_ctx* c, const uint8_t* infos)
{
    if (!c || !infos) return DEMCZ_ERR_INVALID;
    if (c->sticky) return c->sticky;
    if (c->o.world == 1) return DEMCZ_OK;
    if (c->o.world > 8) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "comm_open_peers: at most 8 ranks (one NVSwitch box)");
    if (!c->comm) return demcz_fail(c, DEMCZ_ERR_INVALID, "comm_open_peers: call demcz_comm_init first");
    if (cudaSetDevice(c->device) != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "cudaSetDevice");
    for (int r = 0; r < c->o.world; ++r) {
        if (r == c->o.rank) continue;
        PeerInfo pi;
        memcpy(&pi, infos + (size_t)r * 128, sizeof pi);
        if (pi.ZS != c->ZS || pi.zcap != c->zcap)
            return demcz_fail(c, DEMCZ_ERR_INVALID, "comm_open_peers: rank %d has a different archive geometry", r);
        if (pi.pid == (int64_t)getpid()) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->device, pi.device);
            if (!can) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "comm_open_peers: no P2P access from device %d to %d", c->device, pi.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(pi.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return demcz_fail(c, DEMCZ_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", pi.device, cudaGetErrorString(e));
            cudaGetLastError();
            c->peerZ[r] = (float*)(uintptr_t)pi.ptr;
            c->peer_ipc[r] = false;
        } else {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, pi.handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            c->peerZ[r] = (float*)p;
            c->peer_ipc[r] = true;
        }
    }
    c->peers_open = true;
    // stream memory operations (DEMCZ_MEMOPS=0 switches them off): probe the write on every peer's flag word of this rank
    // (value 0 = its initial content) — a driver that refuses peer addresses leaves the one-warp publish kernel in place
    {
        const char* env = getenv("DEMCZ_MEMOPS");
        const bool on = !(env && atoi(env) == 0);
        MemOps& m = memops();
        c->memop_wait = on && m.wait != nullptr;
        c->memop_write = on && m.write != nullptr;
        if (c->memop_write) {
            for (int r = 0; r < c->o.world && c->memop_write; ++r) {
                if (r == c->o.rank) continue;
                unsigned long long* f = reinterpret_cast<unsigned long long*>(c->peerZ[r] + demcz_flag_off_floats(c)) + c->o.rank;
                if (m.write((CUstream)c->comm_stream, (CUdeviceptr)(uintptr_t)f, 0u, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS) c->memop_write = false;
            }
            if (cudaStreamSynchronize(c->comm_stream) != cudaSuccess) { cudaGetLastError(); c->memop_write = false; }
        }
        if (c->memop_wait) {                     // and the wait on a local word that already holds the value
            if (m.wait((CUstream)c->comm_stream, (CUdeviceptr)(uintptr_t)(c->d_barrier + 4), 0u, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS ||
                cudaStreamSynchronize(c->comm_stream) != cudaSuccess) { cudaGetLastError(); c->memop_wait = false; }
        }
    }
    // every rank takes the same decision (same opts, same driver).  AUTO (measured): copy engines up to
    // 4 GPUs (26.2 vs 18.0 G chain-updates/s on 2, 42.7 on 4), the push kernel above (8 GPUs: 7 peer copies per K-block move
    // only ~290 GB/s next to the step kernel; SM-issued stores started together with the step kernel reach 545 GB/s)
    c->push_engine = c->o.push_engine == DEMCZ_PUSH_KERNEL ? 0 : c->o.push_engine == DEMCZ_PUSH_COPY_ENGINE ? 1 : (c->o.world <= 4 ? 1 : 0);
    if (const char* env = getenv("DEMCZ_PUSH_ENGINE")) c->push_engine = (atoi(env) == DEMCZ_PUSH_COPY_ENGINE) ? 1 : 0;
    c->inkernel_poll = c->push_engine == 1 && c->memop_write;
    if (!c->inkernel_poll) c->memop_wait = false;                  // the wait kernel path keeps the event hand-off
    c->xs[0] = c->comm_stream;
    c->n_xs = 1;
    if (c->push_engine == 1) {
        // one copy engine moves ~300-350 GB/s next to the step kernel (measured); one stream per peer lets several
        // engines work at once (3 streams on 4 GPUs: 580 GB/s aggregate)
        int want = std::min(7, c->o.world - 1);
        if (const char* env = getenv("DEMCZ_CE_STREAMS")) want = std::max(1, std::min(7, atoi(env)));
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int i = 1; i < want; ++i) {
            if (cudaStreamCreateWithPriority(&c->xs[i], cudaStreamNonBlocking, hi) != cudaSuccess) { cudaGetLastError(); break; }
            c->n_xs = i + 1;
        }
    }
    return DEMCZ_OK;
}

// Completes the archive segment starting at physical row `phys_row0` on every rank.  `side`: issue on the side
// stream with the second communicator (overlap mode), else on the compute stream.  The segment's sequence number is
// appends + 1 (this is segment number `appends`).
int demcz_comm_exchange(demcz_ctx* c, int64_t phys_row0, bool side)
{
    if (c->o.world == 1) return DEMCZ_OK;
    const unsigned long long seq = (unsigned long long)c->appends + 1;
    if (phys_row0 + c->o.N_global > c->zcap)
        return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "multi-GPU append segment wraps the archive ring; make archive_capacity and M0 multiples of N");
    cudaStream_t st = side ? c->comm_stream : c->stream;
    ncclComm_t comm = (ncclComm_t)(side ? c->comm2 : c->comm);
    auto mark = [&](void) {                                       // profile=1: bracket the exchange with timing events
        if (!c->o.profile) return;
        while (c->ev_xchg.size() <= c->xchg_used) { cudaEvent_t ev; cudaEventCreate(&ev); c->ev_xchg.push_back(ev); }
        cudaEventRecord(c->ev_xchg[c->xchg_used++], st);
    };
    if (c->peers_open) {
        PeerPtrs pp;
        int np = 0;
        for (int r = 0; r < c->o.world; ++r) if (r != c->o.rank) pp.p[np++] = c->peerZ[r];
        mark();
        if (c->push_engine == 1) {
            // one contiguous copy per peer: this rank's N row slots (with dense slots: payload only).  Peers in ROTATED
            // order (rank+1, rank+2, ...): at any moment every destination is written by one sender, not by all of them
            // (incast on one ingress link); on the side stream(s) round-robin, the peer's flag word right behind its copy
            const size_t off = (size_t)(phys_row0 + c->o.chain_offset) * c->ZS;
            const size_t bytes = (size_t)c->o.N * c->ZS * sizeof(float);
            for (int i = 1; i < c->o.world; ++i) {
                const int r = (c->o.rank + i) % c->o.world;
                cudaStream_t xs = side ? c->xs[(i - 1) % c->n_xs] : st;
                if (cudaMemcpyAsync(c->peerZ[r] + off, c->dZ + off, bytes, cudaMemcpyDeviceToDevice, xs) != cudaSuccess)
                    return demcz_fail(c, DEMCZ_ERR_CUDA, "peer copy of the appended rows failed: %s", cudaGetErrorString(cudaGetLastError()));
                if (c->memop_write) {               // flag word by a stream memory operation: ordered behind the copy, no SM needed
                    unsigned long long* f = reinterpret_cast<unsigned long long*>(c->peerZ[r] + demcz_flag_off_floats(c)) + c->o.rank;
                    if (memops().write((CUstream)xs, (CUdeviceptr)(uintptr_t)f, (cuuint32_t)seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
                        return demcz_fail(c, DEMCZ_ERR_CUDA, "cuStreamWriteValue32 on a peer flag word failed");
                } else {
                    PeerPtrs one; one.p[0] = c->peerZ[r];
                    publish_flags_kernel<<<1, 32, 0, xs>>>(one, 1, demcz_flag_off_floats(c), c->o.rank, seq);
                }
            }
        } else {
            const int C = c->LD / 4;       // payload pieces only (rounding up to whole 32-byte sectors costs bytes and gains nothing: measured)
            const long long total = (long long)c->o.N * C;
            const int grid = (int)std::min<long long>(32, (total + 255) / 256);      // measured: 16..128 CTAs move the same bytes/s (link-bound)
            unsigned int* done = (unsigned int*)(c->d_barrier + (side ? 1 : 0));
            push_rows_kernel<<<grid, 256, 0, st>>>(c->dZ, pp, np, phys_row0 + c->o.chain_offset, c->o.N, c->ZS, C,
                                                   demcz_flag_off_floats(c), c->o.rank, seq, done);
        }
        if (cudaGetLastError() != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "push launch failed");
        mark();
        return DEMCZ_OK;                   // the next gathering launch polls the flag words itself (demcz_comm_need)
    }
    float* seg = c->dZ + (size_t)phys_row0 * c->ZS;
    const size_t count = (size_t)c->o.N * c->ZS;                 // floats per rank
    mark();
    int rc = api().AllGather(seg + (size_t)c->o.rank * count, seg, count, ncclFloat32, comm, st);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllGather", rc);
    mark();
    return DEMCZ_OK;
}

// Segments that must have landed from every peer before the next launch may gather: every appended segment in strict
// mode, all but the newest in overlap mode (the newest is still travelling under that launch and is not drawable).
// Only the peer-push transport uses flag words; the NCCL all-gather is ordered by its stream / events.
unsigned long long demcz_comm_need(const demcz_ctx* c)
{
    if (c->o.world == 1 || !c->peers_open || !c->inkernel_poll) return 0ull;
    const int64_t n = c->o.sync_mode == DEMCZ_SYNC_OVERLAP ? c->appends - 1 : c->appends;
    return n > 0 ? (unsigned long long)n : 0ull;
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

int demcz_comm_wait_done_word(demcz_ctx* c, unsigned int seq)
{
    for (int i = 0; i < c->n_xs; ++i)
        if (memops().wait((CUstream)c->xs[i], (CUdeviceptr)(uintptr_t)(c->d_barrier + 4), (cuuint32_t)seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            return demcz_fail(c, DEMCZ_ERR_CUDA, "cuStreamWaitValue32 failed");
    return DEMCZ_OK;
}

// compute stream: do not pass until every peer has pushed segments 0 .. nseg-1 into this rank's replica
int demcz_comm_wait_segment(demcz_ctx* c, int64_t nseg)
{
    if (c->o.world == 1 || !c->peers_open || nseg <= 0) return DEMCZ_OK;
    wait_flags_kernel<<<1, 32, 0, c->stream>>>(demcz_flags(c), c->o.world, c->o.rank, (unsigned long long)nseg);
    if (cudaGetLastError() != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "wait_flags_kernel launch failed");
    return DEMCZ_OK;
}
