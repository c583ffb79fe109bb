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
    if (cudaMalloc((void**)&c->d_barrier, 4 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(c->d_barrier, 0, 4 * sizeof(unsigned long long)) != cudaSuccess)
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
//
// PACKED (DEMCZ_XCHG_PACKED): the destination is not the slot but `packed_base + 4 t` floats into the peer's
// allocation — this rank's block of the staging ring, rows back to back.  NVLink moves whole 128-byte lines and a
// partly written line costs a full one (tools/push_bench.cu: 5.2 G lines/s either way), so 80-byte rows of a
// 128-byte slot cross the links 1.55x faster packed; the receiver scatters them (unpack_rows_kernel).
template <bool PACKED>
__global__ void __launch_bounds__(256)
push_rows_kernel(const float* __restrict__ Zl, PeerPtrs peers, int npeers, long long row0, int nrows, int ZS, int C,
                 unsigned long long packed_base, unsigned long long flag_off_floats, int my_rank, unsigned long long seq,
                 unsigned int* done_counter)
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
            const size_t src = (size_t)(row0 + row) * ZS + piece * 4;
            off[u] = PACKED ? (size_t)packed_base + (size_t)t * 4 : src;
            if (t < total) v[u] = __ldcg(reinterpret_cast<const float4*>(Zl + src));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t0 + u * stride < total)
                for (int r = 0; r < npeers; ++r) *reinterpret_cast<float4*>(peers.p[r] + off[u]) = v[u];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *done_counter = 0u;
        if ((int)threadIdx.x < npeers) {
            unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.p[threadIdx.x] + flag_off_floats) + my_rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
        }
    }
}

// Receiver side of the packed exchange: the other ranks' rows of one landed segment, staging ring -> row slots.
// `stage`: the segment's buffer [N_global][4 C] floats; rows of this rank's own block [own0, own0 + n_own) are already
// in place (written by the append).  An odd piece count gets its last 32-byte sector completed with slot padding,
// like the append does (whole sectors need no read-modify-write fill in L2).
__global__ void __launch_bounds__(256)
unpack_rows_kernel(const float* __restrict__ stage, float* __restrict__ Zl, long long seg_row0, long long n_rows,
                   long long own0, long long n_own, int ZS, int C, int CW)
{
    constexpr int U = 4;
    const long long total = (n_rows - n_own) * CW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * U) {
        float4 v[U];
        size_t dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long t = t0 + u * stride;
            long long row = t / CW;
            const int piece = (int)(t - row * CW);
            if (row >= own0) row += n_own;
            dst[u] = (size_t)(seg_row0 + row) * ZS + piece * 4;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < total && piece < C) v[u] = __ldcg(reinterpret_cast<const float4*>(stage + (size_t)row * (4 * C) + piece * 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t0 + u * stride < total) *reinterpret_cast<float4*>(Zl + dst[u]) = v[u];
    }
}

// Before a kernel that draws from segment `need - 1`: wait (normally zero spins) until every peer has published it.
__global__ void wait_flags_kernel(const unsigned long long* flags, int world, int my_rank, unsigned long long need)
{
    const int r = threadIdx.x;
    if (r >= world || r == my_rank) return;
    for (unsigned int spin = 0; spin < (1u << 28); ++spin) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
        if (v >= need) return;
        __nanosleep(200);
    }
    __trap();                                   // a peer never delivered: fail the launch instead of hanging
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
    // every rank takes the same decision (same opts); with dense slots (ZS == LD) the in-place push is already packed
    const int lay = c->o.exchange_layout;
    c->xchg_packed = c->ZS != c->LD && lay == DEMCZ_XCHG_PACKED;
    // overlap mode, rows stored in place: with DEMCZ_FUSED_PUSH=1 the next step-kernel launch carries the push itself
    // (demcz_step.cuh, PUSH).  Opt-in: measured at parity with the separate kernel on 2 GPUs.
    const char* env = getenv("DEMCZ_FUSED_PUSH");
    c->fuse_push = c->o.sync_mode == DEMCZ_SYNC_OVERLAP && !c->xchg_packed && env && atoi(env) == 1;
    return DEMCZ_OK;
}

// Completes the archive segment starting at physical row `phys_row0` on every rank.  `side`: issue on the side
// stream with the second communicator (overlap mode), else on the compute stream.
static int exchange_seq(demcz_ctx* c, int64_t phys_row0, bool side, unsigned long long seq);

int demcz_comm_exchange(demcz_ctx* c, int64_t phys_row0, bool side)
{
    return exchange_seq(c, phys_row0, side, (unsigned long long)c->appends + 1);       // this is segment number `appends`
}

int demcz_comm_flush_pending_push(demcz_ctx* c)
{
    if (!c->push_pending) return DEMCZ_OK;
    c->push_pending = false;
    const bool packed = c->xchg_packed;          // (never both; the stand-alone kernel stores in place)
    c->xchg_packed = false;
    int rc = exchange_seq(c, c->push_row0, false, c->push_seq);
    c->xchg_packed = packed;
    return rc;
}

static int exchange_seq(demcz_ctx* c, int64_t phys_row0, bool side, unsigned long long seq)
{
    if (c->o.world == 1) return DEMCZ_OK;
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
        const int C = c->LD / 4;       // payload pieces only (rounding up to whole 32-byte sectors costs bytes and gains nothing: measured)
        const long long total = (long long)c->o.N * C;
        const int grid = (int)std::min<long long>(32, (total + 255) / 256);      // measured: 16..128 CTAs move the same bytes/s (link-bound)
        unsigned int* done = (unsigned int*)(c->d_barrier + (side ? 1 : 0));
        mark();
        if (c->xchg_packed) {
            const int buf = (int)((seq - 1) % kStageDepth);
            const unsigned long long base = demcz_stage_off_floats(c) + ((size_t)buf * c->o.N_global + c->o.chain_offset) * c->LD;
            push_rows_kernel<true><<<grid, 256, 0, st>>>(c->dZ, pp, np, phys_row0 + c->o.chain_offset, c->o.N, c->ZS, C, base,
                                                         demcz_flag_off_floats(c), c->o.rank, seq, done);
        } else {
            push_rows_kernel<false><<<grid, 256, 0, st>>>(c->dZ, pp, np, phys_row0 + c->o.chain_offset, c->o.N, c->ZS, C, 0ull,
                                                          demcz_flag_off_floats(c), c->o.rank, seq, done);
        }
        if (cudaGetLastError() != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "push_rows_kernel launch failed");
        mark();
        if (!side) return demcz_comm_wait_segment(c, (int64_t)seq);               // strict: publish before the next launch
        return DEMCZ_OK;
    }
    float* seg = c->dZ + (size_t)phys_row0 * c->ZS;
    const size_t count = (size_t)c->o.N * c->ZS;                 // floats per rank
    mark();
    int rc = api().AllGather(seg + (size_t)c->o.rank * count, seg, count, ncclFloat32, comm, st);
    if (rc != ncclSuccess) return nccl_fail(c, "ncclAllGather", rc);
    mark();
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

// compute stream: do not pass until every peer has pushed segments 0 .. nseg-1 into this rank's replica
int demcz_comm_wait_segment(demcz_ctx* c, int64_t nseg)
{
    if (c->o.world == 1 || !c->peers_open || nseg <= 0) return DEMCZ_OK;
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(c->dZ + demcz_flag_off_floats(c));
    wait_flags_kernel<<<1, 32, 0, c->stream>>>(flags, c->o.world, c->o.rank, (unsigned long long)nseg);
    if (cudaGetLastError() != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "wait_flags_kernel launch failed");
    if (c->xchg_packed) {
        // segment s was appended at M_base + s N_global; M and appends advance together (demcz_archive_appended)
        const int64_t M_base = c->M - c->appends * c->o.N_global;
        const int C = c->LD / 4, CP = c->ZS / 4, CW = C + ((C % 2 == 1 && C < CP) ? 1 : 0);
        const long long total = (long long)(c->o.N_global - c->o.N) * CW;
        const int grid = (int)std::min<long long>(148 * 4, (total + 256 * 4 - 1) / (256 * 4));
        for (; c->unpacked < nseg; ++c->unpacked) {
            const int64_t s = c->unpacked;
            const float* stage = c->dZ + demcz_stage_off_floats(c) + (size_t)(s % kStageDepth) * c->o.N_global * c->LD;
            unpack_rows_kernel<<<grid, 256, 0, c->stream>>>(stage, c->dZ, (M_base + s * c->o.N_global) % c->zcap, c->o.N_global,
                                                            c->o.chain_offset, c->o.N, c->ZS, C, CW);
            if (cudaGetLastError() != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "unpack_rows_kernel launch failed");
        }
    }
    return DEMCZ_OK;
}
