// demcz_internal.h — context layout and internal entry points shared by the library's
// translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/demcz.h"
#include "philox.cuh"

enum { DEMCZ_LIK_NONE = -1, DEMCZ_LIK_MVN = 0, DEMCZ_LIK_LINREG = 1 };

struct demcz_linreg_state;   // demcz_linreg.cu

struct demcz_ctx {
    demcz_opts o{};
    std::vector<int32_t> block_ptr, block_idx;
    std::vector<double> eps;
    int device = 0;
    cudaStream_t stream = nullptr;
    int LD = 0;          // archive row stride in floats (d rounded up to 4 -> 16-byte aligned rows)
    int ldx = 0;         // SoA leading dimension (N rounded up to 32)
    int ZS = 0;          // archive row slot in floats (power of two >= LD for the closed-form path)
    demcz::PhiloxKeys keys{};
    float shift[64] = {0};

    int lik_kind = DEMCZ_LIK_NONE;
    std::vector<float> mvn_mu, mvn_W;   // W = L^-1, row-major d x d (lower)
    float mvn_c0 = 0.f;
    demcz_linreg_state* linreg = nullptr;

    // device state
    float* dX = nullptr;        // [d][ldx]
    float* dlogp = nullptr;     // [ldx]
    float* dZ = nullptr;        // [zcap][LD]
    int64_t zcap = 0;
    int64_t M = 0;              // rows appended so far (valid rows = min(M, zcap))
    float* dT = nullptr;        // temperature table
    std::vector<float> hT;      // its host copy (the linreg path passes T by value)
    float* dhist = nullptr;     // [G][d][N]   (store_chain = 1: history resident on the device)
    float* dlphist = nullptr;   // [G][N]
    // store_chain = 2: history streamed to pinned host memory, one launch's slice at a time (double-buffered on the
    // device, copied on its own stream under the next launches)
    float* hs_dx[2] = {nullptr, nullptr};        // [slice_gens][d][N]
    void* hs_dlp[2] = {nullptr, nullptr};        // [slice_gens][N] float (closed-form path) or double (linreg)
    float* hs_hx = nullptr;                      // pinned [G][d][N]
    void* hs_hlp = nullptr;                      // pinned [G][N]
    size_t hs_lp_elem = 0;
    int64_t hs_slice_gens = 0, hs_launches = 0;
    cudaStream_t hs_stream = nullptr;
    cudaEvent_t hs_ev_k[2] = {nullptr, nullptr}, hs_ev_c[2] = {nullptr, nullptr};
    uint32_t* tr_r1 = nullptr; uint32_t* tr_r2 = nullptr; float* tr_u = nullptr; float* tr_xprop = nullptr;
    float* tr_lpprop = nullptr; uint8_t* tr_acc = nullptr;
    int* d_chain_changed = nullptr;
    unsigned long long* d_counters = nullptr;   // [0] adaptγ window changed, [1] accepted proposals
    float* dS1[2] = {nullptr, nullptr};
    double* dS2 = nullptr;
    float* dshift = nullptr;
    double* d_scratch = nullptr;
    double* d_stage = nullptr;   // persistent FP64 staging for set_state/get_state: (N*d + N) doubles

    // run state
    int64_t g_done = 0;
    double gamma = 0.0;
    bool state_set = false, state_lp_valid = false, archive_set = false;
    bool stopped_by_rhat = false;
    int64_t proposals = 0;
    std::vector<double> rhat;
    double rhat_max = 0.0;
    bool rhat_valid = false;

    std::vector<cudaEvent_t> ev_pool;
    std::vector<cudaEvent_t> ev_xchg;            // profile=1: brackets of the exchanges of the current run
    size_t xchg_used = 0;
    cudaEvent_t ev_run[2] = {nullptr, nullptr};

    // NCCL (dlopen'ed; demcz_comm.cu) and the peer-to-peer exchange
    void* comm = nullptr;
    void* comm2 = nullptr;                       // communicator of the side stream
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_app[2] = {nullptr, nullptr};  // compute -> side: segment written locally
    cudaEvent_t ev_vis[2] = {nullptr, nullptr};  // side -> compute: segment complete on every rank
    unsigned long long* d_barrier = nullptr;
    float* peerZ[8] = {nullptr};
    bool peer_ipc[8] = {false};
    bool peers_open = false;
    int64_t appends = 0;                         // archive appends so far (overlap mode: visibility lags by one)
    int64_t vis_waited = 0;
    bool step_full = false;                      // the closed-form step launches use the single-full-block kernel
    bool fuse_push = false;                      // overlap + in-place push: issued from inside the next step-kernel launch
    bool push_pending = false;                   // a segment is appended locally but not pushed yet (fused push)
    int64_t push_row0 = 0;                       // its physical first row and
    unsigned long long push_seq = 0;             // the sequence number its flag words will carry
    bool xchg_packed = false;                    // DEMCZ_XCHG_PACKED in effect: rows travel densely through the staging ring
    int64_t unpacked = 0;                        // packed exchange: segments 0 .. unpacked-1 have been scattered into the slots
    int64_t appends_run = 0;                     // appends since create (set_archive does not reset it)

    std::string err;
    int sticky = 0;
};

// One device allocation holds the archive and everything the peers write into, so that a single IPC handle maps it:
//   [zcap row slots][8 x u64 segment flags, 256 B][staging ring: kStageDepth x N_global dense rows of LD floats]   (ring: packed layout)
// Depth 4: a sender refills buffer s % 4 with segment s + 4 only after its K-block s + 4 started, which needed the
// receiver's segment s + 2, which the receiver pushes after its K-block s + 2, which started after it unpacked segment s.
constexpr int kStageDepth = 4;
inline size_t demcz_flag_off_floats(const demcz_ctx* c) { return (size_t)c->zcap * c->ZS; }
inline size_t demcz_stage_off_floats(const demcz_ctx* c) { return (demcz_flag_off_floats(c) + 64 + 31) / 32 * 32; }
inline size_t demcz_stage_floats(const demcz_ctx* c)     // only the packed layout has a staging ring
{
    return (c->o.world > 1 && c->o.exchange_layout == DEMCZ_XCHG_PACKED) ? (size_t)kStageDepth * c->o.N_global * c->LD : 0;
}

int demcz_fail(demcz_ctx* c, int code, const char* fmt, ...);
int demcz_archive_view(demcz_ctx* c, uint32_t* Mdraw, uint32_t* zstart);
int demcz_archive_appended(demcz_ctx* c);
// history destination of the launch covering generations [g, g + n_gens) (0-based): pointers such that index (grel - 1)
// with grel = g + 1 .. g + n_gens is valid; lp elements are float (closed-form path) or double (linreg)
int demcz_hist_begin(demcz_ctx* c, int64_t g, int n_gens, float** hist, void** lphist);
int demcz_hist_end(demcz_ctx* c, int64_t g, int n_gens);

// demcz_comm.cu — no-ops when world == 1
int demcz_comm_exchange(demcz_ctx* c, int64_t phys_row0, bool side);
int demcz_comm_flush_pending_push(demcz_ctx* c);             // fused push: the segment no later launch will carry (end of a run)
int demcz_comm_wait_segment(demcz_ctx* c, int64_t nseg);     // + packed exchange: scatters the landed segments into their slots
int demcz_comm_allreduce_f64(demcz_ctx* c, double* dev, int n);
int demcz_comm_allreduce_u64(demcz_ctx* c, unsigned long long* dev, int n);
void demcz_comm_destroy_internal(demcz_ctx* c);

// demcz_linreg.cu
// out_dev == nullptr: evaluate the context's own N chains into its FP64 log-objective array
int demcz_linreg_eval(demcz_ctx* c, const float* X, int n, int ldx, double* out_dev);
int demcz_linreg_generations(demcz_ctx* c, int64_t g, int n_gens, bool append, bool acc, int half);
void demcz_linreg_free(demcz_ctx* c);
double* demcz_linreg_lp(demcz_ctx* c);           // [ldx] FP64 log-objective of the chains (logpi ~ -n/2 needs more than FP32)
double* demcz_linreg_lphist(demcz_ctx* c);       // [G][N]
unsigned long long demcz_linreg_launches(demcz_ctx* c);
double demcz_linreg_profile_ms(demcz_ctx* c, long long* launches);
