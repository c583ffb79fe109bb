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

struct demcz_linreg_state;   // streaming moments (demcz_lib.cu)
int demcz_moments_begin(demcz_ctx* c);                       // fixes the shift (population mean) at the first generation
int demcz_moments_add_state(demcz_ctx* c);                   // adds the current SoA state (one generation) — linreg path
int demcz_moments_get(demcz_ctx* c, double* mean, double* cov, int64_t* n, int32_t reset);

// demcz_linreg.cu

struct demcz_ctx {
    demcz_opts o{};
    std::vector<int32_t> block_ptr, block_idx;
    std::vector<double> eps;
    int device = 0;
    cudaStream_t stream = nullptr;
    int LD = 0;          // archive row stride in floats (d rounded up to 4 -> 16-byte aligned rows)
    int ldx = 0;         // SoA leading dimension (N rounded up to 32)
    int ZS = 0;          // archive row slot in floats (power of two >= LD for the closed-form path)
    int thin = 1;        // history thinning (opts.thin, at least 1)
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
    // stream memory operations (cuStreamWaitValue32 / cuStreamWriteValue32): the side stream waits on a word the appending
    // step kernel writes (no event on the compute stream -> consecutive step kernels keep their programmatic overlap) and
    // publishes the segment flags without a kernel (no SM slot needed while the step kernel fills the GPU)
    bool memop_wait = false, memop_write = false;
    // The gathering kernels may spin on the peers' flag words ONLY if publishing a segment needs no SM on the sending GPU
    // (copy engines + cuStreamWriteValue32): a publish/push KERNEL could find every SM slot of its GPU held by that GPU's
    // own spinning step kernel -> all ranks wait for each other (observed: 8 GPUs, timed out in flag_wait).  Otherwise a
    // one-warp wait kernel sits between the step kernels (it holds one slot only).
    bool inkernel_poll = false;
    cudaStream_t xs[8] = {nullptr};                 // copy-engine transport: side streams (xs[0] = comm_stream), up to one per peer
    int n_xs = 1;
    bool last_append_by_step_kernel = false;     // the segment just appended carries a done word (closed-form path)
    int push_engine = 0;                         // 0: our push kernel (SM stores over NVLink), 1: copy engines (one peer copy per rank) + flag kernel
    int64_t appends_run = 0;                     // appends since create (set_archive does not reset it)
    bool persistent_ok = false;                  // small-N persistent launches in effect (opts.persistent)

    // opts.moments: streaming first / second moments of every generation of every chain (mean_cov_chain of the full
    // history without the history).  Sums of (x - mom_shift): d_mom = [S1: 64 doubles][S2: 64 x 64 doubles, row-major]
    double* d_mom = nullptr;
    float* dmomshift = nullptr;                  // device copy of mom_shift
    float mom_shift[64] = {0};
    bool mom_shift_set = false;
    int64_t mom_gens = 0;                        // generations accumulated (x N_global points each)

    std::string err;
    int sticky = 0;
};

// One device allocation holds the archive and everything the peers write into, so that a single IPC handle maps it:
//   [zcap row slots][8 x u64 segment flags, 256 B]
inline size_t demcz_flag_off_floats(const demcz_ctx* c) { return (size_t)c->zcap * c->ZS; }
inline size_t demcz_alloc_floats(const demcz_ctx* c) { return (demcz_flag_off_floats(c) + 64 + 31) / 32 * 32; }
inline const unsigned long long* demcz_flags(const demcz_ctx* c)
{
    return reinterpret_cast<const unsigned long long*>(c->dZ + demcz_flag_off_floats(c));
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
int demcz_comm_wait_done_word(demcz_ctx* c, unsigned int seq);   // side stream: cuStreamWaitValue32 on the step kernel's done word
int demcz_comm_wait_segment(demcz_ctx* c, int64_t nseg);     // stream-ordered wait (end of a run; the step kernels poll the flags themselves)
// how many segments must have landed from every peer before the next gathering launch (0: no wait); see demcz_archive_view
unsigned long long demcz_comm_need(const demcz_ctx* c);
int demcz_comm_allreduce_f64(demcz_ctx* c, double* dev, int n);
int demcz_comm_allreduce_u64(demcz_ctx* c, unsigned long long* dev, int n);
void demcz_comm_destroy_internal(demcz_ctx* c);

// streaming moments (demcz_lib.cu)
int demcz_moments_begin(demcz_ctx* c);                       // fixes the shift (population mean) at the first generation
int demcz_moments_add_state(demcz_ctx* c);                   // adds the current SoA state (one generation) — linreg path
int demcz_moments_get(demcz_ctx* c, double* mean, double* cov, int64_t* n, int32_t reset);

// demcz_linreg.cu
// out_dev == nullptr: evaluate the context's own N chains into its FP64 log-objective array
int demcz_linreg_eval(demcz_ctx* c, const float* X, int n, int ldx, double* out_dev);
int demcz_linreg_generations(demcz_ctx* c, int64_t g, int n_gens, bool append, bool acc, int half, bool snooker = false);
bool demcz_is_snooker_gen(const demcz_ctx* c, long long gen_abs);
void demcz_linreg_free(demcz_ctx* c);
double* demcz_linreg_lp(demcz_ctx* c);           // [ldx] FP64 log-objective of the chains (logpi ~ -n/2 needs more than FP32)
double* demcz_linreg_lphist(demcz_ctx* c);       // [G][N]
unsigned long long demcz_linreg_launches(demcz_ctx* c);
double demcz_linreg_profile_ms(demcz_ctx* c, long long* launches);
std::string demcz_linreg_info(demcz_ctx* c);     // "linreg_kind=...;..." or ""
