/*
 * demcz.h — C ABI of libdemcz_cuda.so: the B200-native DEMCz generation loop.
 *
 * Drop-in boundary for the hot path of chrished/DEMC.jl.  The reference has no FFI: its boundary
 * is the Julia call signature itself.  Each entry point below names the reference interface
 * (file:line under the upstream repository) whose work it takes over; the Julia `ccall` shim that
 * binds them is demc.jl_b200/julia/DEMC.jl and the ctypes binding is demc.jl_b200/_lib.py
 * (INTEGRATION.md shows both).
 *
 * Conventions
 *   - Plain C, no CUDA/torch types.  All pointers passed in are HOST pointers borrowed for the
 *     duration of the call; all outputs are caller-allocated.  Matrices are Julia-native:
 *     Float64, column-major, 1 generation-loop "row" = 1 point.
 *   - One context drives ONE GPU (one process per GPU; multi-GPU = one context per rank, wired
 *     with demcz_comm_*).  A context is not thread-safe; distinct contexts are independent.
 *   - Every function returns DEMCZ_OK (0) or a negative demcz_status; demcz_last_error() gives
 *     the text.  CUDA / NCCL errors are sticky on the context.  Nothing throws or aborts, and
 *     there is NO CPU fallback: without a CUDA device demcz_create fails with DEMCZ_ERR_CUDA.
 */
#ifndef DEMCZ_H
#define DEMCZ_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEMCZ_VERSION 200   /* 0.2.0 */
#define DEMCZ_MAX_DIM 32    /* register-resident closed-form likelihood path */
#define DEMCZ_MAX_DIM_LINREG 64

typedef struct demcz_ctx demcz_ctx;

typedef enum {
    DEMCZ_OK = 0,
    DEMCZ_ERR_INVALID = -1,      /* bad argument / call order */
    DEMCZ_ERR_CUDA = -2,
    DEMCZ_ERR_NCCL = -3,
    DEMCZ_ERR_OOM = -4,
    DEMCZ_ERR_NOT_PD = -5,       /* Sigma not positive definite */
    DEMCZ_ERR_UNSUPPORTED = -6
} demcz_status;

enum { DEMCZ_INIT_LAST_N_OF_Z = 0, DEMCZ_INIT_REF_SERIAL_ZEROS = 1 };   /* SURVEY Q1 */
/* STRICT: rows appended after generation g are drawable from generation g+1 (bit-identical to one GPU).
 * OVERLAP: they become drawable one K-block later, on every rank and for every world size (also 1), so that the
 * exchange of a segment runs under the next K-block's compute; results are still independent of the sharding. */
enum { DEMCZ_SYNC_STRICT = 0, DEMCZ_SYNC_OVERLAP = 1 };
/* How a rank's appended rows are laid out in the (replicated) archive and therefore how they cross NVLink once
 * demcz_comm_open_peers was called.  A row of d <= 32 floats normally sits in a power-of-two slot (32/64/128 bytes) so
 * that a gather never straddles a 128-byte line; NVLink also moves whole 128-byte lines (a partial line costs a full
 * one: 5.2 G lines/s measured), so
 *   IN_PLACE   : rows stored straight into the peers' padded slots (best while the push hides under compute),
 *   DENSE_SLOTS: no padding — the archive keeps rows back to back (slot = 4*ceil(d/4) floats), a rank's appended rows
 *                are one contiguous block and cross the links as whole lines; the gather pays for rows that straddle
 *                a line instead (1.39x the DRAM bytes per row at d = 20).  Fixed at demcz_create (it is the layout),
 *   AUTO       : DENSE_SLOTS when world >= 4, else IN_PLACE (measured: 4 GPUs 42.7 vs 31.7 G chain-updates/s on C4).
 *                Results are identical in every layout.
 * (Value 2 was the `packed` staging-ring exchange of version 100; removed: never better than DENSE_SLOTS.) */
enum { DEMCZ_XCHG_AUTO = 0, DEMCZ_XCHG_IN_PLACE = 1, DEMCZ_XCHG_DENSE_SLOTS = 3 };
/* Who moves the rows: our push kernel (SM-issued 16-byte stores into every peer replica) or the copy engines (one
 * contiguous peer copy per rank and segment, no SM time; flags published by a one-warp kernel behind the copies). */
enum { DEMCZ_PUSH_AUTO = 0, DEMCZ_PUSH_KERNEL = 1, DEMCZ_PUSH_COPY_ENGINE = 2 };

/* Mirror of DEMCopt (src/DEMC.jl:24-39) + the keyword arguments of demcz_sample (src/demcz.jl:9)
 * and demcz_anneal (src/demcz_anneal.jl:18), as a POD.  Field meaning == reference option of the
 * same name.  The shim fills it from the Julia struct. */
typedef struct {
    int32_t struct_size;        /* = sizeof(demcz_opts), ABI check */
    int32_t N;                  /* chains on THIS context (opts.N / world when sharded) */
    int32_t K;                  /* append the population to Z every K generations */
    int32_t Ngeneration;
    int32_t Nblocks;
    int32_t d;                  /* Npar */
    const int32_t* block_ptr;   /* CSR of blockindex: Nblocks+1 offsets ...               */
    const int32_t* block_idx;   /* ... into 0-based column indices (blockindex[ib] .- 1)   */
    const double* eps_scale;    /* d.  The jitter eps_scale .* randn (src/demcz.jl:184,186) is drawn with a 16-bit Box-Muller
                                   (|z| <= 4.86, 2^16 levels): meant for the reference's eps_scale ~ 1e-5, not as the main proposal noise */
    double gamma;               /* γ */
    double T0, TN;              /* annealing end points (only used to build the default table) */
    int32_t anneal;             /* 0: demcz_sample semantics, 1: demcz_anneal semantics */
    int32_t autostop;           /* 1 = :Rhat */
    int32_t autostop_every;
    double autostop_Rhat;
    int32_t adapt;              /* adaptγ["adapt"]  (src/demcz_anneal.jl:14) */
    double min_gamma, max_gamma;
    int32_t adapt_every;
    uint64_t seed;              /* Philox key */
    int64_t gen_offset;         /* generations already run by earlier calls (prevrun): the Philox
                                   generation counter continues at gen_offset + 1 */
    int32_t store_chain;        /* MC.chain / MC.log_obj (N x d x G): 0 = not kept, 1 = resident on the device, 2 = streamed to pinned
                                   host memory during the run (one launch's slice at a time, copied under the next launches) */
    int32_t init_mode;          /* DEMCZ_INIT_* */
    int64_t archive_capacity;   /* rows; 0 -> M0 + ceil(N_global*Ngeneration/K) (src/demcz.jl:11) */
    int32_t trace;              /* 1: record per-proposal r1,r2,u,xprop,lp_prop,accept (tests) */
    int32_t profile;            /* 1: time every step-kernel launch with CUDA events */
    /* sharding (one context per GPU). Single GPU: rank 0, world 1, N_global = N, chain_offset 0 */
    int32_t rank, world;
    int64_t N_global;
    int64_t chain_offset;       /* global id of local chain 0 (Philox counter, archive row order) */
    int32_t sync_mode;          /* DEMCZ_SYNC_* */
    int32_t exchange_layout;    /* DEMCZ_XCHG_* (0 -> AUTO) */
    /* ---- version 200 ---- */
    int32_t push_engine;        /* DEMCZ_PUSH_* (0 -> AUTO) */
    int32_t thin;               /* history thinning (store_chain != 0): generation g is kept iff g % thin == 0, so MC.chain is
                                   N x d x floor(G/thin); 0 or 1 = every generation (the reference's behaviour) */
    int32_t moments;            /* 1: accumulate the first and second moments of EVERY generation of every chain on the device
                                   (demcz_get_history_moments = mean_cov_chain of the full history without keeping it) */
    int32_t persistent;         /* small N (one CTA): 1 = run all K-blocks between two host decisions in ONE launch, appends made
                                   visible by a CTA barrier; 0 = one launch per K-block; -1 = auto (on when it applies) */
    double snooker_prob;        /* probability that a generation uses the snooker update of ter Braak & Vrugt (2008), chosen per
                                   generation from the Philox stream; 0 = never (the reference implements the parallel update only) */
    double snooker_gamma;       /* snooker jump size; 0 -> 1.7 */
} demcz_opts;

typedef struct {
    int64_t generations_done;   /* in this call */
    int64_t generations_total;  /* since create */
    int32_t stopped_by_rhat;
    int32_t rhat_valid;
    double rhat_max;
    double final_gamma;
    int64_t M;                  /* archive rows now valid */
    int64_t proposals;          /* since create, this rank */
    int64_t accepted;
    int64_t step_launches;      /* step-kernel launches in this call */
    double step_kernel_ms;      /* sum of their device durations (profile=1), else 0 */
    double device_ms;           /* whole call on the context's stream, CUDA events */
    double exchange_ms;         /* sum of the device durations of the archive exchanges of this call (profile=1) */
    int64_t exchanges;
} demcz_run_info;

int demcz_version(void);

/* replaces: DEMC.demcopt + the prologue of demcz_sample/demcz_anneal (src/DEMC.jl:41-43,
 * src/demcz.jl:9-24, src/demcz_anneal.jl:18-33).  device = CUDA ordinal. */
int demcz_create(demcz_ctx** out, const demcz_opts* opts, int device);
void demcz_destroy(demcz_ctx* ctx);
const char* demcz_last_error(const demcz_ctx* ctx);   /* ctx may be NULL: last create error */

/* replaces the `logobj` closure argument (src/demcz.jl:189) for the likelihoods the reference's
 * tests use.  MvNormal: test/example_normpdf.jl:9-16 (drop_constant=1 with Sigma = I/2 gives the
 * quadratic objective of test/test_anneal.jl:10).  Sigma: d x d column-major. */
int demcz_set_likelihood_mvnormal(demcz_ctx* ctx, int32_t d, const double* mu, const double* Sigma,
                                  int32_t drop_constant);
/* Linreg: -0.5*sum((y .- X*b).^2), test/example_linreg.jl:32.  X: n x p column-major. */
int demcz_set_likelihood_linreg(demcz_ctx* ctx, int64_t n, int32_t p, const double* X, const double* y);

/* A further descriptor on the same tensor-core contraction (not in the reference, whose users pass any
 * closure at src/demcz.jl:189): Bernoulli-logit GLM log-likelihood sum_i [y_i eta_i - log(1 + exp(eta_i))], eta = X b,
 * y_i in {0,1} (any real y is accepted: the y term is linear).  X: n x p column-major. */
int demcz_set_likelihood_logistic(demcz_ctx* ctx, int64_t n, int32_t p, const double* X, const double* y);

/* replaces: Zmat argument + vcat pre-extension (src/demcz.jl:10-12).  Z: M0 x d column-major. */
int demcz_set_archive(demcz_ctx* ctx, const double* Z, int64_t M0);
/* replaces: prevrun state (src/demcz.jl:20-21).  X: N x d col-major; logp: N or NULL (evaluate).
 * Without this call the state is initialised from the archive per init_mode (src/demcz.jl:15-17). */
int demcz_set_state(demcz_ctx* ctx, const double* X, const double* logp);
/* replaces: temperaturefun(ig, Ngen, T0, TN) evaluated per generation (src/demcz_anneal.jl:69);
 * T[ig-1] for ig = 1..len.  Default (no call): tempbaseline (src/demcz_anneal.jl:1-3) from T0,TN. */
int demcz_set_temperature_table(demcz_ctx* ctx, const double* T, int64_t len);

/* replaces: the generation loop `for ig, for ic -> runchain!` with autostop / adaptγ
 * (src/demcz.jl:30-56, 80-93, 167-203; src/demcz_anneal.jl:39-58, 67-80, 142-178).
 * Runs up to max_generations more generations (<=0: all remaining). Blocking. */
int demcz_run(demcz_ctx* ctx, int64_t max_generations, demcz_run_info* info);
/* which kernels this context runs, as "key=value;..." text (measurement / reporting; no reference counterpart):
 * linreg_kind (f16 | tf32), linreg_terms, linreg_kernel, linreg_fallback (why the FP16 kind was not used), step_persistent.
 * Returns the length needed (>= 0); writes at most len-1 characters + NUL. */
int demcz_get_kernel_info(demcz_ctx* ctx, char* buf, int64_t len);
/* measurement switch (no reference counterpart): 1 = bracket every step-kernel launch of later
 * demcz_run calls with CUDA events on the context's stream (demcz_run_info.step_kernel_ms). */
int demcz_set_profile(demcz_ctx* ctx, int32_t on);

/* replaces: the MC result struct (src/DEMC.jl:10-15).  chain: N x d x G, log_obj: N x G
 * (G = generations_total), column-major Float64.  Needs store_chain=1. */
int demcz_get_chain(demcz_ctx* ctx, double* chain, double* log_obj);
int demcz_get_state(demcz_ctx* ctx, double* X, double* logp);            /* Xcurrent, log_objcurrent */
/* returned Zmat (src/demcz.jl:58-62; always Z[1:M,:]). Z may be NULL to query M. */
int demcz_get_archive(demcz_ctx* ctx, double* Z, int64_t* M);
/* population mean / covariance of Xcurrent (1/N normalisation, cf. mean_cov_chain src/utils.jl:96-106) */
int demcz_get_moments(demcz_ctx* ctx, double* mean, double* cov);
/* mean / covariance (1/n normalisation) of archive rows [first_row, first_row + n_rows) in logical order.  The rows
 * appended during a run are the population every K-th generation, so this is mean_cov_chain (src/utils.jl:96-106) of
 * the K-thinned history without keeping MC.chain: the posterior summary for runs with store_chain = 0. */
int demcz_get_archive_moments(demcz_ctx* ctx, int64_t first_row, int64_t n_rows, double* mean, double* cov);
/* mean / covariance (1/n normalisation) over every generation of every chain since create (or since the last call
 * with reset != 0): flatten_chain + mean_cov_chain (src/utils.jl:22-32, 96-106) of the FULL history from streaming
 * device accumulators — no MC.chain needed.  Needs opts.moments = 1.  *n = points accumulated (N_global * generations). */
int demcz_get_history_moments(demcz_ctx* ctx, double* mean, double* cov, int64_t* n, int32_t reset);
/* 64-bit order-independent checksum of archive rows [first_row, first_row + n_rows) in logical order (bit patterns of
 * the FP32 rows mixed with their logical position): equal on every rank of a sharded run iff the replicas are identical.
 * No reference counterpart (the reference shares one Z through a SharedArray, src/demcz.jl:109). */
int demcz_get_archive_checksum(demcz_ctx* ctx, int64_t first_row, int64_t n_rows, uint64_t* checksum);
/* The drawable view of the archive ring as a pure function (no context, no device): which rows a K-block may draw from
 * so that neither this rank's next append nor — in overlap mode — a peer's push of the segment after it can touch them.
 * M rows appended so far, ring capacity zcap, N_global rows per segment, overlap = DEMCZ_SYNC_OVERLAP, `appends` segments
 * appended so far on this archive, `remaining` appends the run will still make.  Outputs: number of drawable rows and the
 * physical row of logical row 0.  Exported for the CPU model check of the exchange protocol (tests/). */
int demcz_archive_view_rule(int64_t M, int64_t zcap, int64_t N_global, int32_t overlap, int64_t appends, int64_t remaining,
                            int64_t* m_out, int64_t* z0_out);
/* last split-Rhat vector (Rhat_gelman, src/utils.jl:2-20), d values */
int demcz_get_rhat(demcz_ctx* ctx, double* rhat);
/* per-chain count of generations (2..G of this context) whose log-objective changed:
 * numerator of accept_ratio in convergence_check (src/utils.jl:61) */
int demcz_get_accept_counts(demcz_ctx* ctx, int64_t* per_chain);

/* replaces: map(logobj, rows) (src/demcz.jl:17,115).  X: n x d col-major -> out[n] */
int demcz_eval_logdensity(demcz_ctx* ctx, const double* X, int64_t n, double* out);

/* step-wise parity hook (needs trace=1): records of all generations run so far, layouts
 * r1,r2,u,lp_prop,accepted: [G][Nblocks][N]; xprop: [G][Nblocks][d][N]; any pointer may be NULL. */
int demcz_get_trace(demcz_ctx* ctx, uint32_t* r1, uint32_t* r2, float* u, float* xprop, float* lp_prop,
                    uint8_t* accepted);

/* multi-GPU (mirrors the sync mode of demcz_sample_par, src/demcz.jl:129-137): every K generations
 * the ranks exchange their appended rows with an NCCL all-gather in place in the archive.
 * Rank 0 calls demcz_comm_unique_id, the host broadcasts the 128 bytes, every rank calls
 * demcz_comm_init before demcz_run. */
int demcz_comm_unique_id(uint8_t id[128]);
int demcz_comm_init(demcz_ctx* ctx, const uint8_t id[128]);
/* Optional, after demcz_set_archive + demcz_comm_init on every rank: exchange the 128-byte peer records
 * (all-gather on the host), then open them.  From then on a rank's appended rows are pushed by our own kernel
 * straight into every peer's archive replica over NVLink (CUDA IPC / peer access) instead of ncclAllGather.
 * infos: world x 128 bytes, in rank order. */
int demcz_comm_peer_info(demcz_ctx* ctx, uint8_t info[128]);
int demcz_comm_open_peers(demcz_ctx* ctx, const uint8_t* infos);

#ifdef __cplusplus
}
#endif
#endif
