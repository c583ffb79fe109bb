/*
 * demcz_oracle.h — CPU restatement of the DEMCz generation loop of chrished/DEMC.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / the timed CPU baseline, never as a fallback.
 *
 * PARITY UNPINNED (w.r.t. the reference): the reference is pure Julia (no Julia toolchain in
 * this image), its tests hold no golden vectors or known-answer values for this path (only
 * statistical bands: test/example_normpdf.jl:49-51), and its RNG (MersenneTwister, Julia 1.1)
 * is replaced by a documented Philox4x32-10 stream.  What IS pinned: Philox against the
 * Random123 known-answer vectors, the MvNormal log-density against scipy, the linreg
 * log-likelihood against a FP64 Gram form, split-R-hat against a numpy transcription of
 * src/utils.jl:2-20, and the C and numpy restatements against each other
 * (tests/test_oracle_*.py, tests/golden/).
 *
 * All reference citations are file:line into /root/reference (read-only upstream).
 */
#ifndef DEMCZ_ORACLE_H
#define DEMCZ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Philox4x32-10 (Salmon et al. 2011, Random123) -------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Counter layout shared with the CUDA path (DESIGN.md "RNG stream"):
 *   key = (seed_lo, seed_hi)
 *   ctr = (global chain id, gen_abs lo32, (block << 16) | call, gen_abs hi32)
 *   call 0      -> words (w_r1, w_r2, w_accept, spare)
 *   call 1 + q  -> 4 words -> 4 Box-Muller pairs (16-bit uniforms) -> jitter normals of columns 8q..8q+7
 * uniform(w)  = ((w >> 9) + 0.5) * 2^-23          (exact in FP32 and FP64, in (0,1))
 * r1 = mulhi32(w_r1, M); j = mulhi32(w_r2, M-1); r2 = j + (j >= r1)   (0-based; Q9)
 */
double orc_uniform(uint32_t w);
void orc_draw_pair(uint32_t w_r1, uint32_t w_r2, uint32_t M, uint32_t* r1, uint32_t* r2);
void orc_box_muller16(uint32_t w, double* z0, double* z1);

/* ---- likelihood descriptors (the reference passes a closure: src/demcz.jl:189) ------------- */
enum { ORC_LIK_MVNORMAL = 0, ORC_LIK_LINREG = 1, ORC_LIK_LOGISTIC = 2 };   /* 2: extension, sum(y*eta - log(1+exp(eta))) */

typedef struct {
    int kind;
    int d;
    /* MvNormal (test/example_normpdf.jl:9-16): mu[d], L = lower Cholesky factor of Sigma
     * (col-major d x d), c0 = -0.5*(d*log(2pi) + logdet Sigma)  (0 if the constant is dropped).
     * The quadratic objective -sum((x-mu)^2) of test/test_anneal.jl:10 is MvNormal(mu, I/2)
     * with c0 = 0. */
    const double* mu;
    const double* L;
    double c0;
    /* Linreg (test/example_linreg.jl:32): -0.5 * sum((y - X b)^2), X col-major n x d */
    int64_t n;
    const double* X;
    const double* y;
} orc_lik;

/* Sigma (col-major d x d) -> lower Cholesky L (col-major) and logdet; returns 0, or -1 if not PD */
int orc_cholesky(int d, const double* Sigma, double* L, double* logdet);
double orc_logdensity(const orc_lik* lik, const double* x);

/* ---- options (mirror of DEMCopt, src/DEMC.jl:24-39) --------------------------------------- */
enum { ORC_APPEND_GEN_SYNC = 0, ORC_APPEND_REF_INLOOP = 1 };   /* SURVEY Q2 */
enum { ORC_INDEX_O1 = 0, ORC_INDEX_FAITHFUL = 1 };              /* src/demcz.jl:176-179 cost model */

typedef struct {
    int N, K, Nblocks, d;
    const int* block_ptr;     /* CSR, Nblocks+1 */
    const int* block_idx;     /* 0-based column indices */
    const double* eps_scale;  /* d */
    double gamma;
    uint64_t seed;
    int64_t gen_offset;       /* absolute generation = gen_offset + ig, ig = 1.. */
    int64_t chain_offset;     /* global id of chain 0 (multi-rank sharding) */
    int append_mode;
    int index_mode;
    int nthreads;             /* chains split over threads inside a generation (gen-sync only) */
    int z_rowmajor;           /* 0: Z col-major like Julia (Zld = row capacity); 1: contiguous rows (stride d) — 'optimised CPU' mode B */
} orc_opts;

/* Per-proposal trace of one generation (all optional, may be NULL). Layout [block][chain]. */
typedef struct {
    uint32_t* r1;       /* Nblocks*N */
    uint32_t* r2;
    double* u;          /* accept uniform */
    double* xprop;      /* [block][k][chain] : Nblocks*d*N */
    double* lp_prop;
    uint8_t* accepted;
} orc_trace;

/* One generation for all chains (update_blocks / update_demcz_chain_block / accept:
 * src/demcz.jl:167-203; tempered: src/demcz_anneal.jl:142-178).
 *   X  : N x d col-major (chain fastest), updated in place     lp : N, updated in place
 *   Z  : archive, Zld x d col-major, M rows valid (frozen during the generation in gen-sync mode)
 *   gen_abs : absolute generation number (Philox counter)
 *   T  : temperature; pass tempered=0 for the plain accept of src/demcz.jl:197-203
 * In ORC_APPEND_REF_INLOOP mode and when append_now != 0, chain ic's row is appended right after
 * its update (src/demcz.jl:88-91 inside :31-33) so later chains of the same generation see it;
 * *M is advanced.  In gen-sync mode the caller appends after the generation. */
void orc_generation(const orc_opts* o, const orc_lik* lik, double* X, double* lp,
                    double* Z, int64_t Zld, int64_t* M, int64_t gen_abs,
                    int tempered, double T, int append_now, const orc_trace* tr);

/* Result / driver mirroring demcz_sample (src/demcz.jl:9-63) and demcz_anneal
 * (src/demcz_anneal.jl:18-65). */
typedef struct {
    int Ngeneration;
    int anneal;               /* 0: demcz_sample, 1: demcz_anneal */
    const double* Ttable;     /* anneal: length Ngeneration, T for ig = 1..Ngeneration */
    int autostop;             /* sample only: 1 = :Rhat */
    int autostop_every;
    double autostop_Rhat;
    int adapt;                /* anneal: adaptγ (src/demcz_anneal.jl:48-57) */
    double min_gamma, max_gamma;
    int adapt_every;
    int init_zeros;           /* 1 = reference serial quirk (start at zeros) */
    int store_chain;
} orc_run_opts;

typedef struct {
    int generations_done;
    int stopped_by_rhat;
    double final_gamma;
    int64_t M;
} orc_run_info;

/* Z: capacity Zld rows (>= M0 + ceil(N*G/K)), M0 valid; X0/lp0 NULL -> init from archive.
 * chain: N x d x G (col-major) or NULL; log_obj: N x G or NULL; Xout N x d; lpout N. */
int orc_run(const orc_opts* o, const orc_run_opts* r, const orc_lik* lik,
            double* Z, int64_t Zld, int64_t M0, const double* X0, const double* lp0,
            double* chain, double* log_obj, double* Xout, double* lpout, double* rhat_out,
            orc_run_info* info);

/* Snooker update of ter Braak & Vrugt (2008) for every chain (one proposal per chain, all coordinates) — an extension:
 * the reference cites the paper (README.md:57-59) and implements the parallel update only.  Trace layout as one block. */
int orc_is_snooker_gen(uint64_t seed, int64_t gen_abs, double prob);
void orc_snooker_generation(const orc_opts* o, const orc_lik* lik, double* X, double* lp, const double* Z, int64_t Zld,
                            int64_t M, int64_t gen_abs, double gamma_s, int tempered, double T, const orc_trace* tr);

/* Rhat_gelman, src/utils.jl:2-20.  chain: N x d x G col-major; out: d */
void orc_rhat_gelman(const double* chain, int N, int G, int d, double* out);
/* tempbaseline, src/demcz_anneal.jl:1-3 */
double orc_tempbaseline(double ig, double Ng, double T0, double TN);

#ifdef __cplusplus
}
#endif
#endif
