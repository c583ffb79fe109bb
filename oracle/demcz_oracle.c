/*
 * demcz_oracle.c — CPU restatement (FP64) of the DEMCz hot path of chrished/DEMC.jl.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED w.r.t. the Julia reference — see demcz_oracle.h.
 * Every function cites the reference file:line (under /root/reference) it follows.
 */
#include "demcz_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10.  Not in the reference (it uses Julia's global MersenneTwister: src/demcz.jl:177,
 * 179,184,186,198); this is the published Random123 algorithm, pinned by its KAT vectors
 * (tests/test_oracle_philox.py).
 * ------------------------------------------------------------------------------------------- */
#define ZAT(Z, Zld, d, rm, r, k) ((Z)[(rm) ? (size_t)(r) * (size_t)(d) + (size_t)(k) : (size_t)(r) + (size_t)(Zld) * (size_t)(k)])
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uniform in (0,1), exactly representable in FP32: replaces rand() of src/demcz.jl:198 */
double orc_uniform(uint32_t w) { return ((double)(w >> 9) + 0.5) * (1.0 / 8388608.0); }

/* i1 = rand(1:M); deleteat!(set, i1); i2 = rand(set)  — src/demcz.jl:176-179.
 * 0-based: r1 uniform on [0,M), j uniform on [0,M-1), r2 = j + (j >= r1). */
void orc_draw_pair(uint32_t w_r1, uint32_t w_r2, uint32_t M, uint32_t* r1, uint32_t* r2)
{
    uint32_t a = (uint32_t)(((uint64_t)w_r1 * (uint64_t)M) >> 32);
    uint32_t j = (uint32_t)(((uint64_t)w_r2 * (uint64_t)(M - 1)) >> 32);
    *r1 = a;
    *r2 = j + (j >= a ? 1u : 0u);
}

/* randn() of src/demcz.jl:184,186 replaced by Box-Muller on ONE Philox word: 16-bit uniforms from the
 * high / low half-word, u = (h + 0.5) 2^-16 (exact in FP32, in (0,1)) */
void orc_box_muller16(uint32_t w, double* z0, double* z1)
{
    double u1 = ((double)(w >> 16) + 0.5) * (1.0 / 65536.0), u2 = ((double)(w & 0xFFFFu) + 0.5) * (1.0 / 65536.0);
    double r = sqrt(-2.0 * log(u1));
    double th = 6.283185307179586476925286766559 * u2;
    *z0 = r * cos(th);
    *z1 = r * sin(th);
}

/* ---------------------------------------------------------------------------------------------
 * Likelihoods
 * ------------------------------------------------------------------------------------------- */
int orc_cholesky(int d, const double* S, double* L, double* logdet)
{
    memset(L, 0, sizeof(double) * (size_t)d * d);
    double ld = 0.0;
    for (int j = 0; j < d; ++j) {
        double s = S[j + (size_t)d * j];
        for (int k = 0; k < j; ++k) s -= L[j + (size_t)d * k] * L[j + (size_t)d * k];
        if (!(s > 0.0)) return -1;
        double ljj = sqrt(s);
        L[j + (size_t)d * j] = ljj;
        ld += 2.0 * log(ljj);
        for (int i = j + 1; i < d; ++i) {
            double t = S[i + (size_t)d * j];
            for (int k = 0; k < j; ++k) t -= L[i + (size_t)d * k] * L[j + (size_t)d * k];
            L[i + (size_t)d * j] = t / ljj;
        }
    }
    if (logdet) *logdet = ld;
    return 0;
}

/* MvNormal: logpdf(MvNormal(mu,Sigma), x) as called at test/example_normpdf.jl:16 (Distributions.jl
 * 0.21.3 / PDMats 0.9.10, not vendored: textbook -0.5*(d log 2pi + logdet) - 0.5*|L^-1 (x-mu)|^2).
 * Linreg: -0.5*sum((y .- X*b).^2), test/example_linreg.jl:32.
 * Quadratic -sum((x .- mu).^2), test/test_anneal.jl:10 == MvNormal with L = I/sqrt(2), c0 = 0. */
double orc_logdensity(const orc_lik* lik, const double* x)
{
    const int d = lik->d;
    if (lik->kind == ORC_LIK_MVNORMAL) {
        double ybuf[64];
        double* y = d <= 64 ? ybuf : (double*)malloc(sizeof(double) * d);
        double q = 0.0;
        for (int i = 0; i < d; ++i) {
            double t = x[i] - lik->mu[i];
            for (int k = 0; k < i; ++k) t -= lik->L[i + (size_t)d * k] * y[k];
            y[i] = t / lik->L[i + (size_t)d * i];
            q += y[i] * y[i];
        }
        if (y != ybuf) free(y);
        return lik->c0 - 0.5 * q;
    } else if (lik->kind == ORC_LIK_LOGISTIC) {
        /* Bernoulli-logit GLM, sum(y*eta - log(1 + exp(eta))), eta = X b: an extension (the reference takes any closure,
         * src/demcz.jl:189); textbook formula with the overflow-safe softplus max(eta,0) + log1p(exp(-|eta|)) */
        const int64_t n = lik->n;
        double* eta = (double*)calloc((size_t)n, sizeof(double));
        for (int k = 0; k < d; ++k) {
            const double bk = x[k];
            const double* col = lik->X + (size_t)n * k;
            for (int64_t i = 0; i < n; ++i) eta[i] += col[i] * bk;
        }
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += lik->y[i] * eta[i] - (fmax(eta[i], 0.0) + log1p(exp(-fabs(eta[i]))));
        free(eta);
        return s;
    } else {
        const int64_t n = lik->n;
        double* r = (double*)malloc(sizeof(double) * (size_t)n);   /* y .- X*b temporary */
        memcpy(r, lik->y, sizeof(double) * (size_t)n);
        for (int k = 0; k < d; ++k) {
            const double bk = x[k];
            const double* col = lik->X + (size_t)n * k;
            for (int64_t i = 0; i < n; ++i) r[i] -= col[i] * bk;
        }
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += r[i] * r[i];
        free(r);
        return -0.5 * s;
    }
}

/* ---------------------------------------------------------------------------------------------
 * One chain, one generation: update_blocks (src/demcz.jl:167-172; tempered src/demcz_anneal.jl:
 * 142-147) -> update_demcz_chain_block (src/demcz.jl:174-195; :149-170) -> accept (:197-203;
 * tempered :172-178).
 * ------------------------------------------------------------------------------------------- */
static void chain_generation(const orc_opts* o, const orc_lik* lik, int ic, double* X, double* lp,
                             const double* Z, int64_t Zld, int64_t M, int64_t gen_abs,
                             int tempered, double T, const orc_trace* tr)
{
    const int d = o->d, N = o->N;
    double xbuf[64], pbuf[64];
    double* x = d <= 64 ? xbuf : (double*)malloc(sizeof(double) * d);
    double* xp = d <= 64 ? pbuf : (double*)malloc(sizeof(double) * d);
    for (int k = 0; k < d; ++k) x[k] = X[ic + (size_t)N * k];     /* mc.Xcurrent[ic, :]  :82 */
    double cur = lp[ic];
    const uint32_t key[2] = { (uint32_t)o->seed, (uint32_t)(o->seed >> 32) };
    const uint32_t cid = (uint32_t)(o->chain_offset + ic);

    for (int ib = 0; ib < o->Nblocks; ++ib) {                     /* for ib in 1:Nblocks  :168 */
        uint32_t ctr[4] = { cid, (uint32_t)gen_abs, ((uint32_t)ib << 16), (uint32_t)((uint64_t)gen_abs >> 32) };
        uint32_t w[4];
        orc_philox4x32_10(ctr, key, w);
        uint32_t r1, r2;
        if (o->index_mode == ORC_INDEX_FAITHFUL) {
            /* set = collect(1:M); i1 = rand(set); deleteat!(set,i1); i2 = rand(set)  :176-179
             * (same draws as the O(1) form, with the reference's O(M) allocation cost) */
            int64_t* set = (int64_t*)malloc(sizeof(int64_t) * (size_t)M);
            for (int64_t i = 0; i < M; ++i) set[i] = i;
            uint32_t a = (uint32_t)(((uint64_t)w[0] * (uint64_t)M) >> 32);
            r1 = (uint32_t)set[a];
            memmove(set + a, set + a + 1, sizeof(int64_t) * (size_t)(M - 1 - a));
            uint32_t j = (uint32_t)(((uint64_t)w[1] * (uint64_t)(M - 1)) >> 32);
            r2 = (uint32_t)set[j];
            free(set);
        } else {
            orc_draw_pair(w[0], w[1], (uint32_t)M, &r1, &r2);
        }
        const int b0 = o->block_ptr[ib], b1 = o->block_ptr[ib + 1];
        const int blocklen = b1 - b0;
        /* γ*(…) if blocklen == 1 else γ/sqrt(2*blocklen)*(…)  :183-187 */
        const double c = blocklen == 1 ? o->gamma : o->gamma / sqrt(2.0 * blocklen);
        for (int k = 0; k < d; ++k) xp[k] = x[k];                 /* de_diffvec = zeros  :180 */
        int have_q = -1; double zn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int t = b0; t < b1; ++t) {
            const int k = o->block_idx[t];
            const int q = k >> 3;
            if (q != have_q) {                                    /* randn(blocklen)  :186 */
                uint32_t cn[4] = { cid, (uint32_t)gen_abs, ((uint32_t)ib << 16) | (uint32_t)(1 + q),
                                   (uint32_t)((uint64_t)gen_abs >> 32) };
                uint32_t wn[4];
                orc_philox4x32_10(cn, key, wn);
                for (int j = 0; j < 4; ++j) orc_box_muller16(wn[j], &zn[2 * j], &zn[2 * j + 1]);   /* columns 8q..8q+7 */
                have_q = q;
            }
            const double diff = c * (ZAT(Z, Zld, d, o->z_rowmajor, r1, k) - ZAT(Z, Zld, d, o->z_rowmajor, r2, k))
                              + o->eps_scale[k] * zn[k & 7];
            xp[k] = x[k] + diff;                                  /* Xproposal = Xcurrent .+ de_diffvec :188 */
        }
        const double lpp = orc_logdensity(lik, xp);              /* logobj(Xproposal)  :189 */
        const double u = orc_uniform(w[2]);
        int acc;
        if (tempered) acc = log(u) < (lpp - cur) / T;             /* src/demcz_anneal.jl:173 (T=0 -> greedy, Q4) */
        else          acc = log(u) < lpp - cur;                   /* src/demcz.jl:198 */
        if (tr) {
            const size_t o1 = (size_t)ib * N + ic;
            if (tr->r1) tr->r1[o1] = r1;
            if (tr->r2) tr->r2[o1] = r2;
            if (tr->u) tr->u[o1] = u;
            if (tr->lp_prop) tr->lp_prop[o1] = lpp;
            if (tr->accepted) tr->accepted[o1] = (uint8_t)acc;
            if (tr->xprop) for (int k = 0; k < d; ++k) tr->xprop[((size_t)ib * d + k) * N + ic] = xp[k];
        }
        if (acc) { for (int k = 0; k < d; ++k) x[k] = xp[k]; cur = lpp; }   /* :190-194 */
    }
    for (int k = 0; k < d; ++k) X[ic + (size_t)N * k] = x[k];     /* mc.Xcurrent[ic,:] .= …  :86 */
    lp[ic] = cur;                                                 /* :87 */
    if (x != xbuf) free(x);
    if (xp != pbuf) free(xp);
}

void orc_generation(const orc_opts* o, const orc_lik* lik, double* X, double* lp,
                    double* Z, int64_t Zld, int64_t* M, int64_t gen_abs,
                    int tempered, double T, int append_now, const orc_trace* tr)
{
    const int N = o->N, d = o->d;
    if (o->append_mode == ORC_APPEND_REF_INLOOP) {
        /* for ic = 1:N runchain!(…) with the append inside — src/demcz.jl:31-33, 88-91 (Q2) */
        for (int ic = 0; ic < N; ++ic) {
            chain_generation(o, lik, ic, X, lp, Z, Zld, *M, gen_abs, tempered, T, tr);
            if (append_now) {
                for (int k = 0; k < d; ++k) ZAT(Z, Zld, d, o->z_rowmajor, *M, k) = X[ic + (size_t)N * k];
                *M += 1;
            }
        }
        return;
    }
    /* generation-synchronous: Z frozen, chains independent (the semantics of the CUDA path;
     * mirrors pmap over chains in demcz_sample_par, src/demcz.jl:137) */
    const int64_t Mg = *M;
    int nt = o->nthreads > 1 ? o->nthreads : 1;
    (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
#endif
    for (int ic = 0; ic < N; ++ic)
        chain_generation(o, lik, ic, X, lp, Z, Zld, Mg, gen_abs, tempered, T, tr);
}

/* ---------------------------------------------------------------------------------------------
 * Rhat_gelman — src/utils.jl:2-20
 * ------------------------------------------------------------------------------------------- */
void orc_rhat_gelman(const double* chain, int N, int G, int d, double* out)
{
    const int n = G / 2;                 /* n = Int(floor(Ngeneration/2))  :4 */
    const int m = 2 * N;                 /* m = Npop*2  :5 */
    double* avg_chains = (double*)malloc(sizeof(double) * (size_t)m);
    for (int k = 0; k < d; ++k) {
        double tot = 0.0;
        for (int j = 0; j < m; ++j) {    /* chainsplit[1:Npop] = chain[:,:,1:n]; [Npop+1:end] = chain[:,:,n+1:2n]  :7-8 */
            const int ic = j % N, half = j / N;
            double s = 0.0;
            for (int t = 0; t < n; ++t) s += chain[ic + (size_t)N * (k + (size_t)d * (half * n + t))];
            avg_chains[j] = s / n;       /* :11 */
            tot += s;
        }
        const double avg_par = tot / ((double)m * n);            /* :10 */
        double B = 0.0, W = 0.0;
        for (int j = 0; j < m; ++j) {
            const int ic = j % N, half = j / N;
            B += (avg_chains[j] - avg_par) * (avg_chains[j] - avg_par);
            double sj = 0.0;
            for (int t = 0; t < n; ++t) {
                const double v = chain[ic + (size_t)N * (k + (size_t)d * (half * n + t))] - avg_chains[j];
                sj += v * v;
            }
            W += sj / (n - 1);           /* sj = 1/(n-1)*sum(…)  :14 */
        }
        B *= (double)n / (m - 1);        /* :13 */
        W /= m;                          /* :15 */
        const double varhat = (double)(n - 1) / n * W + B / n;   /* :16 */
        out[k] = sqrt(varhat / W);       /* :18 */
    }
    free(avg_chains);
}

/* tempbaseline — src/demcz_anneal.jl:1-3 */
double orc_tempbaseline(double ig, double Ng, double T0, double TN) { return T0 * pow(TN / T0, ig / Ng); }

/* ---------------------------------------------------------------------------------------------
 * Drivers: demcz_sample (src/demcz.jl:9-63) / demcz_anneal (src/demcz_anneal.jl:18-65)
 * ------------------------------------------------------------------------------------------- */
int orc_run(const orc_opts* o, const orc_run_opts* r, const orc_lik* lik,
            double* Z, int64_t Zld, int64_t M0, const double* X0, const double* lp0,
            double* chain, double* log_obj, double* Xout, double* lpout, double* rhat_out,
            orc_run_info* info)
{
    const int N = o->N, d = o->d, G = r->Ngeneration, K = o->K;
    orc_opts oo = *o;
    int64_t M = M0;                                               /* M = ones(Int64,1)*nrowZ  :12 */
    double* X = Xout;
    double* lp = lpout;
    if (X0) {                                                     /* prevrun  :20-21 */
        memcpy(X, X0, sizeof(double) * (size_t)N * d);
        if (lp0) memcpy(lp, lp0, sizeof(double) * (size_t)N);
    } else if (r->init_zeros) {                                   /* Zmat[end-N+1:end,:] after vcat zeros: Q1, :11,15 */
        memset(X, 0, sizeof(double) * (size_t)N * d);
    } else {                                                      /* last N rows of Z (documented intent :14; par :113) */
        for (int ic = 0; ic < N; ++ic)
            for (int k = 0; k < d; ++k) X[ic + (size_t)N * k] = ZAT(Z, Zld, d, o->z_rowmajor, M0 - N + ic, k);
    }
    if (!(X0 && lp0)) {                                           /* map(logobj, rows)  :17 */
        double xb[64]; double* x = d <= 64 ? xb : (double*)malloc(sizeof(double) * d);
        for (int ic = 0; ic < N; ++ic) {
            for (int k = 0; k < d; ++k) x[k] = X[ic + (size_t)N * k];
            lp[ic] = orc_logdensity(lik, x);
        }
        if (x != xb) free(x);
    }
    int64_t changed_window = 0;                                   /* count(diff(log_obj window).!=0) */
    double* lp_prev = (double*)malloc(sizeof(double) * (size_t)N);
    int done = 0, stopped = 0;
    for (int ig = 1; ig <= G; ++ig) {                             /* for ig = 1:Ngeneration  :30 */
        memcpy(lp_prev, lp, sizeof(double) * (size_t)N);
        const int append_now = (ig % K) == 0;                     /* mod(ig,K) == 0  :88 */
        const double T = r->anneal ? r->Ttable[ig - 1] : 1.0;     /* temperaturefun(ig, …)  anneal:69 */
        orc_generation(&oo, lik, X, lp, Z, Zld, &M, oo.gen_offset + ig, r->anneal, T, append_now, NULL);
        if (append_now && oo.append_mode == ORC_APPEND_GEN_SYNC) {
            for (int ic = 0; ic < N; ++ic) {
                for (int k = 0; k < d; ++k) ZAT(Z, Zld, d, oo.z_rowmajor, M, k) = X[ic + (size_t)N * k];
                M += 1;
            }
        }
        if (chain) memcpy(chain + (size_t)N * d * (ig - 1), X, sizeof(double) * (size_t)N * d);   /* :84 */
        if (log_obj) memcpy(log_obj + (size_t)N * (ig - 1), lp, sizeof(double) * (size_t)N);     /* :85 */
        done = ig;
        if (r->anneal && r->adapt) {                              /* src/demcz_anneal.jl:48-57 */
            /* diff over the window's columns has adapt_every-1 entries per chain: the first
             * generation of each window is not compared with its predecessor */
            if ((ig - 1) % r->adapt_every != 0)
                for (int ic = 0; ic < N; ++ic) changed_window += (lp[ic] != lp_prev[ic]);
            if (ig % r->adapt_every == 0) {
                const double ar = (double)changed_window / ((double)N * r->adapt_every);
                if (ar < 0.1) oo.gamma = fmax(r->min_gamma, oo.gamma * 0.5);
                else if (ar > 0.5) oo.gamma = fmin(r->max_gamma, oo.gamma * 1.5);
                changed_window = 0;
            }
        }
        if (!r->anneal && r->autostop && chain && ig % r->autostop_every == 0) {   /* :39-55 */
            orc_rhat_gelman(chain + (size_t)N * d * (ig - r->autostop_every), N, r->autostop_every, d, rhat_out);
            double mx = 0.0;
            for (int k = 0; k < d; ++k) mx = rhat_out[k] > mx || rhat_out[k] != rhat_out[k] ? rhat_out[k] : mx;
            if (mx < r->autostop_Rhat) { stopped = 1; break; }
        }
    }
    free(lp_prev);
    if (info) { info->generations_done = done; info->stopped_by_rhat = stopped; info->final_gamma = oo.gamma; info->M = M; }
    return 0;
}


/* ---------------------------------------------------------------------------------------------
 * Snooker update of ter Braak & Vrugt (2008) — cited by the reference (README.md:57-59) but not implemented there;
 * restated from the paper, same Philox stream layout as the CUDA path (csrc/philox.cuh):
 *   generation-level choice: ctr = (0xFFFFFFFF, gen lo, 0xFFFF0000, gen hi), word 0 -> u < prob
 *   per chain: call 0xFFFF0000 -> (w_z, w_z1, w_z2, w_accept); three distinct rows by successive removal;
 *              call 0xFFFF0001 word 0 -> gamma_s = 1.2 + u when gamma_s == 0
 *   x' = x + gamma_s e ((z1 - z2).e)/(e.e), e = x - z; accept with pi(x')/pi(x) |1 + gamma_s proj|^(d-1)
 * ------------------------------------------------------------------------------------------- */
int orc_is_snooker_gen(uint64_t seed, int64_t gen_abs, double prob)
{
    if (!(prob > 0.0)) return 0;
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    const uint32_t ctr[4] = { 0xFFFFFFFFu, (uint32_t)gen_abs, 0xFFFF0000u, (uint32_t)((uint64_t)gen_abs >> 32) };
    uint32_t w[4];
    orc_philox4x32_10(ctr, key, w);
    return orc_uniform(w[0]) < prob;
}

void orc_snooker_generation(const orc_opts* o, const orc_lik* lik, double* X, double* lp, const double* Z, int64_t Zld,
                            int64_t M, int64_t gen_abs, double gamma_s, int tempered, double T, const orc_trace* tr)
{
    const int d = o->d, N = o->N;
    const uint32_t key[2] = { (uint32_t)o->seed, (uint32_t)(o->seed >> 32) };
    for (int ic = 0; ic < N; ++ic) {
        uint32_t ctr[4] = { (uint32_t)(o->chain_offset + ic), (uint32_t)gen_abs, 0xFFFF0000u, (uint32_t)((uint64_t)gen_abs >> 32) };
        uint32_t w[4];
        orc_philox4x32_10(ctr, key, w);
        uint32_t z = (uint32_t)(((uint64_t)w[0] * (uint64_t)M) >> 32);
        uint32_t j1 = (uint32_t)(((uint64_t)w[1] * (uint64_t)(M - 1)) >> 32);
        uint32_t z1 = j1 + (j1 >= z ? 1u : 0u);
        uint32_t z2 = (uint32_t)(((uint64_t)w[2] * (uint64_t)(M - 2)) >> 32);
        const uint32_t lo = z < z1 ? z : z1, hi = z < z1 ? z1 : z;
        if (z2 >= lo) ++z2;
        if (z2 >= hi) ++z2;
        double gs = (double)(float)gamma_s;
        if (gamma_s == 0.0) {
            uint32_t w2[4];
            ctr[2] = 0xFFFF0001u;
            orc_philox4x32_10(ctr, key, w2);
            gs = (double)(float)(1.2f + (float)orc_uniform(w2[0]));            /* the device forms 1.2f + u in FP32 */
        }
        double x[64], xp[64], ee = 0.0, num = 0.0;
        for (int k = 0; k < d; ++k) {
            x[k] = X[ic + (size_t)N * k];
            const double e = x[k] - ZAT(Z, Zld, d, o->z_rowmajor, z, k);
            ee += e * e;
            num += (ZAT(Z, Zld, d, o->z_rowmajor, z1, k) - ZAT(Z, Zld, d, o->z_rowmajor, z2, k)) * e;
        }
        const double proj = ee > 0.0 ? gs * num / ee : 0.0;
        for (int k = 0; k < d; ++k) xp[k] = x[k] + proj * (x[k] - ZAT(Z, Zld, d, o->z_rowmajor, z, k));
        const double lpp = orc_logdensity(lik, xp);
        const double geo = (d - 1) * log(fabs(1.0 + proj));
        const double u = orc_uniform(w[3]);
        const double Tm = tempered ? T : 1.0;
        const int acc = ee > 0.0 && (Tm * log(u) < (lpp - lp[ic]) + Tm * geo);
        if (tr) {
            if (tr->r1) tr->r1[ic] = z1;
            if (tr->r2) tr->r2[ic] = z2;
            if (tr->u) tr->u[ic] = u;
            if (tr->lp_prop) tr->lp_prop[ic] = lpp;
            if (tr->accepted) tr->accepted[ic] = (uint8_t)acc;
            if (tr->xprop) for (int k = 0; k < d; ++k) tr->xprop[(size_t)k * N + ic] = xp[k];
        }
        if (acc) {
            for (int k = 0; k < d; ++k) X[ic + (size_t)N * k] = xp[k];
            lp[ic] = lpp;
        }
    }
}
