// demcz_step.cuh — the fused DEMCz generation kernel for closed-form (register-resident)
// likelihoods: MvNormal logpdf and the quadratic annealing objective.
//
// One launch advances every local chain by `n_gens` generations against a FROZEN archive Z
// (generation-synchronous semantics), i.e. everything runchain! does between two
// archive appends (reference src/demcz.jl:80-93 -> :167-203; tempered src/demcz_anneal.jl:67-80,
// 142-178):
//   for each generation, for each block:
//     Philox draw of r1 != r2 and the accept uniform            (src/demcz.jl:176-179, 198)
//     gather Z[r1,:], Z[r2,:] from the HBM archive                (:184/186)
//     x' = x + c_b (Z[r1]-Z[r2]) + eps .* N(0,1) on the block     (:183-188)
//     lp' = c0 - 0.5 |W (x'-mu)|^2                                 (logobj, :189)
//     accept iff log(u) < (lp'-lp)/T                              (:197-203 / anneal :172-178)
//   history / counters / split-Rhat accumulators                 (:84-87, utils.jl:2-20)
//   optional append of the final state as archive rows           (:88-91)
//
// Mapping: thread = chain (state in registers for the whole launch), warp = 32 chains.  The two
// archive rows of the NEXT proposal are fetched while the current one is evaluated: indices are
// counter-based, so they do not depend on the accept outcome.  The gather is cooperative: the
// warp's 64 rows are cut into 16-byte chunks and consecutive lanes copy consecutive chunks of the
// same row with cp.async (LDGSTS, L2-only), so one warp instruction touches ~32/C rows instead
// of 32 — the L1TEX wavefront count, not DRAM, is what a lane-per-row gather would saturate.
// Rows are staged in shared memory (stride chosen conflict-free for LDS.128) and read back by
// the owning lane.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.cuh"

namespace demcz {

constexpr int kMaxBlocks = 32;

struct BlockDesc {
    uint32_t mask;     // columns of this block (bit k = column k)
    float cscale;      // γ/sqrt(2*blocklen), or γ if blocklen == 1   (src/demcz.jl:183-187)
};

struct StepArgs {
    // chain state, SoA: X[k*ldx + chain]  (== Julia's column-major N x d Xcurrent, in FP32)
    float* X;
    float* logp;
    int N, ldx, d;
    uint32_t chain_offset;       // global id of local chain 0
    // archive: row-major, row stride LD floats (16-byte aligned rows)
    const float* Z;
    uint32_t Mdraw;              // rows eligible for the draw
    uint32_t zstart, zcap;       // ring: physical row = (zstart + logical) mod zcap
    // generations of this launch
    long long gen0;              // absolute number (Philox counter) of the first generation
    long long grel0;             // 1-based number inside this context of the first generation
    int n_gens;
    uint32_t key0, key1;
    int nblocks;
    BlockDesc blocks[kMaxBlocks];
    float eps[32];
    const float* Ttab;           // temperatures, indexed by grel-1; nullptr -> plain accept
    // append (only when the last generation of the launch is a multiple of K)
    float* Zw;                   // archive base for the append, nullptr -> no append
    uint32_t append_row0;        // physical row of global chain 0's new row
    // history (store_chain): hist[(grel-1)*d + k][chain], lphist[grel-1][chain]
    float* hist;
    float* lphist;
    // trace (tests)
    uint32_t* tr_r1; uint32_t* tr_r2; float* tr_u; float* tr_xprop; float* tr_lpprop; uint8_t* tr_acc;
    // counters
    int* chain_changed;                   // per chain, generations whose lp changed (skipping run_skip_gi)
    unsigned long long* win_changed;      // adaptγ window counter (skipping win_skip_gi)
    unsigned long long* accepted_total;
    int run_skip_gi, win_skip_gi;         // index inside the launch of the generation not counted, or -1
    // split-Rhat accumulators of the current half-window
    float* S1;                   // [d][ldx]  sum_t (x - shift)
    double* S2;                  // [d]       sum_chains sum_t (x - shift)^2
    const float* shift;          // [d]
};

template <int LD>
struct MvnParams {
    float mu[LD];
    float W[LD * (LD + 1) / 2];   // packed lower triangle (row-major) of L^-1, Sigma = L L^T
    float c0;
};

template <int LD>
struct StageLayout {
    static constexpr int C = LD / 4;                              // 16-byte chunks per row
    static constexpr int LDS = (C % 2 == 1) ? LD : LD + 4;        // staged row stride (floats): odd #chunks
    static constexpr int ROWS = 64 * LDS;                         // 32 x r1 rows, 32 x r2 rows
    static constexpr int FLOATS = ROWS + 96;                      // + idx[64] + accept word[32]
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log-density of the closed-form family: c0 - 0.5 |W (x - mu)|^2.  W, mu are kernel parameters
// (constant bank), indices are compile-time -> every FFMA takes its matrix entry as a constant
// operand, no loads.
template <int LD>
__device__ __forceinline__ float mvn_logdensity(const MvnParams<LD>& P, const float (&x)[LD])
{
    float dx[LD];
#pragma unroll
    for (int k = 0; k < LD; ++k) dx[k] = x[k] - P.mu[k];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LD; ++i) {
        float y = 0.f;
#pragma unroll
        for (int j = 0; j <= i; ++j) y = fmaf(P.W[i * (i + 1) / 2 + j], dx[j], y);
        q = fmaf(y, y, q);
    }
    return fmaf(-0.5f, q, P.c0);
}

template <int LD, bool ACC, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32)
demcz_step_mvn_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ MvnParams<LD> P)
{
    using SL = StageLayout<LD>;
    constexpr int C = SL::C, LDS = SL::LDS;
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned int red_cnt[2];
    __shared__ float red_s2[ACC ? LD : 1];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + warp * (STAGES * SL::FLOATS);
    const int chain = blockIdx.x * (WARPS * 32) + threadIdx.x;
    const bool active = chain < a.N;
    const int cl = active ? chain : a.N - 1;          // clamped: idle lanes shadow the last chain, never store
    const uint32_t cid = a.chain_offset + (uint32_t)cl;

    if (threadIdx.x < 2) red_cnt[threadIdx.x] = 0u;
    if (ACC && threadIdx.x < LD) red_s2[threadIdx.x] = 0.f;

    float x[LD];
#pragma unroll
    for (int k = 0; k < LD; ++k) x[k] = (k < a.d) ? a.X[(size_t)k * a.ldx + cl] : 0.f;
    float lp = a.logp[cl];

    float s1[ACC ? LD : 1], s2[ACC ? LD : 1], sh[ACC ? LD : 1];
    if (ACC) {
#pragma unroll
        for (int k = 0; k < LD; ++k) {
            sh[k] = (k < a.d) ? a.shift[k] : 0.f;
            s1[k] = (k < a.d) ? a.S1[(size_t)k * a.ldx + cl] : 0.f;
            s2[k] = 0.f;
        }
    }

    const int nb = a.nblocks;
    const int Q = a.n_gens * nb;
    int cnt_run = 0, cnt_win = 0, cnt_acc = 0;

    // ---- producer side of the per-warp pipeline -------------------------------------------
    int pg = 0, pb = 0;
    auto prefetch = [&](int stage) {
        float* st = wbase + stage * SL::FLOATS;
        uint32_t* idx = reinterpret_cast<uint32_t*>(st + SL::ROWS);
        const long long g = a.gen0 + pg;
        const u32x4 w = philox4x32_10(cid, (uint32_t)g, (uint32_t)pb << 16, (uint32_t)((unsigned long long)g >> 32),
                                      a.key0, a.key1);
        uint32_t r1, r2;
        draw_pair(w.x, w.y, a.Mdraw, r1, r2);
        r1 += a.zstart; if (r1 >= a.zcap) r1 -= a.zcap;
        r2 += a.zstart; if (r2 >= a.zcap) r2 -= a.zcap;
        idx[lane] = r1;
        idx[32 + lane] = r2;
        idx[64 + lane] = w.z;
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2 * C; ++it) {
            const int chunk = it * 32 + lane;
            const int row = chunk / C;
            const int piece = chunk - row * C;
            cp_async16(st + row * LDS + piece * 4, a.Z + (size_t)idx[row] * LD + piece * 4);
        }
        if (++pb == nb) { pb = 0; ++pg; }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < Q) prefetch(s);
        cp_async_commit();
    }

    // ---- consumer ---------------------------------------------------------------------------
    int cg = 0, cb = 0;
    float lp_gen0 = lp;
    for (int q = 0; q < Q; ++q) {
        if (q + STAGES - 1 < Q) prefetch((q + STAGES - 1) % STAGES);
        cp_async_commit();                       // (possibly empty) keeps the group count uniform
        cp_async_wait<STAGES - 1>();
        __syncwarp();

        const float* st = wbase + (q % STAGES) * SL::FLOATS;
        const uint32_t* idx = reinterpret_cast<const uint32_t*>(st + SL::ROWS);
        const BlockDesc bd = a.blocks[cb];
        const long long g = a.gen0 + cg;
        const long long grel = a.grel0 + cg;

        float dz[LD];
        {
            const float4* ra = reinterpret_cast<const float4*>(st + lane * LDS);
            const float4* rb = reinterpret_cast<const float4*>(st + (32 + lane) * LDS);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4 A = ra[c], B = rb[c];
                dz[4 * c + 0] = A.x - B.x; dz[4 * c + 1] = A.y - B.y;
                dz[4 * c + 2] = A.z - B.z; dz[4 * c + 3] = A.w - B.w;
            }
        }
        const uint32_t wacc = idx[64 + lane];

        float xp[LD];
#pragma unroll
        for (int q4 = 0; q4 < C; ++q4) {
            if ((bd.mask >> (4 * q4)) & 0xFu) {      // warp-uniform
                const u32x4 wn = philox4x32_10(cid, (uint32_t)g, ((uint32_t)cb << 16) | (uint32_t)(1 + q4),
                                               (uint32_t)((unsigned long long)g >> 32), a.key0, a.key1);
                float z[4];
                box_muller(wn.x, wn.y, z[0], z[1]);
                box_muller(wn.z, wn.w, z[2], z[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 4 * q4 + j;
                    const float diff = fmaf(a.eps[k], z[j], bd.cscale * dz[k]);
                    xp[k] = ((bd.mask >> k) & 1u) ? x[k] + diff : x[k];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) xp[4 * q4 + j] = x[4 * q4 + j];
            }
        }

        const float lpp = mvn_logdensity<LD>(P, xp);
        const float u = uniform01(wacc);
        const float logu = __logf(u);
        const float delta = lpp - lp;
        bool acc;
        if (a.Ttab) acc = logu * a.Ttab[grel - 1] < delta;     // log(u) < Δ/T  (T = 0 -> greedy, Q4)
        else        acc = logu < delta;

        if (a.tr_r1 && active) {
            const size_t o = ((size_t)(grel - 1) * nb + cb) * a.N + chain;
            uint32_t l1 = idx[lane], l2 = idx[32 + lane];
            l1 = l1 >= a.zstart ? l1 - a.zstart : l1 + a.zcap - a.zstart;
            l2 = l2 >= a.zstart ? l2 - a.zstart : l2 + a.zcap - a.zstart;
            a.tr_r1[o] = l1; a.tr_r2[o] = l2; a.tr_u[o] = u; a.tr_lpprop[o] = lpp; a.tr_acc[o] = acc ? 1 : 0;
#pragma unroll
            for (int k = 0; k < LD; ++k)
                if (k < a.d) a.tr_xprop[(((size_t)(grel - 1) * nb + cb) * a.d + k) * a.N + chain] = xp[k];
        }

#pragma unroll
        for (int k = 0; k < LD; ++k) x[k] = acc ? xp[k] : x[k];
        lp = acc ? lpp : lp;
        cnt_acc += (acc && active) ? 1 : 0;
        __syncwarp();                            // stage q%STAGES is free for the next prefetch

        if (++cb == nb) {                        // end of generation: runchain! bookkeeping (src/demcz.jl:84-87)
            cb = 0;
            const bool changed = active && (lp != lp_gen0);
            cnt_run += (changed && cg != a.run_skip_gi) ? 1 : 0;
            cnt_win += (changed && cg != a.win_skip_gi) ? 1 : 0;
            if (a.hist && active) {
#pragma unroll
                for (int k = 0; k < LD; ++k)
                    if (k < a.d) a.hist[((size_t)(grel - 1) * a.d + k) * a.N + chain] = x[k];
                a.lphist[(size_t)(grel - 1) * a.N + chain] = lp;
            }
            if (ACC) {
#pragma unroll
                for (int k = 0; k < LD; ++k) {
                    const float t = x[k] - sh[k];
                    s1[k] += t;
                    s2[k] = fmaf(t, t, s2[k]);
                }
            }
            lp_gen0 = lp;
            ++cg;
        }
    }

    // ---- epilogue ------------------------------------------------------------------------------
    if (active) {
#pragma unroll
        for (int k = 0; k < LD; ++k)
            if (k < a.d) a.X[(size_t)k * a.ldx + chain] = x[k];
        a.logp[chain] = lp;
        if (a.chain_changed) a.chain_changed[chain] += cnt_run;
        if (a.Zw) {                              // Zmat[M+1,:] .= Xcurrent[ic,:]  (src/demcz.jl:88-91)
            uint32_t row = a.append_row0 + a.chain_offset + (uint32_t)chain;
            if (row >= a.zcap) row -= a.zcap;
            float4* dst = reinterpret_cast<float4*>(a.Zw + (size_t)row * LD);
#pragma unroll
            for (int c = 0; c < C; ++c) dst[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        }
        if (ACC) {
#pragma unroll
            for (int k = 0; k < LD; ++k)
                if (k < a.d) a.S1[(size_t)k * a.ldx + chain] = s1[k];
        }
    }
    __syncthreads();                             // red_* initialised
    {
        const int w_win = warp_sum(cnt_win), w_acc = warp_sum(cnt_acc);
        if (lane == 0) { atomicAdd(&red_cnt[0], (unsigned)w_win); atomicAdd(&red_cnt[1], (unsigned)w_acc); }
        if (ACC) {
#pragma unroll
            for (int k = 0; k < LD; ++k) {
                const float v = warp_sum(active ? s2[k] : 0.f);
                if (lane == 0) atomicAdd(&red_s2[k], v);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (red_cnt[0]) atomicAdd(a.win_changed, (unsigned long long)red_cnt[0]);
        if (red_cnt[1]) atomicAdd(a.accepted_total, (unsigned long long)red_cnt[1]);
    }
    if (ACC && threadIdx.x < a.d) atomicAdd(&a.S2[threadIdx.x], (double)red_s2[threadIdx.x]);
}

// Batched log-density: map(logobj, rows) (src/demcz.jl:17).  X SoA [k][ldx].
template <int LD>
__global__ void mvn_eval_kernel(const float* __restrict__ X, int n, int ldx, int d, float* __restrict__ out,
                                const __grid_constant__ MvnParams<LD> P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[LD];
#pragma unroll
    for (int k = 0; k < LD; ++k) x[k] = (k < d) ? X[(size_t)k * ldx + i] : 0.f;
    out[i] = mvn_logdensity<LD>(P, x);
}

}  // namespace demcz
