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
// the owning lane.  Archive rows sit in power-of-two slots (32/64/128 B) so that a row never
// straddles a DRAM fetch atom.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "philox.cuh"
#include "demcz_sync.cuh"

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
    // archive: row-major, row slot ZS floats (power of two; see zslot())
    const float* Z;
    uint32_t Mdraw;              // rows eligible for the draw
    uint32_t zstart, zcap;       // ring: physical row = (zstart + logical) mod zcap
    // generations of this launch
    long long gen0;              // absolute number (Philox counter) of the first generation
    long long grel0;             // 1-based number inside this context of the first generation
    int n_gens;
    PhiloxKeys keys;             // the ten round keys of Philox4x32-10, precomputed on the host
    int nblocks;
    BlockDesc blocks[kMaxBlocks];
    float2 eps2[16];             // eps_scale as column pairs
    const float* Ttab;           // temperatures, indexed by grel-1; nullptr -> plain accept
    // append (only when the last generation of the launch is a multiple of K)
    float* Zw;                   // archive base for the append, nullptr -> no append
    uint32_t append_row0;        // physical row of global chain 0's new row
    // history (store_chain): hist[hidx*d + k][chain], lphist[hidx][chain], hidx = grel/thin - 1
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
    float2 shift2[16];           // population mean at the start of the window (conditioning only), column pairs
    // history thinning: generation grel is stored iff grel % thin == 0, at index grel / thin - 1 (thin >= 1).
    // hist_phase = (grel0 - 1) % thin, hist_idx0 = (grel0 - 1) / thin: computed by the host, the kernel only counts
    int thin, hist_phase;
    unsigned int hist_idx0;
    // multi-GPU: the segment flag words of THIS rank's archive allocation (written by the peers) and how many
    // segments must have landed from every peer before this launch may gather (0: nothing to wait for)
    const unsigned long long* flags;
    int world, my_rank;
    unsigned long long need;
    // small-N persistent launches (PERS kernels, one CTA, world = 1): the launch spans several K-blocks; after every
    // generation g with (kphase + g) % K == 0 the CTA appends its rows, synchronises and the drawable view grows by N
    // rows (from the pers_delay-th append of the launch on: overlap sync mode hides the newest segment)
    int pers_K, pers_kphase, pers_delay;
    // opts.moments (MOM kernels): streaming first / second moments of EVERY generation of every chain,
    // mom[k] += sum (x_k - s_k), mom[64 + 64 k1 + k2] += sum (x_k1 - s_k1)(x_k2 - s_k2) for k1 <= k2 (FP64 global
    // accumulators; s = mom_shift2, fixed for the life of the accumulators)
    double* mom;
    float2 mom_shift2[16];
    // appending launches of a sharded run: the last CTA to finish publishes `done_seq` in `done_word` (release, gpu
    // scope) — the side stream waits on that word (stream memory operation) before it moves the rows to the peers,
    // so the compute stream carries nothing but step kernels.  done_word == nullptr: not used.
    unsigned int* done_word;
    unsigned int* done_ctr;
    unsigned int done_seq;
};

// Packed-FP32 (f32x2) arithmetic: one issue slot per two FMAs (SASS FFMA2 / FADD2 / FMUL2 on sm_100a).
typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

// Kernel-structure switches (compile time; tools/step_tune.cu times the combinations on the GPU):
enum : int {
    OPT_W_SMEM = 1,    // W / mu broadcast from shared memory (LDS.128) instead of the constant bank (LDCU -> uniform regs)
    OPT_JMAJOR = 2,    // quadratic form walks W column by column with LD/2 independent accumulators (ILP) instead of row pairs
    OPT_IDX_VEC = 4,   // gather: each lane copies pieces of consecutive staged rows (LDS.128 index reads, predicated LDGSTS)
    OPT_S2_SMEM = 8,   // Rhat sum-of-squares accumulators live in shared memory instead of registers
    OPT_ABL_NOJIT = 16, OPT_ABL_NOQUAD = 32   // ablations for tools/step_tune.cu only (wrong results)
};

// Closed-form likelihood constants: W = L^-1 (Sigma = L L^T) as ROW PAIRS (W[2 i2][j], W[2 i2 + 1][j]),
// so that y_{2 i2}, y_{2 i2 + 1} accumulate together in one FFMA2 with dx_j broadcast.
//   row-major  : entry (i2, j), j <= 2 i2 + 1, at i2 (i2 + 1) + j
//   column-major (OPT_JMAJOR): entry (j, i2), i2 >= j/2, at wj_off(j) + i2 - j/2; every column padded to an even count
template <int LD>
struct alignas(16) MvnParams {
    static constexpr int H = LD / 2;
    __host__ __device__ static constexpr int wj_cnt(int j) { return ((H - j / 2) + 1) & ~1; }
    __host__ __device__ static constexpr int wj_off(int j) { int o = 0; for (int t = 0; t < j; ++t) o += wj_cnt(t); return o; }
    static constexpr int NW_ROW = H * (H + 1);
    static constexpr int NW_COL = wj_off(LD);
    static constexpr int NW = NW_ROW > NW_COL ? NW_ROW : NW_COL;
    float2 W[NW];
    float2 mu2[H];
    float c0;
};

// archive row slot (floats) for a register width LD: next power of two, >= 8 floats (32-byte sector)
__host__ __device__ constexpr int zslot(int LD) { return LD <= 8 ? 8 : LD <= 16 ? 16 : 32; }

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

// predicated 16-byte async copy (no branch: the predicate rides on the LDGSTS)
__device__ __forceinline__ void cp_async16_if(bool p, float* smem_dst, const float* gmem_src)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n"
                 ::"r"(s), "l"(gmem_src), "r"((int)p) : "memory");
}

// log-density of the closed-form family: c0 - 0.5 |W (x - mu)|^2.  `Wc` / `muc` point either at the kernel
// parameters (constant bank: ptxas stages them through uniform registers, LDCU.128, and feeds FFMA2 its
// .F32x2 operand from there) or at a shared-memory copy (LDS.128 broadcasts).
template <int LD, bool JMAJOR>
__device__ __forceinline__ float mvn_logdensity(const float2* __restrict__ Wc, const float2* __restrict__ muc, float c0,
                                                const float2 (&x2)[LD / 2])
{
    using MP = MvnParams<LD>;
    constexpr int H = LD / 2;
    float dx[LD];
#pragma unroll
    for (int k2 = 0; k2 < H; k2 += 2) {
        const float4 m = *reinterpret_cast<const float4*>(muc + k2);
        const float2 a = fsub2(x2[k2], make_float2(m.x, m.y));
        const float2 b = fsub2(x2[k2 + 1], make_float2(m.z, m.w));
        dx[2 * k2] = a.x; dx[2 * k2 + 1] = a.y; dx[2 * k2 + 2] = b.x; dx[2 * k2 + 3] = b.y;
    }
    float2 q2 = make_float2(0.f, 0.f);
    if constexpr (JMAJOR) {
        float2 y[H];
#pragma unroll
        for (int i2 = 0; i2 < H; ++i2) y[i2] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < LD; ++j) {
            const float2 dj = make_float2(dx[j], dx[j]);
#pragma unroll
            for (int t = 0; t < MP::wj_cnt(j); t += 2) {
                const float4 w = *reinterpret_cast<const float4*>(Wc + MP::wj_off(j) + t);
                const int i2 = j / 2 + t;
                y[i2] = ffma2(make_float2(w.x, w.y), dj, y[i2]);
                if (i2 + 1 < H) y[i2 + 1] = ffma2(make_float2(w.z, w.w), dj, y[i2 + 1]);
            }
        }
#pragma unroll
        for (int i2 = 0; i2 < H; ++i2) q2 = ffma2(y[i2], y[i2], q2);
    } else {
#pragma unroll
        for (int i2 = 0; i2 < H; ++i2) {
            float2 y = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 2 * i2 + 2; j += 2) {
                const float4 w = *reinterpret_cast<const float4*>(Wc + i2 * (i2 + 1) + j);
                y = ffma2(make_float2(w.x, w.y), make_float2(dx[j], dx[j]), y);
                y = ffma2(make_float2(w.z, w.w), make_float2(dx[j + 1], dx[j + 1]), y);
            }
            q2 = ffma2(y, y, q2);
        }
    }
    return fmaf(-0.5f, q2.x + q2.y, c0);
}

// FULL: a single block covering every column (the common case) — no per-column masks.
// DENSE: archive rows back to back (slot = LD floats) instead of power-of-two slots.  The gather then pays for the
// rows that straddle a 128-byte line (measured 1.39x the DRAM bytes per row for d = 20), but a rank's appended rows
// are contiguous, so they cross NVLink as whole lines — the layout of choice when the exchange, not the gather,
// bounds the K-block (8 GPUs).
// MOM: per generation every warp stages its 32 x LD tile of (x - shift) in the staging buffer it has just consumed and
// accumulates the warp's LD x LD cross products in registers: the upper triangle is cut into 4 x 4 blocks, a lane owns
// one block for a subset of the chains (two LDS.128 + 8 FFMA2 per chain), column sums by lane k < LD.  One flush per
// launch: registers -> shared (per CTA) -> FP64 atomics.  mean_cov_chain of the full history without the history.
// PERS: one CTA holds all chains (N <= WARPS * 32) and runs every K-block between two host decisions in ONE launch:
// the append is a store + __syncthreads instead of a kernel boundary (13 us of launch latency per 10 generations at
// N = 100 otherwise).  The gather pipeline drains at every append (the next draw needs the grown archive).
template <int LD, bool ACC, bool FULL, int WARPS, int MINB, int OPT, bool DENSE = false, bool MOM = false, bool PERS = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
demcz_step_mvn_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ MvnParams<LD> P)
{
    using SL = StageLayout<LD>;
    using MP = MvnParams<LD>;
    constexpr int C = SL::C, LDS = SL::LDS, ZS = DENSE ? LD : zslot(LD), CP = ZS / 4, RPI = 32 / CP, STAGES = 2, H = LD / 2;
    constexpr bool W_SMEM = OPT & OPT_W_SMEM, JMAJOR = OPT & OPT_JMAJOR, IDX_VEC = OPT & OPT_IDX_VEC,
                   S2_SMEM = (OPT & OPT_S2_SMEM) && ACC;
    static_assert(!(DENSE && IDX_VEC), "the vectorised index read needs 64 % RPI == 0");
    constexpr int NWC = JMAJOR ? MP::NW_COL : MP::NW_ROW;
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned int red_cnt[2];
    __shared__ float red_s2[ACC ? LD : 1];
    __shared__ __align__(16) float2 scW[W_SMEM ? NWC + H : 1];
    __shared__ float2 s2s[S2_SMEM ? H * WARPS * 32 : 1];
    __shared__ float momG[MOM ? LD * LD + LD : 1];           // CTA-level cross products + column sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + warp * (STAGES * SL::FLOATS);
    const int chain = blockIdx.x * (WARPS * 32) + threadIdx.x;
    const bool active = chain < a.N;
    const int cl = active ? chain : a.N - 1;          // clamped: idle lanes shadow the last chain, never store
    const uint32_t cid = a.chain_offset + (uint32_t)cl;

    // programmatic dependent launch (no-ops without the launch attribute): let the next launch of the stream be
    // scheduled as soon as every CTA of this one is resident; wait here for the predecessor's writes (state, appended
    // rows, counters) — nothing written by a kernel is read or overwritten above this line
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // sharded runs: the newest drawable segment must have landed from every peer (normally it has: zero spins)
    wait_peer_segments(a.flags, a.world, a.my_rank, a.need);
    if (threadIdx.x < 2) red_cnt[threadIdx.x] = 0u;
    if (ACC && threadIdx.x < LD) red_s2[threadIdx.x] = 0.f;
    if (MOM) for (int t = threadIdx.x; t < LD * LD + LD; t += WARPS * 32) momG[t] = 0.f;
    if (W_SMEM)
        for (int t = threadIdx.x; t < NWC + H; t += WARPS * 32) scW[t] = t < NWC ? P.W[t] : P.mu2[t - NWC];
    __syncthreads();
    const float2* Wc = W_SMEM ? scW : P.W;
    const float2* muc = W_SMEM ? scW + NWC : P.mu2;

    float2 x2[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
        x2[k].x = (2 * k < a.d) ? a.X[(size_t)(2 * k) * a.ldx + cl] : 0.f;
        x2[k].y = (2 * k + 1 < a.d) ? a.X[(size_t)(2 * k + 1) * a.ldx + cl] : 0.f;
    }
    float lp = a.logp[cl];

    float2 s1[ACC ? H : 1], s2[(ACC && !S2_SMEM) ? H : 1];
    if (ACC) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
            s1[k].x = (2 * k < a.d) ? a.S1[(size_t)(2 * k) * a.ldx + cl] : 0.f;
            s1[k].y = (2 * k + 1 < a.d) ? a.S1[(size_t)(2 * k + 1) * a.ldx + cl] : 0.f;
            if (S2_SMEM) s2s[k * (WARPS * 32) + threadIdx.x] = make_float2(0.f, 0.f);
            else s2[k] = make_float2(0.f, 0.f);
        }
    }

    const int nb = FULL ? 1 : a.nblocks;
    const int Q = a.n_gens * nb;
    int cnt_run = 0, cnt_win = 0, cnt_acc = 0;

    // ---- producer side of the per-warp pipeline -------------------------------------------
    // lane -> (staged row, 16-byte piece of the row slot), lane-only so that the copy loop has compile-time
    // shared-memory offsets.  IDX_VEC: a lane group owns RPL consecutive staged rows (their indices are read
    // with LDS.128); otherwise copy instruction `it` covers rows it*RPI .. it*RPI+RPI-1.
    constexpr int RPL = (64 + RPI - 1) / RPI;           // copy instructions per stage
    constexpr bool RAGGED = (64 % RPI) != 0;            // (dense slots) the last instruction covers fewer than RPI rows
    const int lrow = lane / CP, lpiece = lane % CP;
    const bool pact = lpiece < C && lrow < RPI;         // (dense slots) 32 % CP lanes have no piece
    const int soff = (IDX_VEC ? lrow * RPL : lrow) * LDS + lpiece * 4;
    int pg = 0, pb = 0;
    uint32_t mdraw_cur = a.Mdraw;                       // PERS: grows by N at the launch's appends
    int pers_appends = 0;
    auto prefetch = [&](int stage) {
        float* st = wbase + stage * SL::FLOATS;
        uint32_t* idx = reinterpret_cast<uint32_t*>(st + SL::ROWS);
        const long long g = a.gen0 + pg;
        const u32x4 w = philox4x32_10(cid, (uint32_t)g, (uint32_t)pb << 16, (uint32_t)((unsigned long long)g >> 32), a.keys);
        uint32_t r1, r2;
        draw_pair(w.x, w.y, PERS ? mdraw_cur : a.Mdraw, r1, r2);
        r1 += a.zstart; if (r1 >= a.zcap) r1 -= a.zcap;
        r2 += a.zstart; if (r2 >= a.zcap) r2 -= a.zcap;
        idx[lane] = r1;
        idx[32 + lane] = r2;
        idx[64 + lane] = w.z;
        __syncwarp();
        const float* zb = a.Z + lpiece * 4;
        if constexpr (IDX_VEC) {
            uint32_t rows[RPL];
#pragma unroll
            for (int v = 0; v < RPL / 4; ++v) {
                const uint4 t = *reinterpret_cast<const uint4*>(idx + lrow * RPL + 4 * v);
                rows[4 * v] = t.x; rows[4 * v + 1] = t.y; rows[4 * v + 2] = t.z; rows[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int it = 0; it < RPL; ++it) cp_async16_if(pact, st + it * LDS + soff, zb + (size_t)rows[it] * ZS);
        } else {
#pragma unroll
            for (int it = 0; it < RPL; ++it) {
                const uint32_t row = idx[it * RPI + lrow];          // (ragged tail: reads the accept words, copy predicated off)
                const bool p = RAGGED && it == RPL - 1 ? (pact && it * RPI + lrow < 64) : pact;
                cp_async16_if(p, st + it * RPI * LDS + soff, zb + (size_t)row * ZS);
            }
        }
        if (!FULL) { if (++pb == nb) { pb = 0; ++pg; } } else { ++pg; }
    };

    if (0 < Q) prefetch(0);
    cp_async_commit();

    // ---- MOM: which 4 x 4 block(s) of the upper triangle and which chains this lane accumulates ---------------------
    constexpr int MB = C * (C + 1) / 2;                                        // blocks
    constexpr int MS = MB > 16 ? 1 : MB > 8 ? 2 : MB > 4 ? 4 : MB > 2 ? 8 : MB > 1 ? 16 : 32;    // chain splits per block
    constexpr int MI = (MB * MS + 31) / 32;                                    // blocks per lane (2 only for LD = 32)
    int mba[MOM ? MI : 1], mbb[MOM ? MI : 1];
    float2 macc[MOM ? MI * 8 : 1];
    float msum = 0.f;                                                          // column sum of column `lane` (lane < LD)
    if (MOM) {
#pragma unroll
        for (int it = 0; it < MI; ++it) {
            const int blk = (lane + 32 * it) / MS;                             // item = (block, split); split = item % MS
            int aa = 0, rem = blk;                                             // row-major enumeration of a <= b
            while (aa < C && rem >= C - aa) { rem -= C - aa; ++aa; }
            mba[it] = blk < MB ? aa : -1;
            mbb[it] = aa + rem;
#pragma unroll
            for (int e = 0; e < 8; ++e) macc[it * 8 + e] = make_float2(0.f, 0.f);
        }
    }

    // ---- consumer ---------------------------------------------------------------------------
    int cg = 0, cb = 0;
    float lp_gen0 = lp;
    for (int q = 0; q < Q; ++q) {
        // PERS: the proposal after an append boundary is prefetched behind the barrier below, not here
        const bool pers_boundary = PERS && (FULL || cb == nb - 1) && ((a.pers_kphase + cg + 1) % a.pers_K == 0);
        if (q + 1 < Q && !pers_boundary) prefetch((q + 1) & 1);
        cp_async_commit();                       // (possibly empty) keeps the group count uniform
        cp_async_wait<1>();
        __syncwarp();

        const float* st = wbase + (q & 1) * SL::FLOATS;
        const uint32_t* idx = reinterpret_cast<const uint32_t*>(st + SL::ROWS);
        const long long g = a.gen0 + cg;
        const long long grel = a.grel0 + cg;
        const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)((unsigned long long)g >> 32);
        const uint32_t mask = FULL ? 0xFFFFFFFFu : a.blocks[cb].mask;
        const float cs = FULL ? a.blocks[0].cscale : a.blocks[cb].cscale;
        const float2 cs2 = make_float2(cs, cs);

        // x' = x + c (Z[r1] - Z[r2]) + eps .* z
        float2 xp[H];
        {
            const float4* ra = reinterpret_cast<const float4*>(st + lane * LDS);
            const float4* rb = reinterpret_cast<const float4*>(st + (32 + lane) * LDS);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4 A = ra[c], B = rb[c];
                xp[2 * c] = fsub2(make_float2(A.x, A.y), make_float2(B.x, B.y));
                xp[2 * c + 1] = fsub2(make_float2(A.z, A.w), make_float2(B.z, B.w));
            }
        }
        const uint32_t wacc = idx[64 + lane];
#pragma unroll
        for (int q8 = 0; q8 < (LD + 7) / 8; ++q8) {
            if ((OPT & OPT_ABL_NOJIT) ? false : (FULL || ((mask >> (8 * q8)) & 0xFFu))) {      // warp-uniform
                const u32x4 wn = philox4x32_10(cid, g_lo, ((uint32_t)cb << 16) | (uint32_t)(1 + q8), g_hi, a.keys);
                const uint32_t ww[4] = {wn.x, wn.y, wn.z, wn.w};
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const int k2 = 4 * q8 + j;               // column pair index
                    if (k2 < H) {
                        xp[k2] = ffma2(cs2, xp[k2], ffma2(a.eps2[k2], box_muller16(ww[j]), x2[k2]));
                        xp[k2 + 1] = ffma2(cs2, xp[k2 + 1], ffma2(a.eps2[k2 + 1], box_muller16(ww[j + 1]), x2[k2 + 1]));
                    }
                }
            }
        }
        if (!FULL) {
#pragma unroll
            for (int k = 0; k < H; ++k) {
                xp[k].x = ((mask >> (2 * k)) & 1u) ? xp[k].x : x2[k].x;
                xp[k].y = ((mask >> (2 * k + 1)) & 1u) ? xp[k].y : x2[k].y;
            }
        }

        const float lpp = (OPT & OPT_ABL_NOQUAD) ? xp[0].x + xp[H - 1].y : mvn_logdensity<LD, JMAJOR>(Wc, muc, P.c0, xp);
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
                if (k < a.d) a.tr_xprop[(((size_t)(grel - 1) * nb + cb) * a.d + k) * a.N + chain] = (k & 1) ? xp[k / 2].y : xp[k / 2].x;
        }

#pragma unroll
        for (int k = 0; k < H; ++k) { x2[k].x = acc ? xp[k].x : x2[k].x; x2[k].y = acc ? xp[k].y : x2[k].y; }
        lp = acc ? lpp : lp;
        cnt_acc += (acc && active) ? 1 : 0;
        __syncwarp();                            // stage q&1 is free for the next prefetch

        if (FULL || ++cb == nb) {                // end of generation: runchain! bookkeeping (src/demcz.jl:84-87)
            cb = 0;
            const bool changed = active && (lp != lp_gen0);
            cnt_run += (changed && cg != a.run_skip_gi) ? 1 : 0;
            cnt_win += (changed && cg != a.win_skip_gi) ? 1 : 0;
            if (a.hist) {                        // thinned history: generation t (counted from the last stored one) is kept iff t % thin == 0
                const int t = a.hist_phase + cg + 1;
                if (active && t % a.thin == 0) {
                    const size_t hi = (size_t)a.hist_idx0 + (size_t)(t / a.thin) - 1;
#pragma unroll
                    for (int k = 0; k < LD; ++k)
                        if (k < a.d) a.hist[(hi * a.d + k) * a.N + chain] = (k & 1) ? x2[k / 2].y : x2[k / 2].x;
                    a.lphist[hi * a.N + chain] = lp;
                }
            }
            if (ACC) {
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    const float2 t = fsub2(x2[k], a.shift2[k]);
                    s1[k] = fadd2(s1[k], t);
                    if (S2_SMEM) { float2& r = s2s[k * (WARPS * 32) + threadIdx.x]; r = ffma2(t, t, r); }
                    else s2[k] = ffma2(t, t, s2[k]);
                }
            }
            if (MOM) {
                // stage (x - shift) of the warp's 32 chains in the buffer consumed above (free until the next iteration's prefetch)
                float* mt = wbase + (q & 1) * SL::FLOATS;
                float4* mrow = reinterpret_cast<float4*>(mt + lane * LDS);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float2 t0 = active ? fsub2(x2[2 * c], a.mom_shift2[2 * c]) : make_float2(0.f, 0.f);
                    const float2 t1 = active ? fsub2(x2[2 * c + 1], a.mom_shift2[2 * c + 1]) : make_float2(0.f, 0.f);
                    mrow[c] = make_float4(t0.x, t0.y, t1.x, t1.y);
                }
                __syncwarp();
                if (lane < LD) {
#pragma unroll 8
                    for (int c = 0; c < 32; ++c) msum += mt[c * LDS + lane];
                }
#pragma unroll
                for (int it = 0; it < MI; ++it) {
                    if (mba[it] >= 0) {
                        const int split = (lane + 32 * it) % MS;
                        const float* pa = mt + 4 * mba[it];
                        const float* pb = mt + 4 * mbb[it];
#pragma unroll 4
                        for (int i = 0; i < 32 / MS; ++i) {
                            const int c = i * MS + split;
                            const float4 A = *reinterpret_cast<const float4*>(pa + c * LDS);
                            const float4 B = *reinterpret_cast<const float4*>(pb + c * LDS);
                            const float2 b01 = make_float2(B.x, B.y), b23 = make_float2(B.z, B.w);
                            macc[it * 8 + 0] = ffma2(make_float2(A.x, A.x), b01, macc[it * 8 + 0]);
                            macc[it * 8 + 1] = ffma2(make_float2(A.x, A.x), b23, macc[it * 8 + 1]);
                            macc[it * 8 + 2] = ffma2(make_float2(A.y, A.y), b01, macc[it * 8 + 2]);
                            macc[it * 8 + 3] = ffma2(make_float2(A.y, A.y), b23, macc[it * 8 + 3]);
                            macc[it * 8 + 4] = ffma2(make_float2(A.z, A.z), b01, macc[it * 8 + 4]);
                            macc[it * 8 + 5] = ffma2(make_float2(A.z, A.z), b23, macc[it * 8 + 5]);
                            macc[it * 8 + 6] = ffma2(make_float2(A.w, A.w), b01, macc[it * 8 + 6]);
                            macc[it * 8 + 7] = ffma2(make_float2(A.w, A.w), b23, macc[it * 8 + 7]);
                        }
                    }
                }
                __syncwarp();                    // the tile is dead: the next iteration may prefetch into it
            }
            lp_gen0 = lp;
            ++cg;
            if (PERS && pers_boundary) {         // append (src/demcz.jl:88-91), make it visible to the CTA, grow the view
                if (active) {
                    uint32_t row = a.append_row0 + (uint32_t)pers_appends * (uint32_t)a.N + a.chain_offset + (uint32_t)chain;
                    float4* dst = reinterpret_cast<float4*>(a.Zw + (size_t)row * ZS);
#pragma unroll
                    for (int c = 0; c < C; ++c) dst[c] = make_float4(x2[2 * c].x, x2[2 * c].y, x2[2 * c + 1].x, x2[2 * c + 1].y);
                    if (C % 2 == 1 && C < CP) dst[C] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                ++pers_appends;
                __threadfence_block();
                __syncthreads();
                if (pers_appends > a.pers_delay) mdraw_cur += (uint32_t)a.N;
                if (q + 1 < Q) prefetch((q + 1) & 1);
                cp_async_commit();
            }
        }
    }

    // ---- epilogue ------------------------------------------------------------------------------
    if (active) {
#pragma unroll
        for (int k = 0; k < LD; ++k)
            if (k < a.d) a.X[(size_t)k * a.ldx + chain] = (k & 1) ? x2[k / 2].y : x2[k / 2].x;
        a.logp[chain] = lp;
        if (a.chain_changed) a.chain_changed[chain] += cnt_run;
        if (!PERS && a.Zw) {                     // Zmat[M+1,:] .= Xcurrent[ic,:]  (src/demcz.jl:88-91)
            uint32_t row = a.append_row0 + a.chain_offset + (uint32_t)chain;
            if (row >= a.zcap) row -= a.zcap;
            float4* dst = reinterpret_cast<float4*>(a.Zw + (size_t)row * ZS);
#pragma unroll
            for (int c = 0; c < C; ++c) dst[c] = make_float4(x2[2 * c].x, x2[2 * c].y, x2[2 * c + 1].x, x2[2 * c + 1].y);
            // complete the last 32-byte sector with slot padding: whole sectors need no read-modify-write fill in L2
            if (C % 2 == 1 && C < CP) dst[C] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (ACC) {
#pragma unroll
            for (int k = 0; k < LD; ++k)
                if (k < a.d) a.S1[(size_t)k * a.ldx + chain] = (k & 1) ? s1[k / 2].y : s1[k / 2].x;
        }
    }
    __syncthreads();                             // red_* initialised
    {
        const int w_win = warp_sum(cnt_win), w_acc = warp_sum(cnt_acc);
        if (lane == 0) { atomicAdd(&red_cnt[0], (unsigned)w_win); atomicAdd(&red_cnt[1], (unsigned)w_acc); }
        if (ACC) {
#pragma unroll
            for (int k = 0; k < LD; ++k) {
                const float2 sv = S2_SMEM ? s2s[(k / 2) * (WARPS * 32) + threadIdx.x] : s2[S2_SMEM ? 0 : k / 2];
                const float v = warp_sum(active ? ((k & 1) ? sv.y : sv.x) : 0.f);
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

    if (MOM) {                                   // registers -> CTA accumulators -> FP64 global (upper triangle, k1 <= k2)
#pragma unroll
        for (int it = 0; it < MI; ++it) {
            if (mba[it] >= 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float* g = momG + (4 * mba[it] + i) * LD + 4 * mbb[it];
                    atomicAdd(g + 0, macc[it * 8 + 2 * i].x); atomicAdd(g + 1, macc[it * 8 + 2 * i].y);
                    atomicAdd(g + 2, macc[it * 8 + 2 * i + 1].x); atomicAdd(g + 3, macc[it * 8 + 2 * i + 1].y);
                }
            }
        }
        if (lane < LD) atomicAdd(momG + LD * LD + lane, msum);
        __syncthreads();
        for (int t = threadIdx.x; t < LD * LD + LD; t += WARPS * 32) {
            if (t < LD * LD) {
                const int k1 = t / LD, k2 = t - k1 * LD;
                if (k1 <= k2 && k2 < a.d) atomicAdd(a.mom + 64 + 64 * k1 + k2, (double)momG[t]);
            } else if (t - LD * LD < a.d) {
                atomicAdd(a.mom + (t - LD * LD), (double)momG[t]);
            }
        }
    }

    if (a.done_word) {                           // the __syncthreads above ordered every thread's append stores before this point
        if (threadIdx.x == 0) {
            __threadfence();                     // cumulative: the CTA's rows are visible at gpu scope before the ticket
            if (atomicAdd(a.done_ctr, 1u) == gridDim.x - 1) {
                *a.done_ctr = 0u;
                __threadfence();                 // acquire side of the ticket: every CTA's rows happen-before the word
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.done_word), "r"(a.done_seq) : "memory");
            }
        }
    }
}

// Batched log-density: map(logobj, rows) (src/demcz.jl:17).  X SoA [k][ldx].
template <int LD>
__global__ void mvn_eval_kernel(const float* __restrict__ X, int n, int ldx, int d, float* __restrict__ out,
                                const __grid_constant__ MvnParams<LD> P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 x2[LD / 2];
#pragma unroll
    for (int k = 0; k < LD / 2; ++k) {
        x2[k].x = (2 * k < d) ? X[(size_t)(2 * k) * ldx + i] : 0.f;
        x2[k].y = (2 * k + 1 < d) ? X[(size_t)(2 * k + 1) * ldx + i] : 0.f;
    }
    out[i] = mvn_logdensity<LD, false>(P.W, P.mu2, P.c0, x2);
}

}  // namespace demcz
