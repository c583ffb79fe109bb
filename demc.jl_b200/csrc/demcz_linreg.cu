// demcz_linreg.cu — the linear-regression Gaussian log-likelihood path of the DEMCz generation loop:
//     logobj(b) = -0.5 * sum((y .- X*b).^2)                       (reference test/example_linreg.jl:32)
// evaluated for ALL chains of a generation at once as the one dense contraction of the sampler,
// X (n x p) times the proposals (N x p)^T, on the 5th-generation tensor cores:
//
//   centred form   SSR(b) = SSR(b0) - 2 d^T g + |X d|^2,  d = b - b0,  g = X^T (y - X b0)    (FP64 constants,
//                  b0 = OLS computed at registration).  Same GEMM shape as y - X b, but its output is O(|d|),
//                  so the FP32-equivalent product keeps Δlogπ accurate although logπ ~ -n/2.
//   tensor cores   D[chain, row] = sum_k d[chain,k] X[row,k] with tcgen05.mma kind::tf32, M = 128 chains
//                  (TMEM lane = chain), N = TILE_N data rows, accumulators in TMEM; FP32-equivalent accuracy by
//                  the 3xTF32 split  d_hi X_hi + d_lo X_hi + d_hi X_lo  (X pre-split once, d split per proposal)
//   TMA            X_hi / X_lo row tiles and the d_hi / d_lo chain tile arrive by cp.async.bulk.tensor (128-byte
//                  swizzle, K-major) on mbarriers; a 3-role warp-specialised CTA per SM: TMA producer, MMA issuer,
//                  4 epilogue warps
//   fused epilogue tcgen05.ld the accumulator tile, square-and-sum along the data rows inside the thread that
//                  owns the chain (no cross-lane reduction), FP32 per tile, FP64 across tiles; two TMEM
//                  accumulator stages so the epilogue of tile t overlaps the MMAs of tile t+1
//   determinism    fixed (chain tile, row chunk) -> CTA schedule, partial sums reduced in fixed order
//
// Around the GEMM, per generation and block: propose (Philox, archive gather, x', d split, d^T g) and
// accept/select (tempered Metropolis in FP64, history, counters, Rhat accumulators, archive append) — the
// same semantics as demcz_step.cuh (reference src/demcz.jl:167-203, 80-93; src/demcz_anneal.jl:142-178).
#include "demcz_internal.h"
#include "demcz_step.cuh"
#include "demcz_sync.cuh"
#include "demcz_tcgen05.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <vector>

using namespace demcz;

#define CUL(c, call)                                                                                  \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return demcz_fail((c), e_ == cudaErrorMemoryAllocation ? DEMCZ_ERR_OOM : DEMCZ_ERR_CUDA,  \
                              "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
    } while (0)

struct demcz_linreg_state {
    int64_t n = 0;
    int p = 0, ppad = 0, ksteps = 0, ksteps16 = 0;  // K steps of 8 TF32 / 16 FP16 elements covering p
    float *Xhi = nullptr, *Xlo = nullptr;          // [n][ppad] row-major, TF32-rounded hi / lo parts
    double *d_b0 = nullptr, *d_g = nullptr;        // [64]
    double ssr0 = 0.0;
    int Npad = 0;                                  // chains rounded up to 128 (MMA M)
    float *Dhi = nullptr, *Dlo = nullptr;          // [Npad][ppad] delta = x' - b0, split
    float* Xprop = nullptr;                        // [d][ldx]
    double *lin = nullptr, *partial = nullptr;     // [Npad], [R][Npad]
    double* lp64 = nullptr;                        // [ldx] current log-objective
    double* lphist64 = nullptr;                    // [G][N]
    uint32_t* wacc = nullptr;                      // [ldx] accept word of the pending proposal
    float* geo = nullptr;                          // [ldx] snooker generations: log Jacobian factor of the pending proposal
    int tile_n = 0, tiles = 0, R = 0, tiles_per_chunk = 0;
    int terms = 3;                                 // 3 MMA terms per product (FP32-equivalent) or 1 (measurement only)
    // kind 1 (default): 2-way FP16 split on tcgen05 kind::f16 — x = hi + lo, both FP16, operands pre-scaled by powers of
    // two (one exponent per regressor column folded into X, its inverse into d; one exponent per chain for d), the
    // same three products d_lo X_hi + d_hi X_lo + d_hi X_hi at twice the TF32 MMA rate and half the X bytes.
    // kind 0: the 3xTF32 split (fallback when the FP16 path's self-test fails; DEMCZ_LINREG_KIND=tf32 forces it).
    int kind = 1;
    int link = 0;                                  // 0: Gaussian linear regression, 1: Bernoulli-logit GLM (FP16 kind only)
    __half* Xh = nullptr;                          // [n][2 ppad] row-major: [hi(ppad) | lo(ppad)] of X[:,k] * 2^xexp[k]
    uint32_t* Dh = nullptr;                        // [Npad][ppad] words: [hi pairs (ppad/2) | lo pairs (ppad/2)] of d * 2^(Dexp - xexp)
    int* Dexp = nullptr;                           // [Npad] per-chain exponent
    int xexp[64] = {0};                            // per-column exponent
    int swap_halves = 0;                           // which half of a 32-bit TMEM column holds the even K element (probed)
    // the two half-populations are pipelined: the latency-bound prepare / accept kernels of one half run under the GEMM of
    // the other (two side streams, forked from and joined to the context's stream once per demcz_linreg_generations call)
    cudaStream_t hstream[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    std::string fallback_reason;                   // why kind 1 was not used (self-test message), if it was asked for
    CUtensorMap tmXh;
    CUtensorMap tmXhi, tmXlo;
    int sms = 148;
    unsigned long long ssr_launches = 0;
    std::vector<cudaEvent_t> ev;                   // profile=1: brackets of the SSR kernel launches of the last run
    size_t ev_used = 0;
};

namespace {

// ------------------------------------------------------------------------------------------------
// registration-time kernels (FP64, one-off)
// ------------------------------------------------------------------------------------------------
// X col-major FP64 (n x p) -> row-major [n][ppad] TF32 hi / lo
__global__ void split_x_kernel(const double* __restrict__ X, long long n, int p, int ppad, float* __restrict__ hi,
                               float* __restrict__ lo)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < ppad; ++k) {
        float h = 0.f, l = 0.f;
        if (k < p) {
            const double x = X[i + n * k];
            h = tc::to_tf32((float)x);
            l = tc::to_tf32((float)(x - (double)h));
        }
        hi[i * ppad + k] = h;
        lo[i * ppad + k] = l;
    }
}

// column-wise max |X[:,k]| (as the bit pattern of a non-negative float: integer max == float max)
__global__ void colabsmax_kernel(const double* __restrict__ X, long long n, int p, int* __restrict__ out)
{
    const int k = blockIdx.y;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf((float)X[i + n * k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out + k, __float_as_int(m));
}

// X col-major FP64 (n x p) -> row-major [n][2 ppad] FP16: [hi | lo] of X[:,k] * 2^xexp[k]  (|scaled| < 2^15)
struct XExp { int e[64]; };
__global__ void split_x_f16_kernel(const double* __restrict__ X, long long n, int p, int ppad, XExp xe, __half* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __half* row = out + i * (2 * ppad);
    for (int k = 0; k < ppad; ++k) {
        __half h = __float2half_rn(0.f), l = h;
        if (k < p) {
            const double x = scalbn(X[i + n * k], xe.e[k]);
            h = __float2half_rn((float)x);
            l = __float2half_rn((float)(x - (double)__half2float(h)));
        }
        row[k] = h;
        row[ppad + k] = l;
    }
}

// Per-block partial Gram matrices of [X | y]: out[block][a*(p+1) + b] = sum over the block's rows of col_a * col_b
// (a, b <= p; column p is y).  The host adds the blocks in order: registration is bit-reproducible.
__global__ void gram_kernel(const double* __restrict__ X, const double* __restrict__ y, long long n, int p,
                            long long rows_per_block, double* __restrict__ out)
{
    extern __shared__ double srow[];                  // [32 rows][p + 1]
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(n, r0 + rows_per_block);
    const int P1 = p + 1;
    const int npairs = P1 * P1;
    double acc[20];                                    // up to ceil(65*65/256) = 17 pairs per thread
#pragma unroll
    for (int t = 0; t < 20; ++t) acc[t] = 0.0;
    for (long long base = r0; base < r1; base += 32) {
        const int nr = (int)min((long long)32, r1 - base);
        __syncthreads();
        for (int t = threadIdx.x; t < nr * P1; t += blockDim.x) {
            const int k = t / nr, r = t - k * nr;      // consecutive threads -> consecutive rows (coalesced)
            srow[r * P1 + k] = k < p ? X[base + r + n * k] : y[base + r];
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 20; ++t) {
            const int pr = threadIdx.x + t * blockDim.x;
            if (pr < npairs) {
                const int a = pr / P1, b = pr - a * P1;
                double s = 0.0;
                for (int r = 0; r < nr; ++r) s += srow[r * P1 + a] * srow[r * P1 + b];
                acc[t] += s;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 20; ++t) {
        const int pr = threadIdx.x + t * blockDim.x;
        if (pr < npairs) out[(size_t)blockIdx.x * npairs + pr] = acc[t];
    }
}

// Per-block partials of g = X^T (y - X b0) and ssr0 = |y - X b0|^2: out[block][0..p) = g, out[block][p] = ssr0.
// Each warp owns a row of the shared accumulator, the block's warps and then (on the host) the blocks are added
// in a fixed order.
__global__ void resid_kernel(const double* __restrict__ X, const double* __restrict__ y, long long n, int p,
                             const double* __restrict__ b0, double* __restrict__ out)
{
    __shared__ double sb[64];
    __shared__ double sg[8][65];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) sb[threadIdx.x] = threadIdx.x < p ? b0[threadIdx.x] : 0.0;
    for (int t = threadIdx.x; t < 8 * 65; t += blockDim.x) (&sg[0][0])[t] = 0.0;
    __syncthreads();
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;                  // warp-uniform trip count: the shuffles below need all lanes
        const bool valid = i < n;
        double r = valid ? y[i] : 0.0;
        if (valid) for (int k = 0; k < p; ++k) r -= X[i + n * k] * sb[k];
        for (int k = 0; k < p; ++k) {
            double v = valid ? X[i + n * k] * r : 0.0;
            v = warp_sum(v);
            if (lane == 0) sg[warp][k] += v;
        }
        double v = warp_sum(r * r);
        if (lane == 0) sg[warp][64] += v;
    }
    __syncthreads();
    if (threadIdx.x <= p) {
        const int k = threadIdx.x < p ? threadIdx.x : 64;
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sg[w][k];
        out[(size_t)blockIdx.x * (p + 1) + threadIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// per-proposal kernels around the GEMM
// ------------------------------------------------------------------------------------------------
struct LinArgs {
    // state
    float* X; int N, ldx, d, ppad;
    int chain_lo, chain_hi;       // this launch handles chains [chain_lo, chain_hi) (the two half-populations are pipelined)
    uint32_t chain_offset;
    const float* Z; int ZS; uint32_t Mdraw, zstart, zcap;
    long long gen, grel;          // absolute / context-relative generation of this proposal
    PhiloxKeys keys;
    int block, nblocks; uint32_t mask_lo, mask_hi; float cscale;
    float eps[64];
    __device__ bool in_block(int k) const { return (((k < 32 ? mask_lo : mask_hi) >> (k & 31)) & 1u) != 0; }
    __device__ uint32_t group_mask(int k) const { return ((k < 32 ? mask_lo : mask_hi) >> (k & 24)) & 0xFFu; }   // 8-column jitter group of k
    // outputs of propose
    float* Xprop; uint32_t* wacc; float *Dhi, *Dlo; double* lin;
    // FP16-split operands (kind 1): packed words + per-chain exponent; column exponents; TMEM half order
    int kind; uint32_t* Dh; int* Dexp; int xexp[64]; int swap_halves;
    const double *b0, *g;
    // trace
    uint32_t *tr_r1, *tr_r2; float *tr_u, *tr_xprop;
    // snooker generation (see snooker_mvn_kernel in demcz_lib.cu): mode 2 of the prepare kernel; geo = (d-1) log|1 + gamma_s proj|
    float gamma_s; float* geo;
    // multi-GPU: segment flag words to poll before the first gather (demcz_sync.cuh); need = 0: nothing to wait for
    const unsigned long long* flags; int world, my_rank; unsigned long long need;
};

// Philox draw, archive gather, x' = x + c (Z[r1]-Z[r2]) + eps .* z on the block (src/demcz.jl:174-188), then the
// GEMM operands: d = x' - b0 split into TF32 hi / lo, and the FP64 linear term d^T g.
// propose == false: d from the current state (initial log-objective, demcz_eval_logdensity).
__global__ void linreg_prepare_kernel(const LinArgs a, bool propose, bool snooker = false)
{
    const int chain = a.chain_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (a.need != 0ull) {                              // sharded runs: the newest drawable segment has landed from every peer
        wait_peer_segments(a.flags, a.world, a.my_rank, a.need);
        __syncthreads();
    }
    if (chain >= a.chain_hi) return;
    const uint32_t cid = a.chain_offset + (uint32_t)chain;
    const uint32_t g_lo = (uint32_t)a.gen, g_hi = (uint32_t)((unsigned long long)a.gen >> 32);
    const float *z1 = nullptr, *z2 = nullptr;
    float sproj = 0.f;                                 // snooker: gamma_s ((z1 - z2) . e) / (e . e), e = x - z
    const float* zs = nullptr;
    if (propose && snooker) {
        const u32x4 w = philox4x32_10(cid, g_lo, kSnookerCall, g_hi, a.keys);
        uint32_t z, r1, r2;
        draw_triple(__umulhi(w.x, a.Mdraw), __umulhi(w.y, a.Mdraw - 1u), __umulhi(w.z, a.Mdraw - 2u), z, r1, r2);
        float gs = a.gamma_s;
        if (gs == 0.f) gs = 1.2f + uniform01(philox4x32_10(cid, g_lo, kSnookerCall + 1u, g_hi, a.keys).x);
        if (a.tr_r1) {
            const size_t o = ((size_t)(a.grel - 1) * a.nblocks) * a.N + chain;
            a.tr_r1[o] = r1; a.tr_r2[o] = r2; a.tr_u[o] = uniform01(w.w);
        }
        auto phys = [&](uint32_t r) { r += a.zstart; if (r >= a.zcap) r -= a.zcap; return a.Z + (size_t)r * a.ZS; };
        zs = phys(z); z1 = phys(r1); z2 = phys(r2);
        float ee = 0.f, num = 0.f;
        for (int k = 0; k < a.d; ++k) {
            const float e = a.X[(size_t)k * a.ldx + chain] - zs[k];
            ee = fmaf(e, e, ee);
            num = fmaf(z1[k] - z2[k], e, num);
        }
        sproj = ee > 0.f ? gs * (num / ee) : 0.f;
        a.geo[chain] = ee > 0.f ? (float)(a.d - 1) * __logf(fabsf(1.f + sproj)) : -INFINITY;      // x == z: rejected
        a.wacc[chain] = w.w;
    } else if (propose) {
        const u32x4 w = philox4x32_10(cid, g_lo, (uint32_t)a.block << 16, g_hi, a.keys);
        uint32_t r1, r2;
        draw_pair(w.x, w.y, a.Mdraw, r1, r2);
        if (a.tr_r1) {
            const size_t o = ((size_t)(a.grel - 1) * a.nblocks + a.block) * a.N + chain;
            a.tr_r1[o] = r1; a.tr_r2[o] = r2; a.tr_u[o] = uniform01(w.z);
        }
        r1 += a.zstart; if (r1 >= a.zcap) r1 -= a.zcap;
        r2 += a.zstart; if (r2 >= a.zcap) r2 -= a.zcap;
        z1 = a.Z + (size_t)r1 * a.ZS;
        z2 = a.Z + (size_t)r2 * a.ZS;
        a.wacc[chain] = w.z;
    }
    // Columns in groups of 8 (= one Philox jitter call): every load of a group is issued before its first store, so
    // the (possibly aliasing) stores serialise groups, not columns — 4 memory round trips for p = 32 instead of 32.
    // Arithmetic and summation order are those of a plain column loop.
    const int LD4 = (a.d + 3) & ~3;                 // floats of an archive row that exist (rows are 16-byte aligned)
    double lin = 0.0;
    float dmax = 0.f;                               // kind 1: max_k |d_k| 2^-xexp[k]
    for (int k0 = 0; k0 < a.d; k0 += 8) {
        float x[8], dz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (k0 + j < a.d) ? a.X[(size_t)(k0 + j) * a.ldx + chain] : 0.f;
        float2 zn[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (propose && snooker) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dz[j] = (k0 + j < a.d) ? x[j] - zs[k0 + j] : 0.f;          // e = x - z
        } else if (propose) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a0 = *reinterpret_cast<const float4*>(z1 + k0), b0 = *reinterpret_cast<const float4*>(z2 + k0);
            const float4 a1 = (k0 + 4 < LD4) ? *reinterpret_cast<const float4*>(z1 + k0 + 4) : z;
            const float4 b1 = (k0 + 4 < LD4) ? *reinterpret_cast<const float4*>(z2 + k0 + 4) : z;
            dz[0] = a0.x - b0.x; dz[1] = a0.y - b0.y; dz[2] = a0.z - b0.z; dz[3] = a0.w - b0.w;
            dz[4] = a1.x - b1.x; dz[5] = a1.y - b1.y; dz[6] = a1.z - b1.z; dz[7] = a1.w - b1.w;
            if (a.group_mask(k0) != 0) {
                const u32x4 wn = philox4x32_10(cid, g_lo, ((uint32_t)a.block << 16) | (uint32_t)(1 + (k0 >> 3)), g_hi, a.keys);
                zn[0] = box_muller16(wn.x); zn[1] = box_muller16(wn.y); zn[2] = box_muller16(wn.z); zn[3] = box_muller16(wn.w);
            }
        }
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            float xp = x[j];
            if (propose && snooker) {
                if (k < a.d) xp = fmaf(sproj, dz[j], x[j]);
            } else if (propose && k < a.d && a.in_block(k)) {
                const float2 zz = zn[j >> 1];
                xp = fmaf(a.cscale, dz[j], fmaf(a.eps[k], (j & 1) ? zz.y : zz.x, x[j]));
            }
            x[j] = xp;
            const double dd = (double)xp - (k < a.d ? a.b0[k] : 0.0);
            const float df = (float)dd;
            if (a.kind == 1) {
                if (k < a.d) dmax = fmaxf(dmax, fabsf(scalbnf(df, -a.xexp[k])));
                continue;
            }
            hi[j] = k < a.d ? tc::to_tf32(df) : 0.f;
            lo[j] = k < a.d ? tc::to_tf32(df - hi[j]) : 0.f;
            // the GEMM sees hi + lo; use the same value in the linear term
            if (k < a.d) lin += ((double)hi[j] + (double)lo[j]) * a.g[k];
        }
        if (propose) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j;
                if (k < a.d) {
                    a.Xprop[(size_t)k * a.ldx + chain] = x[j];
                    if (a.tr_xprop) a.tr_xprop[(((size_t)(a.grel - 1) * a.nblocks + a.block) * a.d + k) * a.N + chain] = x[j];
                }
            }
        }
        if (a.kind == 1) continue;
        // operand rows are ppad (32 or 64) floats and zero beyond d: whole 16-byte stores
        float4* dh = reinterpret_cast<float4*>(a.Dhi + (size_t)chain * a.ppad + k0);
        float4* dl = reinterpret_cast<float4*>(a.Dlo + (size_t)chain * a.ppad + k0);
        dh[0] = make_float4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_float4(hi[4], hi[5], hi[6], hi[7]);
        dl[0] = make_float4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_float4(lo[4], lo[5], lo[6], lo[7]);
    }
    if (a.kind == 1) {
        // ---- FP16 split: d''_k = d_k 2^(f - xexp[k]) with the per-chain exponent f such that max_k |d''_k| is in [2^14, 2^15)
        // (exact power-of-two scalings), hi = fp16(d''), lo = fp16(d'' - hi): 22 significant bits for the large components,
        // absolute error <= 2^-25 (FP16 subnormal spacing) for the small ones, i.e. <= 2^-39 of the largest
        int f = 0;
        if (dmax > 0.f) { int ex; frexpf(dmax, &ex); f = min(100, max(-100, 15 - ex)); }
        const float* xsrc = propose ? a.Xprop : a.X;               // pass 1 left x' there (same thread: program order)
        uint32_t* drow = a.Dh + (size_t)chain * a.ppad;
        const int HW = a.ppad / 2;                                 // words of the hi part
        for (int k0 = 0; k0 < a.ppad; k0 += 8) {
            uint32_t wh[4], wl[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                unsigned short h16[2], l16[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int k = k0 + j + t;
                    float v = 0.f;
                    if (k < a.d) v = scalbnf((float)((double)xsrc[(size_t)k * a.ldx + chain] - a.b0[k]), f - a.xexp[k]);
                    const __half h = __float2half_rn(v);
                    const __half l = __float2half_rn(v - __half2float(h));
                    h16[t] = __half_as_ushort(h); l16[t] = __half_as_ushort(l);
                    // the GEMM sees hi + lo; use the same value in the linear term
                    if (k < a.d) lin += scalbn((double)__half2float(h) + (double)__half2float(l), a.xexp[k] - f) * a.g[k];
                }
                const int e = a.swap_halves ? 1 : 0;               // which K element sits in the low half-word of a TMEM column
                wh[j >> 1] = (uint32_t)h16[e] | ((uint32_t)h16[1 - e] << 16);
                wl[j >> 1] = (uint32_t)l16[e] | ((uint32_t)l16[1 - e] << 16);
            }
            *reinterpret_cast<uint4*>(drow + (k0 >> 1)) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            *reinterpret_cast<uint4*>(drow + HW + (k0 >> 1)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
        a.Dexp[chain] = f;
    }
    a.lin[chain] = lin;
}

struct AccArgs {
    float* X; const float* Xprop; double* lp; const uint32_t* wacc;
    int N, ldx, d, Npad, R;
    int chain_lo, chain_hi;
    const double *partial, *lin; double ssr0; int link;
    float T; int tempered;
    long long grel; int block, nblocks; bool last_block;
    // generation bookkeeping (last block only)
    double* lp_gen0;              // [ldx] log-objective at the start of the generation
    int* chain_changed; unsigned long long *win_changed, *accepted_total; int count_run, count_win;
    float* hist; double* lphist; long long hidx;      // history slot of this generation (thinning), or hist == nullptr
    float* Zw; int ZS; uint32_t append_row0, zcap, chain_offset;
    float* S1; double* S2; const float* shift;
    float* tr_lpprop; uint8_t* tr_acc;
    const float* geo;             // snooker generation: log of the Jacobian factor (d-1) log(|x'-z| / |x-z|), or nullptr
};

// log-objective of the proposal from the GEMM partials, tempered Metropolis accept (src/demcz.jl:197-203,
// src/demcz_anneal.jl:172-178) in FP64, select, and runchain!'s bookkeeping (src/demcz.jl:84-91).
__global__ void linreg_accept_kernel(const AccArgs a)
{
    __shared__ unsigned int red[2];
    __shared__ double red_s2[64];
    if (threadIdx.x < 2) red[threadIdx.x] = 0;
    if (threadIdx.x < 64) red_s2[threadIdx.x] = 0.0;
    __syncthreads();
    const int chain = a.chain_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = chain < a.chain_hi;
    int acc_i = 0, win_i = 0;
    if (active) {
        double q = 0.0;
        for (int r0 = 0; r0 < a.R; r0 += 16) {               // 16 loads in flight, summed in the fixed order r = 0, 1, ...
            double pr[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] = (r0 + j < a.R) ? a.partial[(size_t)(r0 + j) * a.Npad + chain] : 0.0;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (r0 + j < a.R) q += pr[j];
        }
        const double lpp = a.link ? (a.lin[chain] - q) : -0.5 * (a.ssr0 - 2.0 * a.lin[chain] + q);     // logistic: y.eta - sum softplus(eta)
        const double lp = a.lp[chain];
        if (a.block == 0 && a.lp_gen0) a.lp_gen0[chain] = lp;
        const double logu = log((double)uniform01(a.wacc[chain]));
        const double delta = lpp - lp;
        const double geo = a.geo ? (double)a.geo[chain] : 0.0;
        const bool acc = a.tempered ? (logu * (double)a.T < delta + (double)a.T * geo) : (logu < delta + geo);     // T = 0 -> greedy (Q4); NaN -> reject
        if (a.tr_lpprop) {
            const size_t o = ((size_t)(a.grel - 1) * a.nblocks + a.block) * a.N + chain;
            a.tr_lpprop[o] = (float)lpp;
            a.tr_acc[o] = acc ? 1 : 0;
        }
        double lpn = lp;
        if (acc) {
            for (int k0 = 0; k0 < a.d; k0 += 8) {             // loads of a group before its stores (see linreg_prepare_kernel)
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) if (k0 + j < a.d) t[j] = a.Xprop[(size_t)(k0 + j) * a.ldx + chain];
#pragma unroll
                for (int j = 0; j < 8; ++j) if (k0 + j < a.d) a.X[(size_t)(k0 + j) * a.ldx + chain] = t[j];
            }
            a.lp[chain] = lpp;
            lpn = lpp;
            acc_i = 1;
        }
        if (a.last_block) {
            const bool changed = lpn != (a.lp_gen0 ? a.lp_gen0[chain] : lp);
            if (changed && a.count_run) a.chain_changed[chain] += 1;
            win_i = (changed && a.count_win) ? 1 : 0;
            if (a.hist || a.Zw) {
                uint32_t row = a.append_row0 + a.chain_offset + (uint32_t)chain;
                if (row >= a.zcap) row -= a.zcap;
                const int LD4 = (a.d + 3) & ~3;
                for (int k0 = 0; k0 < a.d; k0 += 8) {
                    float t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = (k0 + j < a.d) ? a.X[(size_t)(k0 + j) * a.ldx + chain] : 0.f;
                    if (a.hist) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (k0 + j < a.d) a.hist[((size_t)a.hidx * a.d + k0 + j) * a.N + chain] = t[j];
                    }
                    if (a.Zw) {                                // rows are 16-byte aligned; the padding up to 4*ceil(d/4) is written as zeros
                        float4* dst = reinterpret_cast<float4*>(a.Zw + (size_t)row * a.ZS + k0);
                        dst[0] = make_float4(t[0], t[1], t[2], t[3]);
                        if (k0 + 4 < LD4) dst[1] = make_float4(t[4], t[5], t[6], t[7]);
                    }
                }
                if (a.hist) a.lphist[(size_t)a.hidx * a.N + chain] = lpn;
            }
        }
    }
    if (a.last_block && a.S1) {                         // split-Rhat accumulators (src/utils.jl:2-20)
        for (int k0 = 0; k0 < a.d; k0 += 8) {
            float t[8], s1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool on = active && k0 + j < a.d;
                t[j] = on ? a.X[(size_t)(k0 + j) * a.ldx + chain] - a.shift[k0 + j] : 0.f;
                s1[j] = on ? a.S1[(size_t)(k0 + j) * a.ldx + chain] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (k0 + j < a.d) {                            // warp-uniform
                    if (active) a.S1[(size_t)(k0 + j) * a.ldx + chain] = s1[j] + t[j];
                    const double v = warp_sum((double)t[j] * (double)t[j]);
                    if ((threadIdx.x & 31) == 0) atomicAdd(&red_s2[k0 + j], v);
                }
            }
        }
    }
    const int wa = warp_sum(acc_i), ww = warp_sum(win_i);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[0], (unsigned)wa); atomicAdd(&red[1], (unsigned)ww); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (red[0]) atomicAdd(a.accepted_total, (unsigned long long)red[0]);
        if (red[1]) atomicAdd(a.win_changed, (unsigned long long)red[1]);
    }
    if (a.last_block && a.S1 && threadIdx.x < a.d) atomicAdd(&a.S2[threadIdx.x], red_s2[threadIdx.x]);
}

__global__ void linreg_finish_eval_kernel(const double* __restrict__ partial, const double* __restrict__ lin, int n, int Npad,
                                          int R, double ssr0, int link, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q = 0.0;
    for (int r = 0; r < R; ++r) q += partial[(size_t)r * Npad + i];
    out[i] = link ? (lin[i] - q) : -0.5 * (ssr0 - 2.0 * lin[i] + q);
}

// ------------------------------------------------------------------------------------------------
// the tensor-core kernel:  partial[rc][chain] = sum over the rows of chunk rc of (x_row . d_chain)^2
// ------------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 8;          // two epilogue warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int kSsrThreads = 64 + 32 * kEpiWarps;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..: epilogue

template <int KB, int TILE_N, int STAGES>
struct SsrSmem {
    static constexpr int B_BYTES = KB * TILE_N * 128;              // one of X_hi / X_lo for one row tile
    static constexpr int STAGE_BYTES = 2 * B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;             // barriers + tmem holder + 1024-byte alignment slack
};

// TMEM columns: [0, 2 TILE_N) two accumulator stages | d_hi (KB*32 columns) | d_lo (KB*32 columns)
template <int KB, int TILE_N, int STAGES, int TERMS>
__global__ void __launch_bounds__(kSsrThreads, 1)
linreg_ssr_kernel(const __grid_constant__ CUtensorMap tmXhi, const __grid_constant__ CUtensorMap tmXlo,
                  const float* __restrict__ Dhi, const float* __restrict__ Dlo, int ppad, int ksteps, int tiles,
                  int tiles_per_chunk, int CT, int units, int npad, double* __restrict__ partial)
{
    using L = SsrSmem<KB, TILE_N, STAGES>;
    constexpr int ACC = 2;                                         // TMEM accumulator stages
    constexpr uint32_t A_COL = ACC * TILE_N;                       // first column of the chain tile
    constexpr uint32_t TMEM_COLS = 512;
    static_assert(A_COL + 2 * KB * 32 <= TMEM_COLS, "TMEM budget");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t* a_full = bars + 2 * STAGES;  // chain tile stored in TMEM
    uint64_t* tfull = a_full + 1;          // [ACC] MMA -> epilogue
    uint64_t* tempty = tfull + ACC;        // [ACC] epilogue -> MMA
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmXhi); tc::tma_prefetch_desc(&tmXlo);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(a_full, kEpiWarps);
        for (int s = 0; s < ACC; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], kEpiWarps); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_holder, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer: streams the X tiles of all this CTA's units =====================
        // (whole warp in the loop, one elected lane issues — see linreg_ssr_f16_kernel)
        int stage = 0; uint32_t ph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int rc = u / CT;
            const int t0 = rc * tiles_per_chunk, t1 = min(tiles, t0 + tiles_per_chunk);
            for (int t = t0; t < t1; ++t) {
                tc::mbar_wait(&empty[stage], ph ^ 1);
                uint8_t* bh = sB + stage * L::STAGE_BYTES;
                uint8_t* bl = bh + L::B_BYTES;
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&full[stage], L::STAGE_BYTES);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        tc::tma_load_2d(bh + kb * TILE_N * 128, &tmXhi, kb * 32, t * TILE_N, &full[stage]);
                        tc::tma_load_2d(bl + kb * TILE_N * 128, &tmXlo, kb * 32, t * TILE_N, &full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp in the loop, one elected lane issues) =====================
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, TILE_N);
        int stage = 0, as = 0; uint32_t ph = 0, aph = 0, tph = 0;
        const uint32_t a_hi = tmem_base + A_COL, a_lo = tmem_base + A_COL + KB * 32;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int rc = u / CT;
            tc::mbar_wait(a_full, aph);                        // the epilogue warps stored this unit's chain tile
            aph ^= 1;
            tc::tc_fence_after();
            const int t0 = rc * tiles_per_chunk, t1 = min(tiles, t0 + tiles_per_chunk);
            for (int t = t0; t < t1; ++t) {
                tc::mbar_wait(&tempty[as], tph ^ 1);
                tc::mbar_wait(&full[stage], ph);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_N);
                const uint64_t b_hi0 = tc::make_sw128_kmajor_desc(sB + stage * L::STAGE_BYTES);
                const uint64_t b_lo0 = tc::make_sw128_kmajor_desc(sB + stage * L::STAGE_BYTES + L::B_BYTES);
                if (tc::elect_one()) {
                    uint32_t accum = 0;
                    // a k-step is 8 TF32: 8 TMEM columns of the chain tile, 32 bytes inside the 128-byte swizzle row of X
                    // (descriptor start address + 2); a k-block is the next [rows][128 B] panel.  Cross terms first.
                    if (TERMS == 3) {
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t bo = (uint64_t)(((ks >> 2) * TILE_N * 128 + (ks & 3) * 32) >> 4);
                            tc::mma_tf32_ts(d_tmem, a_lo + ks * 8, b_hi0 + bo, idesc, accum); accum = 1;
                            tc::mma_tf32_ts(d_tmem, a_hi + ks * 8, b_lo0 + bo, idesc, accum);
                        }
                    }
#pragma unroll 4
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t bo = (uint64_t)(((ks >> 2) * TILE_N * 128 + (ks & 3) * 32) >> 4);
                        tc::mma_tf32_ts(d_tmem, a_hi + ks * 8, b_hi0 + bo, idesc, accum); accum = 1;
                    }
                    tc::mma_commit(&empty[stage]);             // X tile consumed -> producer may refill
                    tc::mma_commit(&tfull[as]);                // accumulator ready -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; ph ^= 1; }
                if (++as == ACC) { as = 0; tph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps: chain tile -> TMEM, then square-and-sum =====================
        constexpr int HALVES = kEpiWarps / 4;                  // column ranges per lane quarter
        constexpr int COLS = TILE_N / HALVES;
        static_assert(COLS % 32 == 0, "epilogue column split");
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                      // which column range of the tile
        int as = 0; uint32_t tph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int rc = u / CT, ct = u - rc * CT;
            // All MMAs of the previous unit are complete (its last accumulator was consumed below), so the chain tile
            // in TMEM can be replaced: lane = chain, columns = the KB*32 coefficients, hi by one warp of the quarter,
            // lo by the other.
            {
                const float* src = (half == 0 ? Dhi : Dlo) + (size_t)(ct * 128 + quarter * 32 + lane) * ppad;
                const uint32_t acol = A_COL + (half == 0 ? 0 : KB * 32);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(src + kb * 32 + j);
                        v[j] = t4.x; v[j + 1] = t4.y; v[j + 2] = t4.z; v[j + 3] = t4.w;
                    }
                    tc::tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acol + kb * 32, v);
                }
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(a_full);
            }
            const int t0 = rc * tiles_per_chunk, t1 = min(tiles, t0 + tiles_per_chunk);
            double accd = 0.0;
            for (int t = t0; t < t1; ++t) {
                tc::mbar_wait(&tfull[as], tph);
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TILE_N + half * COLS);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int c = 0; c < COLS; c += 32) {
                    float v[32];
                    tc::tmem_ld32(taddr + c, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        s0 = fmaf(v[j], v[j], s0); s1 = fmaf(v[j + 1], v[j + 1], s1);
                        s2 = fmaf(v[j + 2], v[j + 2], s2); s3 = fmaf(v[j + 3], v[j + 3], s3);
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[as]);
                accd += (double)((s0 + s1) + (s2 + s3));
                if (++as == ACC) { as = 0; tph ^= 1; }
            }
            partial[(size_t)(rc * HALVES + half) * npad + ct * 128 + quarter * 32 + lane] = accd;
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// the same contraction on kind::f16 with the 2-way FP16 split.  Per data row the B operand is ONE row of
// [hi(PPAD) | lo(PPAD)] halves (4 PPAD bytes: PPAD/32 boxes of 128 bytes), the chain tile in TMEM is
// [hi pairs | lo pairs] (PPAD 32-bit columns).  A K step is 16 halves = 32 bytes of the swizzled row = 8 TMEM columns,
// so the descriptor arithmetic is that of the TF32 kernel; half as many MMAs and half the X bytes per product.
// ------------------------------------------------------------------------------------------------
template <int PPAD, int TILE_N, int STAGES>
struct SsrSmemH {
    static constexpr int KBB = PPAD / 32;                          // 128-byte boxes per data row: hi|lo of 32, or hi, lo of 64
    static constexpr int STAGE_BYTES = KBB * TILE_N * 128;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

// TMEM columns: [0, 2 TILE_N) two accumulator stages | d_hi (PPAD/2 columns) | d_lo (PPAD/2 columns)
// LINK selects the per-element function the epilogue sums over the data rows of (X b)[row, chain]:
//   0: v^2                  Gaussian linear regression (SSR)
//   1: softplus(eta)        Bernoulli-logit GLM: sum_i log(1 + exp(eta_i)); the y . eta part of the log-likelihood is the
//                           FP64 linear term b . (X^T y) of the prepare kernel.  MUFU ex2 + lg2 per element.
template <int PPAD, int TILE_N, int STAGES, int TERMS, int LINK>
__global__ void __launch_bounds__(kSsrThreads, 1)
linreg_ssr_f16_kernel(const __grid_constant__ CUtensorMap tmXh, const uint32_t* __restrict__ Dh, const int* __restrict__ Dexp,
                      int ksteps, int tiles, int tiles_per_chunk, int CT, int units, int npad, double* __restrict__ partial, int n_rows)
{
    using L = SsrSmemH<PPAD, TILE_N, STAGES>;
    constexpr int KBB = L::KBB;
    constexpr int ACC = 2;
    constexpr uint32_t A_COL = ACC * TILE_N;
    constexpr uint32_t TMEM_COLS = 512;
    static_assert(A_COL + PPAD <= TMEM_COLS, "TMEM budget");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* a_full = bars + 2 * STAGES;
    uint64_t* tfull = a_full + 1;
    uint64_t* tempty = tfull + ACC;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmXh);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(a_full, kEpiWarps);
        for (int s = 0; s < ACC; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], kEpiWarps); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_holder, TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The WHOLE warp runs the loop (warp-uniform control flow and addresses, so the compiler keeps them in uniform
        // registers); one elected lane issues.  A loop under `if (lane == 0)` instead makes ptxas wrap every TMA / MMA
        // instruction in an ELECT ... BRA.U.ANY serialisation loop (~150 cycles per instruction: measured, the MMA
        // issue rate — not the tensor pipe — bounded the kernel).
        int stage = 0; uint32_t ph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int rc = u / CT;
            const int t0 = rc * tiles_per_chunk, t1 = min(tiles, t0 + tiles_per_chunk);
            for (int t = t0; t < t1; ++t) {
                tc::mbar_wait(&empty[stage], ph ^ 1);
                uint8_t* b = sB + stage * L::STAGE_BYTES;
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&full[stage], L::STAGE_BYTES);
#pragma unroll
                    for (int kb = 0; kb < KBB; ++kb) tc::tma_load_2d(b + kb * TILE_N * 128, &tmXh, kb * 64, t * TILE_N, &full[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp in the loop, one elected lane issues) =====================
        constexpr uint32_t idesc = tc::make_idesc_f16(128, TILE_N);
        int stage = 0, as = 0; uint32_t ph = 0, aph = 0, tph = 0;
        const uint32_t a_hi = tmem_base + A_COL, a_lo = tmem_base + A_COL + PPAD / 2;
        // descriptor offset (16-byte units) of K step ks of the hi (part 0) / lo (part 1) half of a data row
        auto boff = [](int part, int ks) -> uint64_t {
            const int e = part * PPAD + ks * 16;                           // first half-element of the step inside the row
            return (uint64_t)(((e >> 6) * TILE_N * 128 + (e & 63) * 2) >> 4);
        };
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int rc = u / CT;
            tc::mbar_wait(a_full, aph);
            aph ^= 1;
            tc::tc_fence_after();
            const int t0 = rc * tiles_per_chunk, t1 = min(tiles, t0 + tiles_per_chunk);
            for (int t = t0; t < t1; ++t) {
                tc::mbar_wait(&tempty[as], tph ^ 1);
                tc::mbar_wait(&full[stage], ph);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_N);
                const uint64_t b0 = tc::make_sw128_kmajor_desc(sB + stage * L::STAGE_BYTES);
                if (tc::elect_one()) {
                    uint32_t accum = 0;
                    if (TERMS == 3) {                                      // cross terms first
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) {
                            tc::mma_f16_ts(d_tmem, a_lo + ks * 8, b0 + boff(0, ks), idesc, accum); accum = 1;
                            tc::mma_f16_ts(d_tmem, a_hi + ks * 8, b0 + boff(1, ks), idesc, accum);
                        }
                    }
#pragma unroll 4
                    for (int ks = 0; ks < ksteps; ++ks) {
                        tc::mma_f16_ts(d_tmem, a_hi + ks * 8, b0 + boff(0, ks), idesc, accum); accum = 1;
                    }
                    tc::mma_commit(&empty[stage]);
                    tc::mma_commit(&tfull[as]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; ph ^= 1; }
                if (++as == ACC) { as = 0; tph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        constexpr int HALVES = kEpiWarps / 4;
        constexpr int COLS = TILE_N / HALVES;
        static_assert(COLS % 32 == 0, "epilogue column split");
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        int as = 0; uint32_t tph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int rc = u / CT, ct = u - rc * CT;
            const int chain = ct * 128 + quarter * 32 + lane;
            {
                // chain tile -> TMEM: PPAD words per chain, 32 columns per store; PPAD = 32: one store (first warp of the
                // quarter), PPAD = 64: hi words by one warp, lo words by the other
                if (half < PPAD / 32) {
                    const uint32_t* src = Dh + (size_t)chain * PPAD + half * 32;
                    uint32_t v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const uint4 t4 = *reinterpret_cast<const uint4*>(src + j);
                        v[j] = t4.x; v[j + 1] = t4.y; v[j + 2] = t4.z; v[j + 3] = t4.w;
                    }
                    tc::tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + A_COL + half * 32, v);
                    tc::tmem_st_wait();
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(a_full);
            }
            const int fexp = Dexp[chain];
            const float sc = __int_as_float((127 - fexp) << 23);          // 2^-f (|f| <= 100)
            const int t0 = rc * tiles_per_chunk, t1 = min(tiles, t0 + tiles_per_chunk);
            double accd = 0.0;
            for (int t = t0; t < t1; ++t) {
                tc::mbar_wait(&tfull[as], tph);
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TILE_N + half * COLS);
                // all of the warp's columns in flight at once, ONE wait, then the accumulator stage goes back to the MMA
                // issuer before the arithmetic starts (the values are in registers)
                float v[COLS];
#pragma unroll
                for (int c = 0; c < COLS; c += 32) tc::tmem_ld32(taddr + c, *reinterpret_cast<float(*)[32]>(&v[c]));
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[as]);
                if constexpr (LINK == 0) {
                    float2 q[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int j = 0; j < COLS; j += 8) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 t2 = make_float2(v[j + 2 * e], v[j + 2 * e + 1]);
                            q[e] = ffma2(t2, t2, q[e]);
                        }
                    }
                    accd += (double)(((q[0].x + q[0].y) + (q[1].x + q[1].y)) + ((q[2].x + q[2].y) + (q[3].x + q[3].y)));
                } else {
                    // softplus(eta) = max(eta, 0) + log1p(exp(-|eta|)), eta = accumulator * 2^-f; rows beyond n (zero-filled by
                    // TMA, eta = 0 -> ln 2) are masked in the last tile
                    const int valid = n_rows - (t * TILE_N + half * COLS);
                    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
                    auto sp = [&](float a) {
                        const float eta = a * sc;
                        float tt, l2;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(tt) : "f"(-1.4426950408889634f * fabsf(eta)));
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(1.f + tt));
                        return fmaf(l2, 0.6931471805599453f, fmaxf(eta, 0.f));
                    };
                    if (valid >= COLS) {
#pragma unroll
                        for (int j = 0; j < COLS; j += 4) { q0 += sp(v[j]); q1 += sp(v[j + 1]); q2 += sp(v[j + 2]); q3 += sp(v[j + 3]); }
                    } else {
#pragma unroll
                        for (int j = 0; j < COLS; j += 4) {
                            q0 += j < valid ? sp(v[j]) : 0.f; q1 += j + 1 < valid ? sp(v[j + 1]) : 0.f;
                            q2 += j + 2 < valid ? sp(v[j + 2]) : 0.f; q3 += j + 3 < valid ? sp(v[j + 3]) : 0.f;
                        }
                    }
                    accd += (double)((q0 + q1) + (q2 + q3));
                }
                if (++as == ACC) { as = 0; tph ^= 1; }
            }
            // LINK 0: undo the chain's power-of-two scale (the accumulators hold (2^f x.d)^2 sums)
            partial[(size_t)(rc * HALVES + half) * npad + chain] = LINK == 0 ? scalbn(accd, -2 * fexp) : accd;
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major FP32 [rows][ppad] -> boxes of [box_rows][32 floats = 128 B], 128-byte swizzle, OOB rows read as zero
int make_map(demcz_ctx* c, CUtensorMap* m, const float* base, uint64_t rows, int ppad, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return demcz_fail(c, DEMCZ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)ppad, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ppad * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return demcz_fail(c, DEMCZ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return DEMCZ_OK;
}

// row-major FP16 [rows][2 ppad] -> boxes of [box_rows][64 halves = 128 B], 128-byte swizzle, OOB rows read as zero
int make_map_f16(demcz_ctx* c, CUtensorMap* m, const __half* base, uint64_t rows, int ppad, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return demcz_fail(c, DEMCZ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)(2 * ppad), (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)(2 * ppad) * sizeof(__half)};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return demcz_fail(c, DEMCZ_ERR_CUDA, "cuTensorMapEncodeTiled (FP16) failed with CUresult %d", (int)r);
    return DEMCZ_OK;
}

template <int PPAD, int TILE_N, int STAGES, int TERMS, int LINK>
cudaError_t launch_ssr_f16_t(demcz_ctx* c, demcz_linreg_state* s, const CUtensorMap& tm, const uint32_t* dh, const int* dexp, int tiles,
                             int tiles_per_chunk, int R, int CT, int npad, double* partial, int n_rows, cudaStream_t st)
{
    auto kern = linreg_ssr_f16_kernel<PPAD, TILE_N, STAGES, TERMS, LINK>;
    constexpr int smem = SsrSmemH<PPAD, TILE_N, STAGES>::TOTAL;
    static bool configured[64] = {false};
    const int dev = c->device & 63;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int units = CT * R;
    const int grid = std::min(units, s->sms);
    kern<<<grid, kSsrThreads, smem, st>>>(tm, dh, dexp, s->ksteps16, tiles, tiles_per_chunk, CT, units, npad, partial, n_rows);
    return cudaGetLastError();
}

constexpr int kTileNH = 192;          // data rows per MMA of the FP16 kernel: 2 N + PPAD <= 512 TMEM columns for PPAD = 32 and 64

cudaError_t launch_ssr_f16(demcz_ctx* c, demcz_linreg_state* s, const CUtensorMap& tm, const uint32_t* dh, const int* dexp, int tiles,
                           int tiles_per_chunk, int R, int CT, int npad, double* partial, int n_rows = 1 << 30, int link = -1,
                           cudaStream_t st = nullptr)
{
    if (link < 0) link = s->link;
    if (!st) st = c->stream;
#define DEMCZ_F16_CASE(PP, ST, TE, LK) return launch_ssr_f16_t<PP, kTileNH, ST, TE, LK>(c, s, tm, dh, dexp, tiles, tiles_per_chunk, R, CT, npad, partial, n_rows, st)
    if (s->ppad == 32) {
        if (link == 1) { if (s->terms == 3) DEMCZ_F16_CASE(32, 6, 3, 1); DEMCZ_F16_CASE(32, 6, 1, 1); }
        if (s->terms == 3) DEMCZ_F16_CASE(32, 6, 3, 0);
        DEMCZ_F16_CASE(32, 6, 1, 0);
    }
    if (link == 1) { if (s->terms == 3) DEMCZ_F16_CASE(64, 4, 3, 1); DEMCZ_F16_CASE(64, 4, 1, 1); }
    if (s->terms == 3) DEMCZ_F16_CASE(64, 4, 3, 0);
    DEMCZ_F16_CASE(64, 4, 1, 0);
#undef DEMCZ_F16_CASE
}

template <int KB, int TILE_N, int STAGES, int TERMS>
cudaError_t launch_ssr_t(demcz_ctx* c, demcz_linreg_state* s, const float* dhi, const float* dlo, int CT, int npad, double* partial, cudaStream_t st)
{
    auto kern = linreg_ssr_kernel<KB, TILE_N, STAGES, TERMS>;
    constexpr int smem = SsrSmem<KB, TILE_N, STAGES>::TOTAL;
    static bool configured[64] = {false};       // per instantiation and per device (function attributes are per device)
    const int dev = c->device & 63;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int units = CT * s->R;
    const int grid = std::min(units, s->sms);
    kern<<<grid, kSsrThreads, smem, st>>>(s->tmXhi, s->tmXlo, dhi, dlo, s->ppad, s->ksteps, s->tiles, s->tiles_per_chunk, CT,
                                                 units, npad, partial);
    return cudaGetLastError();
}

constexpr int kTileN32 = 192, kTileN64 = 128;     // data rows per MMA (N) for p <= 32 / p <= 64: 2 N + chain-tile columns <= 512 TMEM columns

cudaError_t launch_ssr(demcz_ctx* c, demcz_linreg_state* s, const float* dhi, const float* dlo, int CT, int npad, double* partial,
                       const uint32_t* dh = nullptr, const int* dexp = nullptr, cudaStream_t st = nullptr)
{
    if (!st) st = c->stream;
    if (s->kind == 1)
        return launch_ssr_f16(c, s, s->tmXh, dh ? dh : s->Dh, dexp ? dexp : s->Dexp, s->tiles, s->tiles_per_chunk, s->R, CT, npad, partial, (int)s->n, -1, st);
    if (s->ppad == 32) {
        if (s->terms == 3) return launch_ssr_t<1, kTileN32, 4, 3>(c, s, dhi, dlo, CT, npad, partial, st);
        return launch_ssr_t<1, kTileN32, 4, 1>(c, s, dhi, dlo, CT, npad, partial, st);
    }
    if (s->terms == 3) return launch_ssr_t<2, kTileN64, 3, 3>(c, s, dhi, dlo, CT, npad, partial, st);
    return launch_ssr_t<2, kTileN64, 3, 1>(c, s, dhi, dlo, CT, npad, partial, st);
}

inline int grid1(long long n, int t = 256) { return (int)((n + t - 1) / t); }

}  // namespace

// ------------------------------------------------------------------------------------------------
// registration
// ------------------------------------------------------------------------------------------------
void demcz_linreg_free(demcz_ctx* c)
{
    demcz_linreg_state* s = c->linreg;
    if (!s) return;
    void* ptrs[] = {s->Xhi, s->Xlo, s->d_b0, s->d_g, s->Dhi, s->Dlo, s->Xprop, s->lin, s->partial, s->lp64, s->lphist64, s->wacc,
                    s->Xh, s->Dh, s->Dexp, s->geo};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    for (int h = 0; h < 2; ++h) { if (s->hstream[h]) cudaStreamDestroy(s->hstream[h]); if (s->ev_join[h]) cudaEventDestroy(s->ev_join[h]); }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    delete s;
    c->linreg = nullptr;
}

static int linreg_register(demcz_ctx* c, demcz_linreg_state* s, int64_t n, int32_t p, const double* X, const double* y);

static void linreg_init_geometry(demcz_linreg_state* s, int64_t n, int32_t p)
{
    s->n = n; s->p = p; s->ppad = p <= 32 ? 32 : 64; s->ksteps = (p + 7) / 8; s->ksteps16 = (p + 15) / 16;
}

// One launch of the FP16 kernel on a hand-made tile whose result tells how the hardware reads the packed FP16 chain
// tile from tensor memory: the even K element of a pair in the low half-word of the 32-bit column (swap_halves = 0) or
// in the high one (1).  Anything else -> DEMCZ_ERR_UNSUPPORTED and the caller falls back to the 3xTF32 kernel.
static int linreg_f16_selftest(demcz_ctx* c, demcz_linreg_state* s)
{
    const int ppad = s->ppad, W = 2 * ppad, TN = kTileNH;
    std::vector<__half> hx((size_t)TN * W, __float2half(0.f));
    auto setx = [&](int col, float v) { for (int r = 0; r < TN; ++r) hx[(size_t)r * W + col] = __float2half(v); };
    setx(0, 1.f); setx(1, 2.f); setx(16, 3.f); setx(17, 5.f);
    if (ppad == 64) { setx(32, 7.f); setx(33, 13.f); setx(48, 11.f); setx(49, 17.f); }
    setx(ppad + 0, 0.5f);                                           // X_lo[:, 0]
    std::vector<uint32_t> hd((size_t)128 * ppad, 0u);
    std::vector<int> hexp(128, 0);
    auto one = [](float v) { return (uint32_t)__half_as_ushort(__float2half(v)); };
    const int HW = ppad / 2;
    hd[0 * ppad + 0] = one(1.f);                                    // chain 0: d_hi[0] = 1
    hd[1 * ppad + 8] = one(1.f);                                    // chain 1: d_hi[16] = 1 (second K step)
    hd[2 * ppad + HW + 0] = one(1.f);                               // chain 2: d_lo[0] = 1
    hd[3 * ppad + 0] = one(4.f); hexp[3] = 2;                       // chain 3: d_hi[0] = 4 at chain exponent 2
    if (ppad == 64) { hd[4 * ppad + 16] = one(1.f); hd[5 * ppad + 24] = one(1.f); }     // d_hi[32], d_hi[48]
    const double nat[6] = {2.25, 9.0, 1.0, 2.25, 49.0, 121.0}, swp[6] = {4.0, 25.0, 4.0, 4.0, 169.0, 289.0};
    const int ncheck = ppad == 64 ? 6 : 4;

    __half* dx = nullptr; uint32_t* dd = nullptr; int* de = nullptr; double* dp = nullptr;
    CUL(c, cudaMalloc((void**)&dx, hx.size() * sizeof(__half)));
    cudaError_t e = cudaMalloc((void**)&dd, hd.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&de, 128 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dp, 2 * 128 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx, hx.data(), hx.size() * sizeof(__half), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dd, hd.data(), hd.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(de, hexp.data(), 128 * sizeof(int), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dp, 0, 2 * 128 * sizeof(double), c->stream);
    int rc = DEMCZ_OK;
    CUtensorMap tm;
    if (e == cudaSuccess) rc = make_map_f16(c, &tm, dx, (uint64_t)TN, ppad, TN);
    double hp[256] = {0};
    if (e == cudaSuccess && rc == DEMCZ_OK) {
        const int ks_save = s->ksteps16;
        s->ksteps16 = ppad / 16;
        e = launch_ssr_f16(c, s, tm, dd, de, 1, 1, 1, 1, 128, dp, 1 << 30, 0);
        s->ksteps16 = ks_save;
        if (e == cudaSuccess) e = cudaMemcpyAsync(hp, dp, sizeof hp, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(dx); if (dd) cudaFree(dd); if (de) cudaFree(de); if (dp) cudaFree(dp);
    if (rc != DEMCZ_OK) return rc;
    CUL(c, e);
    bool is_nat = true, is_swp = true;
    char got[256]; int off = 0;
    for (int i = 0; i < ncheck; ++i) {
        const double v = (hp[i] + hp[128 + i]) / TN;              // two column halves of the tile, TN identical rows
        off += snprintf(got + off, sizeof got - off, " %.4g", v);
        if (fabs(v - nat[i]) > 1e-6 * nat[i]) is_nat = false;
        if (fabs(v - swp[i]) > 1e-6 * swp[i]) is_swp = false;
    }
    if (const char* dbg = getenv("DEMCZ_DEBUG")) if (atoi(dbg)) fprintf(stderr, "demcz: FP16 TMEM layout self-test:%s (natural %d, swapped %d)\n", got, is_nat, is_swp);
    if (is_nat) { s->swap_halves = 0; return DEMCZ_OK; }
    if (is_swp) { s->swap_halves = 1; return DEMCZ_OK; }
    demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "FP16 tensor-memory operand self-test failed (per-row sums:%s); using the 3xTF32 kernel", got);
    c->sticky = 0;
    return DEMCZ_ERR_UNSUPPORTED;
}

static int set_likelihood_xb(demcz_ctx* c, int64_t n, int32_t p, const double* X, const double* y, int link);

extern "C" int demcz_set_likelihood_linreg(demcz_ctx* c, int64_t n, int32_t p, const double* X, const double* y)
{
    return set_likelihood_xb(c, n, p, X, y, 0);
}

extern "C" int demcz_set_likelihood_logistic(demcz_ctx* c, int64_t n, int32_t p, const double* X, const double* y)
{
    return set_likelihood_xb(c, n, p, X, y, 1);
}

static int set_likelihood_xb(demcz_ctx* c, int64_t n, int32_t p, const double* X, const double* y, int link)
{
    if (!c) return DEMCZ_ERR_INVALID;
    if (c->sticky) return c->sticky;
    if (cudaSetDevice(c->device) != cudaSuccess) return demcz_fail(c, DEMCZ_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
    if (!X || !y || n < 1 || p != c->o.d) return demcz_fail(c, DEMCZ_ERR_INVALID, "set_likelihood_linreg: n=%lld p=%d, opts.d=%d", (long long)n, p, c->o.d);
    if (p > DEMCZ_MAX_DIM_LINREG) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "set_likelihood_linreg: p=%d > %d", p, DEMCZ_MAX_DIM_LINREG);
    if (n >= ((int64_t)1 << 31)) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "set_likelihood_linreg: n >= 2^31 rows");
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, c->device);
    if (major != 10) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "the linreg likelihood needs tcgen05 (sm_100a); device has compute capability %d.x", major);
    demcz_linreg_free(c);
    demcz_linreg_state* s = new demcz_linreg_state();
    c->linreg = s;
    linreg_init_geometry(s, n, p);
    s->link = link;
    if (const char* kd = getenv("DEMCZ_LINREG_KIND")) s->kind = (strcmp(kd, "tf32") == 0) ? 0 : 1;
    if (link == 1 && s->kind != 1) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "the logistic likelihood runs on the FP16 tcgen05 kernel only (DEMCZ_LINREG_KIND=tf32 is set)");
    return linreg_register(c, s, n, p, X, y);
}

// (re-)builds everything that depends on the MMA kind; called again with kind 0 if the FP16 self-test fails
static int linreg_register(demcz_ctx* c, demcz_linreg_state* s, int64_t n, int32_t p, const double* X, const double* y)
{
    const demcz_opts& o = c->o;
    s->tile_n = s->kind == 1 ? kTileNH : (s->ppad == 32 ? kTileN32 : kTileN64);
    s->tiles = (int)((n + s->tile_n - 1) / s->tile_n);
    s->R = std::max(1, std::min(37, s->tiles / 8));                 // row chunks: 37 | 148 keeps (chain tiles x chunks) balanced over the SMs
    s->tiles_per_chunk = (s->tiles + s->R - 1) / s->R;
    s->R = (s->tiles + s->tiles_per_chunk - 1) / s->tiles_per_chunk;
    s->Npad = (o.N + 127) / 128 * 128;
    if (const char* t = getenv("DEMCZ_LINREG_TERMS")) s->terms = atoi(t) == 1 ? 1 : 3;
    cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, c->device);

    const size_t np = (size_t)n * s->ppad;
    double *dX = nullptr, *dy = nullptr, *dscr = nullptr;
    if (s->kind == 1) {
        CUL(c, cudaMalloc((void**)&s->Xh, 2 * np * sizeof(__half)));
        CUL(c, cudaMalloc((void**)&s->Dh, (size_t)s->Npad * s->ppad * sizeof(uint32_t)));
        CUL(c, cudaMalloc((void**)&s->Dexp, (size_t)s->Npad * sizeof(int)));
        CUL(c, cudaMemsetAsync(s->Dh, 0, (size_t)s->Npad * s->ppad * sizeof(uint32_t), c->stream));
        CUL(c, cudaMemsetAsync(s->Dexp, 0, (size_t)s->Npad * sizeof(int), c->stream));
    } else {
        CUL(c, cudaMalloc((void**)&s->Xhi, np * sizeof(float)));
        CUL(c, cudaMalloc((void**)&s->Xlo, np * sizeof(float)));
    }
    CUL(c, cudaMalloc((void**)&s->d_b0, 64 * sizeof(double)));
    CUL(c, cudaMalloc((void**)&s->d_g, 64 * sizeof(double)));
    if (s->kind == 0) {
        CUL(c, cudaMalloc((void**)&s->Dhi, (size_t)s->Npad * s->ppad * sizeof(float)));
        CUL(c, cudaMalloc((void**)&s->Dlo, (size_t)s->Npad * s->ppad * sizeof(float)));
        CUL(c, cudaMemsetAsync(s->Dhi, 0, (size_t)s->Npad * s->ppad * sizeof(float), c->stream));
        CUL(c, cudaMemsetAsync(s->Dlo, 0, (size_t)s->Npad * s->ppad * sizeof(float), c->stream));
    }
    CUL(c, cudaMalloc((void**)&s->Xprop, (size_t)c->ldx * o.d * sizeof(float)));
    CUL(c, cudaMalloc((void**)&s->lin, (size_t)s->Npad * sizeof(double)));
    CUL(c, cudaMalloc((void**)&s->partial, (size_t)s->R * (kEpiWarps / 4) * s->Npad * sizeof(double)));
    CUL(c, cudaMalloc((void**)&s->lp64, (size_t)c->ldx * 2 * sizeof(double)));      // [lp | lp at generation start]
    CUL(c, cudaMalloc((void**)&s->wacc, (size_t)c->ldx * sizeof(uint32_t)));
    CUL(c, cudaMalloc((void**)&s->geo, (size_t)c->ldx * sizeof(float)));
    CUL(c, cudaMemsetAsync(s->lp64, 0, (size_t)c->ldx * 2 * sizeof(double), c->stream));
    if (o.store_chain == 1 && o.Ngeneration / c->thin > 0) CUL(c, cudaMalloc((void**)&s->lphist64, (size_t)(o.Ngeneration / c->thin) * o.N * sizeof(double)));

    // ---- upload (FP64, as the caller holds it), split, OLS reference point -------------------------
    const int P1 = p + 1, npairs = P1 * P1;
    const long long rpb = std::max<long long>(32, (n + 295) / 296);
    const int gblocks = (int)((n + rpb - 1) / rpb);
    const int rblocks = std::min(grid1(n), 592);
    const size_t scr = std::max((size_t)gblocks * npairs, (size_t)rblocks * P1);
    CUL(c, cudaMalloc((void**)&dX, (size_t)n * p * sizeof(double)));
    cudaError_t e = cudaMalloc((void**)&dy, (size_t)n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dscr, scr * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dX, X, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    std::vector<double> hbuf(scr), G((size_t)npairs, 0.0);
    if (e == cudaSuccess && s->kind == 1) {
        // one power-of-two exponent per regressor column: max |X[:,k]| 2^xexp[k] in [2^14, 2^15)
        int* dmax = reinterpret_cast<int*>(dscr);
        e = cudaMemsetAsync(dmax, 0, 64 * sizeof(int), c->stream);
        if (e == cudaSuccess) {
            colabsmax_kernel<<<dim3(std::min(grid1(n), 592), p), 256, 0, c->stream>>>(dX, n, p, dmax);
            e = cudaGetLastError();
        }
        int hmax[64] = {0};
        if (e == cudaSuccess) e = cudaMemcpyAsync(hmax, dmax, (size_t)p * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        XExp xe;
        for (int k = 0; k < 64; ++k) {
            float m; memcpy(&m, &hmax[k], sizeof m);
            int ex = 0;
            if (k < p && m > 0.f && std::isfinite(m)) frexpf(m, &ex); else ex = 15;
            s->xexp[k] = xe.e[k] = 15 - ex;
        }
        if (e == cudaSuccess) {
            split_x_f16_kernel<<<grid1(n), 256, 0, c->stream>>>(dX, n, p, s->ppad, xe, s->Xh);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) {
        if (s->kind == 0) split_x_kernel<<<grid1(n), 256, 0, c->stream>>>(dX, n, p, s->ppad, s->Xhi, s->Xlo);
        gram_kernel<<<gblocks, 256, 32 * P1 * sizeof(double), c->stream>>>(dX, dy, n, p, rpb, dscr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hbuf.data(), dscr, (size_t)gblocks * npairs * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    std::vector<double> b0(64, 0.0);
    if (e == cudaSuccess) {
        for (int b = 0; b < gblocks; ++b)
            for (int t = 0; t < npairs; ++t) G[t] += hbuf[(size_t)b * npairs + t];
        // OLS by Cholesky of X^T X (host, FP64); a singular Gram matrix falls back to b0 = 0 (still exact, just
        // less well conditioned)
        std::vector<double> Lc((size_t)p * p, 0.0);
        bool pd = s->link == 0;                        // the GLM needs eta = X b itself: no centring, b0 = 0 and g = X^T y
        for (int j = 0; j < p && pd; ++j) {
            double d = G[(size_t)j * P1 + j];
            for (int k = 0; k < j; ++k) d -= Lc[(size_t)j * p + k] * Lc[(size_t)j * p + k];
            if (!(d > 1e-12 * G[(size_t)j * P1 + j]) || !(d > 0.0)) { pd = false; break; }
            const double ljj = sqrt(d);
            Lc[(size_t)j * p + j] = ljj;
            for (int i = j + 1; i < p; ++i) {
                double t = G[(size_t)i * P1 + j];
                for (int k = 0; k < j; ++k) t -= Lc[(size_t)i * p + k] * Lc[(size_t)j * p + k];
                Lc[(size_t)i * p + j] = t / ljj;
            }
        }
        if (pd) {
            std::vector<double> z(p);
            for (int i = 0; i < p; ++i) {
                double t = G[(size_t)i * P1 + p];
                for (int k = 0; k < i; ++k) t -= Lc[(size_t)i * p + k] * z[k];
                z[i] = t / Lc[(size_t)i * p + i];
            }
            for (int i = p - 1; i >= 0; --i) {
                double t = z[i];
                for (int k = i + 1; k < p; ++k) t -= Lc[(size_t)k * p + i] * b0[k];
                b0[i] = t / Lc[(size_t)i * p + i];
            }
        }
        e = cudaMemcpyAsync(s->d_b0, b0.data(), 64 * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    }
    if (e == cudaSuccess) {
        resid_kernel<<<rblocks, 256, 0, c->stream>>>(dX, dy, n, p, s->d_b0, dscr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hbuf.data(), dscr, (size_t)rblocks * P1 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) {
        std::vector<double> g(64, 0.0);
        double ssr0 = 0.0;
        for (int b = 0; b < rblocks; ++b) {
            for (int k = 0; k < p; ++k) g[k] += hbuf[(size_t)b * P1 + k];
            ssr0 += hbuf[(size_t)b * P1 + p];
        }
        s->ssr0 = ssr0;
        e = cudaMemcpy(s->d_g, g.data(), 64 * sizeof(double), cudaMemcpyHostToDevice);
    }
    cudaFree(dX);
    if (dy) cudaFree(dy);
    if (dscr) cudaFree(dscr);
    CUL(c, e);

    int rc = DEMCZ_OK;
    if (s->kind == 1) {
        rc = make_map_f16(c, &s->tmXh, s->Xh, (uint64_t)n, s->ppad, s->tile_n);
        if (rc == DEMCZ_OK) rc = linreg_f16_selftest(c, s);
        if (rc == DEMCZ_ERR_UNSUPPORTED) {
            // the FP16 operand layout this library assumes is not what the hardware reads: use the 3xTF32 kernel
            const std::string why = c->err;
            if (s->link == 1) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "logistic likelihood: %s", why.c_str());
            demcz_linreg_free(c);
            demcz_linreg_state* s2 = new demcz_linreg_state();
            c->linreg = s2;
            linreg_init_geometry(s2, n, p);
            s2->kind = 0;
            s2->fallback_reason = why;
            return linreg_register(c, s2, n, p, X, y);
        }
    } else {
        rc = make_map(c, &s->tmXhi, s->Xhi, (uint64_t)n, s->ppad, s->tile_n);
        if (rc == DEMCZ_OK) rc = make_map(c, &s->tmXlo, s->Xlo, (uint64_t)n, s->ppad, s->tile_n);
    }
    if (rc != DEMCZ_OK) return rc;
    c->lik_kind = DEMCZ_LIK_LINREG;
    c->state_lp_valid = false;
    return DEMCZ_OK;
}

// ------------------------------------------------------------------------------------------------
// batched evaluation: map(logobj, rows) (src/demcz.jl:17)
// ------------------------------------------------------------------------------------------------
int demcz_linreg_eval(demcz_ctx* c, const float* X, int n, int ldx, double* out_dev)
{
    demcz_linreg_state* s = c->linreg;
    if (!s) return demcz_fail(c, DEMCZ_ERR_INVALID, "linreg likelihood not set");
    const bool own_state = (out_dev == nullptr);           // the context's own chains -> lp64
    const int npad = (n + 127) / 128 * 128;
    float *dhi = s->Dhi, *dlo = s->Dlo; double *lin = s->lin, *partial = s->partial;
    uint32_t* dh = s->Dh; int* dexp = s->Dexp;
    void* tmp = nullptr;
    if (npad > s->Npad) {                                   // evaluation of more points than chains: temporaries
        const size_t fb = (size_t)npad * s->ppad * sizeof(float);
        const size_t total = 2 * fb + (size_t)npad * sizeof(double) * (1 + s->R * (kEpiWarps / 4));
        CUL(c, cudaMalloc(&tmp, total));
        CUL(c, cudaMemsetAsync(tmp, 0, total, c->stream));
        dhi = (float*)tmp; dlo = (float*)((char*)tmp + fb);
        dh = (uint32_t*)tmp; dexp = (int*)((char*)tmp + fb);       // FP16 kind: packed words + exponents in the same space
        lin = (double*)((char*)tmp + 2 * fb); partial = lin + npad;
    }
    LinArgs a;
    memset(&a, 0, sizeof a);
    a.X = const_cast<float*>(X); a.N = n; a.ldx = ldx; a.d = c->o.d; a.ppad = s->ppad;
    a.chain_lo = 0; a.chain_hi = n;
    a.Dhi = dhi; a.Dlo = dlo; a.lin = lin; a.b0 = s->d_b0; a.g = s->d_g;
    a.kind = s->kind; a.Dh = dh; a.Dexp = dexp; memcpy(a.xexp, s->xexp, sizeof a.xexp); a.swap_halves = s->swap_halves;
    linreg_prepare_kernel<<<grid1(n, 128), 128, 0, c->stream>>>(a, false);
    cudaError_t e = cudaGetLastError();
    const int use_npad = npad > s->Npad ? npad : s->Npad;
    if (e == cudaSuccess) e = launch_ssr(c, s, dhi, dlo, npad / 128, use_npad, partial, dh, dexp);
    if (e == cudaSuccess) {
        linreg_finish_eval_kernel<<<grid1(n), 256, 0, c->stream>>>(partial, lin, n, use_npad, s->R * (kEpiWarps / 4), s->ssr0, s->link, own_state ? s->lp64 : out_dev);
        e = cudaGetLastError();
    }
    s->ssr_launches++;
    if (tmp) { if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); cudaFree(tmp); }
    CUL(c, e);
    return DEMCZ_OK;
}

double* demcz_linreg_lp(demcz_ctx* c) { return c->linreg ? c->linreg->lp64 : nullptr; }
double* demcz_linreg_lphist(demcz_ctx* c) { return c->linreg ? c->linreg->lphist64 : nullptr; }
unsigned long long demcz_linreg_launches(demcz_ctx* c) { return c->linreg ? c->linreg->ssr_launches : 0; }

std::string demcz_linreg_info(demcz_ctx* c)
{
    demcz_linreg_state* s = c->linreg;
    if (!s) return "";
    char buf[512];
    if (s->kind == 1)
        snprintf(buf, sizeof buf, "linreg_kind=f16;linreg_terms=%d;linreg_link=%s;linreg_kernel=linreg_ssr_f16_kernel<%d,%d> (tcgen05 kind::f16, 2-way FP16 split: %d MMA terms, A in TMEM%s%s);",
                 s->terms, s->link ? "logit" : "gaussian", s->ppad, kTileNH, s->terms, s->swap_halves ? ", swapped halves" : "", s->link ? ", softplus epilogue" : "");
    else
        snprintf(buf, sizeof buf, "linreg_kind=tf32;linreg_terms=%d;linreg_kernel=linreg_ssr_kernel<%d,%d> (tcgen05 kind::tf32, 3xTF32 split, A in TMEM);linreg_fallback=%s;",
                 s->terms, s->ppad / 32, s->tile_n, s->fallback_reason.c_str());
    return buf;
}

double demcz_linreg_profile_ms(demcz_ctx* c, long long* launches)
{
    demcz_linreg_state* s = c->linreg;
    double tot = 0.0;
    if (launches) *launches = 0;
    if (!s) return 0.0;
    for (size_t i = 0; i + 1 < s->ev_used; i += 2) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, s->ev[i], s->ev[i + 1]) == cudaSuccess) tot += t;
    }
    if (launches) *launches = (long long)(s->ev_used / 2);
    s->ev_used = 0;
    return tot;
}

// ------------------------------------------------------------------------------------------------
// generations [g, g + n_gens) of this context (0-based g), archive frozen; append after the last one if asked
// ------------------------------------------------------------------------------------------------
int demcz_linreg_generations(demcz_ctx* c, int64_t g, int n_gens, bool append, bool acc, int half, bool snooker)
{
    demcz_linreg_state* s = c->linreg;
    if (!s) return demcz_fail(c, DEMCZ_ERR_INVALID, "linreg likelihood not set");
    const demcz_opts& o = c->o;
    const int CT = s->Npad / 128;
    uint32_t Mdraw = 0, zstart = 0;
    { int rc = demcz_archive_view(c, &Mdraw, &zstart); if (rc != DEMCZ_OK) return rc; }
    float* hist = nullptr; void* lphist = nullptr;
    { int rc = demcz_hist_begin(c, g, n_gens, &hist, &lphist); if (rc != DEMCZ_OK) return rc; }
    if (snooker && Mdraw < 3) return demcz_fail(c, DEMCZ_ERR_INVALID, "demcz_run: the snooker update needs 3 drawable archive rows");

    // Two half-populations (split at a multiple of the 128-chain MMA tile), each on its own stream: prepare / accept of
    // one half run under the GEMM of the other.  Chains are independent while the archive is frozen and every per-chain
    // sum keeps its order, so the result is bit-identical to one stream.  Not split: a single chain tile, per-launch
    // timing (profile), the per-generation moments kernel (it reads the whole state).
    static const bool split_env = !(getenv("DEMCZ_LINREG_SPLIT") && atoi(getenv("DEMCZ_LINREG_SPLIT")) == 0);
    const int nh = (split_env && CT >= 2 && !o.profile && !o.moments) ? 2 : 1;
    const int CT0 = nh == 2 ? (CT + 1) / 2 : CT;
    cudaStream_t st[2] = {c->stream, c->stream};
    if (nh == 2) {
        if (!s->hstream[0]) {
            for (int h = 0; h < 2; ++h) {
                CUL(c, cudaStreamCreateWithFlags(&s->hstream[h], cudaStreamNonBlocking));
                CUL(c, cudaEventCreateWithFlags(&s->ev_join[h], cudaEventDisableTiming));
            }
            CUL(c, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        }
        CUL(c, cudaEventRecord(s->ev_fork, c->stream));
        for (int h = 0; h < 2; ++h) { st[h] = s->hstream[h]; CUL(c, cudaStreamWaitEvent(st[h], s->ev_fork, 0)); }
    }
    for (int gi = 0; gi < n_gens; ++gi) {
        const int64_t grel = g + gi + 1;
        const int nb_run = snooker ? 1 : o.Nblocks;              // a snooker generation is ONE proposal moving all coordinates
        for (int b = 0; b < nb_run; ++b) {
            for (int h = 0; h < nh; ++h) {
                const int ct_lo = h == 0 ? 0 : CT0, ct_n = nh == 1 ? CT : (h == 0 ? CT0 : CT - CT0);
                const int chain_lo = ct_lo * 128, chain_hi = std::min(o.N, (ct_lo + ct_n) * 128);
                if (chain_lo >= chain_hi) continue;
                const int n_sub = chain_hi - chain_lo;
                LinArgs a;
                memset(&a, 0, sizeof a);
                a.X = c->dX; a.N = o.N; a.ldx = c->ldx; a.d = o.d; a.ppad = s->ppad;
                a.chain_lo = chain_lo; a.chain_hi = chain_hi;
                a.chain_offset = (uint32_t)o.chain_offset;
                a.Z = c->dZ; a.ZS = c->ZS;
                a.Mdraw = Mdraw; a.zcap = (uint32_t)c->zcap; a.zstart = zstart;
                a.gen = o.gen_offset + grel; a.grel = grel;
                a.keys = c->keys;
                a.block = b; a.nblocks = o.Nblocks;
                uint64_t mask = 0;
                for (int t = c->block_ptr[b]; t < c->block_ptr[b + 1]; ++t) mask |= (uint64_t)1 << c->block_idx[t];
                a.mask_lo = (uint32_t)mask; a.mask_hi = (uint32_t)(mask >> 32);
                const int blocklen = c->block_ptr[b + 1] - c->block_ptr[b];
                a.cscale = (float)(blocklen == 1 ? c->gamma : c->gamma / sqrt(2.0 * blocklen));   // src/demcz.jl:183-187
                for (int k = 0; k < o.d; ++k) a.eps[k] = (float)c->eps[k];
                a.Xprop = s->Xprop; a.wacc = s->wacc; a.Dhi = s->Dhi; a.Dlo = s->Dlo; a.lin = s->lin; a.b0 = s->d_b0; a.g = s->d_g;
                a.kind = s->kind; a.Dh = s->Dh; a.Dexp = s->Dexp; memcpy(a.xexp, s->xexp, sizeof a.xexp); a.swap_halves = s->swap_halves;
                a.tr_r1 = c->tr_r1; a.tr_r2 = c->tr_r2; a.tr_u = c->tr_u; a.tr_xprop = c->tr_xprop;
                a.flags = demcz_flags(c); a.world = o.world; a.my_rank = o.rank;
                a.need = (gi == 0 && b == 0) ? demcz_comm_need(c) : 0ull;          // once per archive view (and per half)
                a.gamma_s = (float)o.snooker_gamma; a.geo = s->geo;
                linreg_prepare_kernel<<<grid1(n_sub, 128), 128, 0, st[h]>>>(a, true, snooker);
                CUL(c, cudaGetLastError());

                if (o.profile) {
                    while (s->ev.size() < s->ev_used + 2) { cudaEvent_t ev; CUL(c, cudaEventCreate(&ev)); s->ev.push_back(ev); }
                    CUL(c, cudaEventRecord(s->ev[s->ev_used], st[h]));
                }
                // the GEMM of this half: operand rows, exponents and partial sums of chains [chain_lo, ...) through offset pointers
                const size_t po = (size_t)chain_lo * s->ppad;
                CUL(c, launch_ssr(c, s, s->Dhi ? s->Dhi + po : nullptr, s->Dlo ? s->Dlo + po : nullptr, ct_n, s->Npad, s->partial + chain_lo,
                                  s->Dh ? s->Dh + po : nullptr, s->Dexp ? s->Dexp + chain_lo : nullptr, st[h]));
                if (o.profile) { CUL(c, cudaEventRecord(s->ev[s->ev_used + 1], st[h])); s->ev_used += 2; }
                if (h == 0) s->ssr_launches++;

                AccArgs q;
                memset(&q, 0, sizeof q);
                q.X = c->dX; q.Xprop = s->Xprop; q.lp = s->lp64; q.wacc = s->wacc;
                q.N = o.N; q.ldx = c->ldx; q.d = o.d; q.Npad = s->Npad; q.R = s->R * (kEpiWarps / 4);
                q.chain_lo = chain_lo; q.chain_hi = chain_hi;
                q.partial = s->partial; q.lin = s->lin; q.ssr0 = s->ssr0; q.link = s->link;
                q.tempered = o.anneal ? 1 : 0;
                q.T = o.anneal ? c->hT[(size_t)grel - 1] : 1.f;
                q.grel = grel; q.block = b; q.nblocks = o.Nblocks; q.last_block = (b == nb_run - 1);
                q.geo = snooker ? s->geo : nullptr;
                q.lp_gen0 = o.Nblocks > 1 ? s->lp64 + c->ldx : nullptr;
                q.chain_changed = c->d_chain_changed; q.win_changed = c->d_counters + 0; q.accepted_total = c->d_counters + 1;
                q.count_run = (grel != 1) ? 1 : 0;
                q.count_win = (o.anneal && o.adapt && (grel - 1) % o.adapt_every == 0) ? 0 : 1;
                const bool keep = hist != nullptr && grel % c->thin == 0;           // history thinning (opts.thin)
                q.hist = keep ? hist : nullptr; q.lphist = (double*)lphist; q.hidx = grel / c->thin - 1;
                const bool do_append = append && gi == n_gens - 1;
                q.Zw = do_append ? c->dZ : nullptr; q.ZS = c->ZS;
                q.append_row0 = (uint32_t)(c->M % c->zcap); q.zcap = (uint32_t)c->zcap; q.chain_offset = (uint32_t)o.chain_offset;
                if (acc) { q.S1 = c->dS1[half]; q.S2 = c->dS2 + (size_t)half * o.d; q.shift = c->dshift; }
                q.tr_lpprop = c->tr_lpprop; q.tr_acc = c->tr_acc;
                linreg_accept_kernel<<<grid1(n_sub, 128), 128, 0, st[h]>>>(q);
                CUL(c, cudaGetLastError());
            }
        }
        if (o.moments) { int rc = demcz_moments_add_state(c); if (rc != DEMCZ_OK) return rc; }
    }
    if (nh == 2)
        for (int h = 0; h < 2; ++h) { CUL(c, cudaEventRecord(s->ev_join[h], st[h])); CUL(c, cudaStreamWaitEvent(c->stream, s->ev_join[h], 0)); }
    return demcz_hist_end(c, g, n_gens);
}
