// demcz_lib.cu — libdemcz_cuda.so: context, host-side generation loop and the C ABI of
// include/demcz.h.  Reference behaviour mirrored (file:line under the upstream repo):
//   demcz_sample   src/demcz.jl:9-63        runchain!      src/demcz.jl:80-93
//   demcz_anneal   src/demcz_anneal.jl:18-65 (adaptγ :48-57, tempbaseline :1-3)
//   Rhat_gelman    src/utils.jl:2-20
// No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/demcz.h"
#include "demcz_step_host.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace demcz;

static std::string g_create_error;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
int demcz_fail(demcz_ctx* c, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) {
        c->err = buf;
        if (code == DEMCZ_ERR_CUDA || code == DEMCZ_ERR_NCCL) c->sticky = code;
    } else {
        g_create_error = buf;
    }
    return code;
}

#define CU(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return demcz_fail((c), e_ == cudaErrorMemoryAllocation ? DEMCZ_ERR_OOM : DEMCZ_ERR_CUDA,  \
                              "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
    } while (0)

#define CHECK_CTX(c)                                                                    \
    do {                                                                                \
        if (!(c)) return DEMCZ_ERR_INVALID;                                             \
        if ((c)->sticky) return (c)->sticky;                                            \
        if (cudaSetDevice((c)->device) != cudaSuccess)                                  \
            return demcz_fail((c), DEMCZ_ERR_CUDA, "cudaSetDevice(%d) failed", (c)->device); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// small layout / reduction kernels
// ------------------------------------------------------------------------------------------------
namespace {

// Julia column-major Float64 (rows x d, leading dim ldin) -> archive rows (row-major, stride LD, FP32)
__global__ void import_rows_kernel(const double* __restrict__ in, long long rows, long long ldin, int d, int LD,
                                   float* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    for (int k = 0; k < LD; ++k) out[i * LD + k] = k < d ? (float)in[i + ldin * k] : 0.f;   // LD = row slot (floats)
}

__global__ void export_rows_kernel(const float* __restrict__ Z, long long rows, long long zstart, long long zcap,
                                   int d, int LD, double* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    long long p = zstart + i;
    if (p >= zcap) p -= zcap;
    for (int k = 0; k < d; ++k) out[i + rows * k] = (double)Z[p * LD + k];
}

// column-major Float64 (n x d) <-> SoA FP32 [k][ldx]
__global__ void import_soa_kernel(const double* __restrict__ in, long long n, int d, long long ldx, float* __restrict__ X)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < d; ++k) X[k * ldx + i] = (float)in[i + n * k];
}

__global__ void export_soa_kernel(const float* __restrict__ X, long long n, int d, long long ldx, double* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < d; ++k) out[i + n * k] = (double)X[k * ldx + i];
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, long long n, double* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, long long n, float* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// X = Zmat[M-N+1:M, :]  (src/demcz.jl:15 as intended, :113): archive rows -> SoA state
__global__ void rows_to_soa_kernel(const float* __restrict__ Z, long long row0, long long zcap, int n, int d, int LD,
                                   int ldx, float* __restrict__ X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long p = row0 + i;
    if (p >= zcap) p -= zcap;
    for (int k = 0; k < d; ++k) X[(size_t)k * ldx + i] = Z[p * LD + k];
}

__device__ __forceinline__ double block_sum(double v, double* sm)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;   // valid on thread 0
}

// sums[k] += sum_i X[k][i]   (grid: (blocks, d))
__global__ void colsum_kernel(const float* __restrict__ X, int n, int ldx, double* __restrict__ sums)
{
    __shared__ double sm[32];
    const int k = blockIdx.y;
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += X[(size_t)k * ldx + i];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) atomicAdd(&sums[k], s);
}

__global__ void shift_from_sums_kernel(const double* __restrict__ sums, int d, double inv_n, float* __restrict__ shift,
                                       double* __restrict__ mean_out)
{
    const int k = threadIdx.x;
    if (k < d) {
        if (shift) shift[k] = (float)(sums[k] * inv_n);
        if (mean_out) mean_out[k] = sums[k] * inv_n;
    }
}

// cov[k1][k2] += sum_i (x1 - m1)(x2 - m2)    (grid: (blocks, d*d))
__global__ void cov_kernel(const float* __restrict__ X, int n, int ldx, int d, const double* __restrict__ mean,
                           double* __restrict__ cov)
{
    __shared__ double sm[32];
    const int k1 = blockIdx.y / d, k2 = blockIdx.y % d;
    const double m1 = mean[k1], m2 = mean[k2];
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        s += ((double)X[(size_t)k1 * ldx + i] - m1) * ((double)X[(size_t)k2 * ldx + i] - m2);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) atomicAdd(&cov[k1 * d + k2], s);
}

// streaming moments of one generation from the SoA state: S1[k] += sum_i (x_k - s_k), S2[k1][k2] += sum_i (x_k1 - s_k1)(x_k2 - s_k2)
// grid: (blocks, d * d + d); blockIdx.y < d*d: pair (k1, k2) with k1 <= k2 (others idle), else column sums
__global__ void state_moments_kernel(const float* __restrict__ X, int n, int ldx, int d, const float* __restrict__ shift,
                                     double* __restrict__ mom)
{
    __shared__ double sm[32];
    const int y = blockIdx.y;
    double s = 0.0;
    if (y < d * d) {
        const int k1 = y / d, k2 = y % d;
        if (k1 > k2) return;
        const double s1 = shift[k1], s2 = shift[k2];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            s += ((double)X[(size_t)k1 * ldx + i] - s1) * ((double)X[(size_t)k2 * ldx + i] - s2);
        s = block_sum(s, sm);
        if (threadIdx.x == 0) atomicAdd(&mom[64 + k1 * 64 + k2], s);
    } else {
        const int k = y - d * d;
        const double s1 = shift[k];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (double)X[(size_t)k * ldx + i] - s1;
        s = block_sum(s, sm);
        if (threadIdx.x == 0) atomicAdd(&mom[k], s);
    }
}

// archive rows (row-major slots, ring) -> per-column sums / centred cross products, FP64.  grid: (blocks, d) / (blocks, d*d)
__global__ void rows_colsum_kernel(const float* __restrict__ Z, long long row0, long long n, long long zstart, long long zcap,
                                   int ZS, double* __restrict__ sums)
{
    __shared__ double sm[32];
    const int k = blockIdx.y;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long p = zstart + row0 + i;
        if (p >= zcap) p -= zcap;
        s += Z[p * ZS + k];
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) atomicAdd(&sums[k], s);
}

__global__ void rows_cov_kernel(const float* __restrict__ Z, long long row0, long long n, long long zstart, long long zcap,
                                int ZS, int d, const double* __restrict__ mean, double* __restrict__ cov)
{
    __shared__ double sm[32];
    const int k1 = blockIdx.y / d, k2 = blockIdx.y % d;
    const double m1 = mean[k1], m2 = mean[k2];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long p = zstart + row0 + i;
        if (p >= zcap) p -= zcap;
        s += ((double)Z[p * ZS + k1] - m1) * ((double)Z[p * ZS + k2] - m2);
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) atomicAdd(&cov[k1 * d + k2], s);
}

// 64-bit checksum of archive rows: sum over (logical row i, column k) of mix64(bits(Z[i][k]) ^ (i * 64 + k) * golden).
// Order-independent (a sum), position-sensitive (the position is mixed in): equal iff the replicas hold the same rows.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void rows_checksum_kernel(const float* __restrict__ Z, long long row0, long long n, long long zstart, long long zcap,
                                     int ZS, int d, unsigned long long* __restrict__ out)
{
    unsigned long long h = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long p = zstart + row0 + i;
        if (p >= zcap) p -= zcap;
        for (int k = 0; k < d; ++k)
            h += mix64((unsigned long long)__float_as_uint(Z[p * ZS + k]) ^ ((unsigned long long)((row0 + i) * 64 + k) * 0x9E3779B97F4A7C15ull));
    }
    h = warp_sum(h);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);
}

// Gelman-Rubin partial sums over the local chain-halves (Rhat_gelman, src/utils.jl:2-20):
//   part[0*d+k] += sum_j mean_j    part[1*d+k] += sum_j mean_j^2    part[2*d+k] += sum_j S1_j^2
// with mean_j = shift + S1_j/n.  Warp-shuffle -> block -> one FP64 atomic per block.
// grid: (blocks, d, 2 halves)
__global__ void rhat_partial_kernel(const float* __restrict__ S1h0, const float* __restrict__ S1h1, int n_chains,
                                    int ldx, int d, const float* __restrict__ shift, double inv_n,
                                    double* __restrict__ part)
{
    __shared__ double sm[32];
    const int k = blockIdx.y;
    const float* S1 = blockIdx.z ? S1h1 : S1h0;
    const double sh = shift[k];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chains; i += gridDim.x * blockDim.x) {
        const double s = S1[(size_t)k * ldx + i];
        const double m = sh + s * inv_n;
        a += m; b += m * m; c += s * s;
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) {
        atomicAdd(&part[0 * d + k], a);
        atomicAdd(&part[1 * d + k], b);
        atomicAdd(&part[2 * d + k], c);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// kernel dispatch (the step-kernel instantiations live in demcz_step_launch.cu, compiled once per width group)
// ------------------------------------------------------------------------------------------------
namespace {

template <int LD>
cudaError_t launch_eval_t(demcz_ctx* c, const float* X, int n, int ldx, float* out)
{
    const MvnParams<LD> P = demcz_make_mvn_params<LD>(c, false);
    mvn_eval_kernel<LD><<<(n + 127) / 128, 128, 0, c->stream>>>(X, n, ldx, c->o.d, out, P);
    return cudaGetLastError();
}

cudaError_t launch_eval(demcz_ctx* c, const float* X, int n, int ldx, float* out)
{
    switch (c->LD) {
        case 4:  return launch_eval_t<4>(c, X, n, ldx, out);
        case 8:  return launch_eval_t<8>(c, X, n, ldx, out);
        case 12: return launch_eval_t<12>(c, X, n, ldx, out);
        case 16: return launch_eval_t<16>(c, X, n, ldx, out);
        case 20: return launch_eval_t<20>(c, X, n, ldx, out);
        case 24: return launch_eval_t<24>(c, X, n, ldx, out);
        case 28: return launch_eval_t<28>(c, X, n, ldx, out);
        case 32: return launch_eval_t<32>(c, X, n, ldx, out);
    }
    return cudaErrorInvalidValue;
}

inline int grid1d(long long n, int threads = 256) { return (int)((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------
// Snooker generation (ter Braak & Vrugt 2008, the update the reference's README cites but does not implement:
// README.md:57-59): one proposal per chain along the line through the chain and an archive point z,
//     x' = x + gamma_s * e * ((z1 - z2) . e) / (e . e),   e = x - z,   z, z1, z2 three distinct archive rows,
// accepted with probability min(1, pi(x')/pi(x) * (|x' - z| / |x - z|)^(d-1)); |x' - z| / |x - z| = |1 + gamma_s proj|.
// Ignores the block structure (moves all coordinates).  A generation is a snooker generation as a whole (chosen from
// the Philox stream by the host), so the fused step kernel is untouched; this kernel is thread-per-chain.
// ------------------------------------------------------------------------------------------------
struct SnookerArgs {
    float* X; float* logp; int N, ldx, d;
    uint32_t chain_offset;
    const float* Z; int ZS; uint32_t Mdraw, zstart, zcap;
    long long gen, grel;
    PhiloxKeys keys;
    float gamma_s;                 // 0: uniform on [1.2, 2.2]
    const float* Ttab;             // temperature table or nullptr
    int nblocks;                   // trace layout only
    float* hist; float* lphist; long long hidx;      // history slot of this generation or hist == nullptr
    uint32_t *tr_r1, *tr_r2; float *tr_u, *tr_xprop, *tr_lpprop; uint8_t* tr_acc;
    int* chain_changed; unsigned long long *win_changed, *accepted_total; int count_run, count_win;
    float* S1; double* S2; float shift[32];
    float* Zw; uint32_t append_row0;
    const unsigned long long* flags; int world, my_rank; unsigned long long need;
};

template <int LD>
__global__ void snooker_mvn_kernel(const __grid_constant__ SnookerArgs a, const __grid_constant__ MvnParams<LD> P)
{
    __shared__ unsigned int red[2];
    __shared__ float red_s2[LD];
    wait_peer_segments(a.flags, a.world, a.my_rank, a.need);
    if (threadIdx.x < 2) red[threadIdx.x] = 0u;
    if (threadIdx.x < LD) red_s2[threadIdx.x] = 0.f;
    __syncthreads();
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = chain < a.N;
    const int cl = active ? chain : a.N - 1;
    const uint32_t cid = a.chain_offset + (uint32_t)cl;
    const uint32_t g_lo = (uint32_t)a.gen, g_hi = (uint32_t)((unsigned long long)a.gen >> 32);
    float x[LD];
#pragma unroll
    for (int k = 0; k < LD; ++k) x[k] = k < a.d ? a.X[(size_t)k * a.ldx + cl] : 0.f;
    const float lp = a.logp[cl];

    const u32x4 w = philox4x32_10(cid, g_lo, kSnookerCall, g_hi, a.keys);
    uint32_t z, z1, z2;
    draw_triple(__umulhi(w.x, a.Mdraw), __umulhi(w.y, a.Mdraw - 1u), __umulhi(w.z, a.Mdraw - 2u), z, z1, z2);
    float gs = a.gamma_s;
    if (gs == 0.f) gs = 1.2f + uniform01(philox4x32_10(cid, g_lo, kSnookerCall + 1u, g_hi, a.keys).x);
    auto phys = [&](uint32_t r) { r += a.zstart; if (r >= a.zcap) r -= a.zcap; return a.Z + (size_t)r * a.ZS; };
    const float *pz = phys(z), *p1 = phys(z1), *p2 = phys(z2);
    float ee = 0.f, num = 0.f, e[LD];
#pragma unroll
    for (int k = 0; k < LD; ++k) {
        const float zk = k < a.d ? pz[k] : 0.f, dk = k < a.d ? p1[k] - p2[k] : 0.f;
        e[k] = x[k] - zk;
        ee = fmaf(e[k], e[k], ee);
        num = fmaf(dk, e[k], num);
    }
    const float proj = ee > 0.f ? gs * (num / ee) : 0.f;              // x == z: no direction, the proposal is x itself (rejected below)
    float2 xp2[LD / 2];
#pragma unroll
    for (int k = 0; k < LD; k += 2) xp2[k / 2] = make_float2(fmaf(proj, e[k], x[k]), fmaf(proj, e[k + 1], x[k + 1]));
    const float lpp = mvn_logdensity<LD, false>(P.W, P.mu2, P.c0, xp2);
    const float geo = (float)(a.d - 1) * __logf(fabsf(1.f + proj));   // (d-1) log(|x'-z| / |x-z|); -inf when x' = z
    const float u = uniform01(w.w);
    const float logu = __logf(u);
    const float T = a.Ttab ? a.Ttab[a.grel - 1] : 1.f;
    const bool acc = ee > 0.f && (logu * T < (lpp - lp) + T * geo);   // NaN -> reject; T = 0 -> greedy on the objective

    if (a.tr_r1 && active) {
        const size_t o = ((size_t)(a.grel - 1) * a.nblocks) * a.N + chain;      // block slot 0 of this generation
        a.tr_r1[o] = z1; a.tr_r2[o] = z2; a.tr_u[o] = u; a.tr_lpprop[o] = lpp; a.tr_acc[o] = acc ? 1 : 0;
#pragma unroll
        for (int k = 0; k < LD; ++k)
            if (k < a.d) a.tr_xprop[(((size_t)(a.grel - 1) * a.nblocks) * a.d + k) * a.N + chain] = (k & 1) ? xp2[k / 2].y : xp2[k / 2].x;
    }
    float lpn = lp;
    if (acc) {
#pragma unroll
        for (int k = 0; k < LD; ++k) x[k] = (k & 1) ? xp2[k / 2].y : xp2[k / 2].x;
        lpn = lpp;
    }
    if (active) {
        if (acc) {
#pragma unroll
            for (int k = 0; k < LD; ++k) if (k < a.d) a.X[(size_t)k * a.ldx + chain] = x[k];
            a.logp[chain] = lpn;
        }
        if (acc && lpn != lp && a.count_run) a.chain_changed[chain] += 1;      // generations whose log-objective changed (src/utils.jl:61)
        if (a.hist) {
#pragma unroll
            for (int k = 0; k < LD; ++k) if (k < a.d) a.hist[((size_t)a.hidx * a.d + k) * a.N + chain] = x[k];
            a.lphist[(size_t)a.hidx * a.N + chain] = lpn;
        }
        if (a.Zw) {
            uint32_t row = a.append_row0 + a.chain_offset + (uint32_t)chain;
            if (row >= a.zcap) row -= a.zcap;
            float* dst = a.Zw + (size_t)row * a.ZS;
#pragma unroll
            for (int k = 0; k < LD; ++k) dst[k] = x[k];             // LD <= ZS: the 16-byte padded row
        }
    }
    const bool changed = active && acc && lpn != lp;
    if (a.S1) {
#pragma unroll
        for (int k = 0; k < LD; ++k) {
            if (k < a.d) {                                         // warp-uniform
                const float t = active ? x[k] - a.shift[k] : 0.f;
                if (active) a.S1[(size_t)k * a.ldx + chain] += t;
                const float v = warp_sum(t * t);
                if ((threadIdx.x & 31) == 0) atomicAdd(&red_s2[k], v);
            }
        }
    }
    const int wa = warp_sum((acc && active) ? 1 : 0), ww = warp_sum((changed && a.count_win) ? 1 : 0);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[0], (unsigned)wa); atomicAdd(&red[1], (unsigned)ww); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (red[0]) atomicAdd(a.accepted_total, (unsigned long long)red[0]);
        if (red[1]) atomicAdd(a.win_changed, (unsigned long long)red[1]);
    }
    if (a.S1 && threadIdx.x < a.d) atomicAdd(&a.S2[threadIdx.x], (double)red_s2[threadIdx.x]);
}

template <int LD>
cudaError_t launch_snooker_t(demcz_ctx* c, const SnookerArgs& a)
{
    const MvnParams<LD> P = demcz_make_mvn_params<LD>(c, false);
    snooker_mvn_kernel<LD><<<(a.N + 127) / 128, 128, 0, c->stream>>>(a, P);
    return cudaGetLastError();
}

cudaError_t launch_snooker(demcz_ctx* c, const SnookerArgs& a)
{
    switch (c->LD) {
        case 4:  return launch_snooker_t<4>(c, a);
        case 8:  return launch_snooker_t<8>(c, a);
        case 12: return launch_snooker_t<12>(c, a);
        case 16: return launch_snooker_t<16>(c, a);
        case 20: return launch_snooker_t<20>(c, a);
        case 24: return launch_snooker_t<24>(c, a);
        case 28: return launch_snooker_t<28>(c, a);
        case 32: return launch_snooker_t<32>(c, a);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

// Is generation `gen_abs` (absolute, 1-based) a snooker generation?  Same answer on every rank and in the oracle.
bool demcz_is_snooker_gen(const demcz_ctx* c, long long gen_abs)
{
    if (!(c->o.snooker_prob > 0.0)) return false;
    const u32x4 w = philox4x32_10_host(0xFFFFFFFFu, (uint32_t)gen_abs, kSnookerCall, (uint32_t)((unsigned long long)gen_abs >> 32), c->keys);
    const double u = ((double)(w.x >> 9) + 0.5) * (1.0 / 8388608.0);
    return u < c->o.snooker_prob;
}

// ------------------------------------------------------------------------------------------------
// ABI: lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" int demcz_version(void) { return DEMCZ_VERSION; }

extern "C" const char* demcz_last_error(const demcz_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int demcz_create(demcz_ctx** out, const demcz_opts* o, int device)
{
    if (!out || !o) return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: null argument");
    *out = nullptr;
    if (o->struct_size != (int32_t)sizeof(demcz_opts))
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: struct_size %d != %zu (ABI mismatch)", o->struct_size,
                          sizeof(demcz_opts));
    if (o->N < 1 || o->K < 1 || o->Ngeneration < 0 || o->d < 1 || o->Nblocks < 1 || !o->block_ptr || !o->block_idx ||
        !o->eps_scale)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: bad N/K/Ngeneration/d/Nblocks or null arrays");
    if (o->d > DEMCZ_MAX_DIM_LINREG)
        return demcz_fail(nullptr, DEMCZ_ERR_UNSUPPORTED, "demcz_create: d=%d > %d", o->d, DEMCZ_MAX_DIM_LINREG);
    if (o->Nblocks > kMaxBlocks)
        return demcz_fail(nullptr, DEMCZ_ERR_UNSUPPORTED, "demcz_create: Nblocks=%d > %d", o->Nblocks, kMaxBlocks);
    if (o->world < 1 || o->rank < 0 || o->rank >= o->world || o->N_global < o->N || o->chain_offset < 0 ||
        o->chain_offset + o->N > o->N_global)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: inconsistent rank/world/N_global/chain_offset");
    if (o->autostop && (o->autostop_every < 4))
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: autostop_every must be >= 4");
    if (o->exchange_layout != DEMCZ_XCHG_AUTO && o->exchange_layout != DEMCZ_XCHG_IN_PLACE && o->exchange_layout != DEMCZ_XCHG_DENSE_SLOTS)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: exchange_layout %d (0 auto, 1 in_place, 3 dense_slots)", o->exchange_layout);
    if (o->push_engine < DEMCZ_PUSH_AUTO || o->push_engine > DEMCZ_PUSH_COPY_ENGINE)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: push_engine %d", o->push_engine);
    if (o->thin < 0 || o->moments < 0 || o->moments > 1 || o->persistent < -1 || o->persistent > 1)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: bad thin/moments/persistent");
    if (!(o->snooker_prob >= 0.0 && o->snooker_prob <= 1.0) || !(o->snooker_gamma >= 0.0))
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: snooker_prob must be in [0,1], snooker_gamma >= 0");
    if (o->anneal && o->adapt && o->adapt_every < 1)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: adapt_every must be >= 1");
    for (int b = 0; b < o->Nblocks; ++b) {
        if (o->block_ptr[b + 1] <= o->block_ptr[b])
            return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: empty block %d", b);
        for (int t = o->block_ptr[b]; t < o->block_ptr[b + 1]; ++t)
            if (o->block_idx[t] < 0 || o->block_idx[t] >= o->d)
                return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: block %d has column %d outside 0..%d", b,
                                  o->block_idx[t], o->d - 1);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return demcz_fail(nullptr, DEMCZ_ERR_CUDA, "demcz_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev)
        return demcz_fail(nullptr, DEMCZ_ERR_INVALID, "demcz_create: device %d of %d", device, ndev);

    demcz_ctx* c = new demcz_ctx();
    c->o = *o;
    c->device = device;
    c->block_ptr.assign(o->block_ptr, o->block_ptr + o->Nblocks + 1);
    c->block_idx.assign(o->block_idx, o->block_idx + o->block_ptr[o->Nblocks]);
    c->eps.assign(o->eps_scale, o->eps_scale + o->d);
    c->o.block_ptr = c->block_ptr.data();
    c->o.block_idx = c->block_idx.data();
    c->o.eps_scale = c->eps.data();
    c->gamma = o->gamma;
    c->LD = (o->d + 3) / 4 * 4;
    c->ZS = o->d <= DEMCZ_MAX_DIM ? zslot(c->LD) : c->LD;
    // rows back to back instead of padded slots when the exchange rather than the gather bounds a K-block (demcz.h)
    if (o->exchange_layout == DEMCZ_XCHG_DENSE_SLOTS || (o->exchange_layout == DEMCZ_XCHG_AUTO && o->world >= 4)) c->ZS = c->LD;
    c->ldx = (o->N + 31) / 32 * 32;
    c->keys = make_philox_keys(o->seed);
    if (const char* gr = getenv("DEMCZ_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(gr));
    c->rhat.assign(o->d, NAN);

    auto bail = [&](int code) { demcz_destroy(c); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(demcz_fail(nullptr, DEMCZ_ERR_CUDA, "cudaSetDevice(%d)", device));
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(demcz_fail(nullptr, DEMCZ_ERR_CUDA, "cudaStreamCreate"));
    const size_t nx = (size_t)c->ldx * o->d;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) { e = cudaMalloc(p, bytes ? bytes : 16); if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes ? bytes : 16, c->stream); }
    };
    alloc((void**)&c->dX, nx * sizeof(float));
    alloc((void**)&c->dlogp, (size_t)c->ldx * sizeof(float));
    alloc((void**)&c->d_chain_changed, (size_t)c->ldx * sizeof(int));
    alloc((void**)&c->d_counters, 4 * sizeof(unsigned long long));
    alloc((void**)&c->d_scratch, (size_t)(8 * o->d + o->d * o->d + 16) * sizeof(double));
    alloc((void**)&c->d_stage, ((size_t)o->N * o->d + o->N) * sizeof(double));
    if (o->autostop) {
        alloc((void**)&c->dS1[0], nx * sizeof(float));
        alloc((void**)&c->dS1[1], nx * sizeof(float));
        alloc((void**)&c->dS2, (size_t)2 * o->d * sizeof(double));
        alloc((void**)&c->dshift, (size_t)o->d * sizeof(float));
    }
    c->thin = o->thin > 1 ? o->thin : 1;
    // small-N persistent launches: every chain in one CTA of the step kernel, one GPU, padded slots, no streaming moments
    // (the streamed history's slice buffers hold one K-block: store_chain = 2 keeps one launch per K-block)
    c->persistent_ok = o->persistent != 0 && o->world == 1 && o->N <= 128 && !o->moments && o->store_chain != 2 && o->d <= DEMCZ_MAX_DIM &&
                       c->ZS == zslot(c->LD);
    if (o->store_chain == 1 && o->Ngeneration / c->thin > 0) {
        alloc((void**)&c->dhist, (size_t)(o->Ngeneration / c->thin) * o->d * o->N * sizeof(float));
        alloc((void**)&c->dlphist, (size_t)(o->Ngeneration / c->thin) * o->N * sizeof(float));
    }
    if (o->trace && o->Ngeneration > 0) {
        const size_t np = (size_t)o->Ngeneration * o->Nblocks * o->N;
        alloc((void**)&c->tr_r1, np * 4); alloc((void**)&c->tr_r2, np * 4); alloc((void**)&c->tr_u, np * 4);
        alloc((void**)&c->tr_lpprop, np * 4); alloc((void**)&c->tr_acc, np);
        alloc((void**)&c->tr_xprop, np * o->d * 4);
    }
    if (e != cudaSuccess)
        return bail(demcz_fail(nullptr, e == cudaErrorMemoryAllocation ? DEMCZ_ERR_OOM : DEMCZ_ERR_CUDA,
                               "demcz_create: device allocation failed: %s", cudaGetErrorString(e)));
    if (o->anneal) {   // default schedule: tempbaseline(ig, Ngeneration, T0, TN)  (src/demcz_anneal.jl:1-3; Q3)
        std::vector<double> T((size_t)o->Ngeneration);
        for (int ig = 1; ig <= o->Ngeneration; ++ig) T[ig - 1] = o->T0 * pow(o->TN / o->T0, (double)ig / o->Ngeneration);
        *out = c;
        int rc = demcz_set_temperature_table(c, T.data(), (int64_t)T.size());
        if (rc != DEMCZ_OK) { *out = nullptr; g_create_error = c->err; return bail(rc); }
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return bail(demcz_fail(nullptr, DEMCZ_ERR_CUDA, "sync after create"));
    *out = c;
    return DEMCZ_OK;
}

extern "C" void demcz_destroy(demcz_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    demcz_comm_destroy_internal(c);
    demcz_linreg_free(c);
    void* ptrs[] = {c->dX, c->dlogp, c->dZ, c->dT, c->dhist, c->dlphist, c->tr_r1, c->tr_r2, c->tr_u, c->tr_xprop,
                    c->tr_lpprop, c->tr_acc, c->d_chain_changed, c->d_counters, c->dS1[0], c->dS1[1], c->dS2, c->dshift,
                    c->d_scratch, c->d_stage, c->d_mom, c->dmomshift};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t ev : c->ev_pool) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->ev_xchg) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (c->hs_dx[i]) cudaFree(c->hs_dx[i]);
        if (c->hs_dlp[i]) cudaFree(c->hs_dlp[i]);
        if (c->hs_ev_k[i]) cudaEventDestroy(c->hs_ev_k[i]);
        if (c->hs_ev_c[i]) cudaEventDestroy(c->hs_ev_c[i]);
    }
    if (c->hs_hx) cudaFreeHost(c->hs_hx);
    if (c->hs_hlp) cudaFreeHost(c->hs_hlp);
    if (c->hs_stream) cudaStreamDestroy(c->hs_stream);
    if (c->ev_run[0]) { cudaEventDestroy(c->ev_run[0]); cudaEventDestroy(c->ev_run[1]); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ------------------------------------------------------------------------------------------------
// ABI: likelihood / archive / state
// ------------------------------------------------------------------------------------------------
extern "C" int demcz_set_likelihood_mvnormal(demcz_ctx* c, int32_t d, const double* mu, const double* Sigma,
                                             int32_t drop_constant)
{
    CHECK_CTX(c);
    if (!mu || !Sigma || d != c->o.d) return demcz_fail(c, DEMCZ_ERR_INVALID, "set_likelihood_mvnormal: d=%d, opts.d=%d", d, c->o.d);
    if (d > DEMCZ_MAX_DIM) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "set_likelihood_mvnormal: d=%d > %d", d, DEMCZ_MAX_DIM);
    // FP64 Cholesky Sigma = L L^T, then W = L^-1 (lower), rounded to FP32 for the kernels
    std::vector<double> L((size_t)d * d, 0.0), W((size_t)d * d, 0.0);
    double logdet = 0.0;
    for (int j = 0; j < d; ++j) {
        double s = Sigma[j + (size_t)d * j];
        for (int k = 0; k < j; ++k) s -= L[(size_t)j * d + k] * L[(size_t)j * d + k];
        if (!(s > 0.0)) return demcz_fail(c, DEMCZ_ERR_NOT_PD, "set_likelihood_mvnormal: Sigma is not positive definite (pivot %d)", j);
        const double ljj = sqrt(s);
        L[(size_t)j * d + j] = ljj;
        logdet += 2.0 * log(ljj);
        for (int i = j + 1; i < d; ++i) {
            double t = Sigma[i + (size_t)d * j];
            for (int k = 0; k < j; ++k) t -= L[(size_t)i * d + k] * L[(size_t)j * d + k];
            L[(size_t)i * d + j] = t / ljj;
        }
    }
    for (int j = 0; j < d; ++j) {                 // columns of L^-1 by forward substitution
        for (int i = j; i < d; ++i) {
            double t = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) t -= L[(size_t)i * d + k] * W[(size_t)k * d + j];
            W[(size_t)i * d + j] = t / L[(size_t)i * d + i];
        }
    }
    c->mvn_mu.assign(d, 0.f);
    c->mvn_W.assign((size_t)d * d, 0.f);
    for (int k = 0; k < d; ++k) c->mvn_mu[k] = (float)mu[k];
    for (size_t t = 0; t < (size_t)d * d; ++t) c->mvn_W[t] = (float)W[t];
    c->mvn_c0 = drop_constant ? 0.f : (float)(-0.5 * (d * log(2.0 * M_PI) + logdet));
    c->lik_kind = DEMCZ_LIK_MVN;
    c->state_lp_valid = false;
    return DEMCZ_OK;
}

extern "C" int demcz_set_archive(demcz_ctx* c, const double* Z, int64_t M0)
{
    CHECK_CTX(c);
    if (!Z || M0 < 2) return demcz_fail(c, DEMCZ_ERR_INVALID, "set_archive: need at least 2 rows (got %lld)", (long long)M0);
    const demcz_opts& o = c->o;
    int64_t cap = o.archive_capacity > 0 ? o.archive_capacity
                                         : M0 + (o.N_global * (int64_t)o.Ngeneration + o.K - 1) / o.K;   // src/demcz.jl:11
    if (cap < M0) cap = M0;
    // 2^31: ring positions are formed as 32-bit sums (logical row + start of the ring) inside the kernels
    if (cap >= (int64_t)1 << 31) return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "set_archive: capacity %lld >= 2^31 rows", (long long)cap);
    if (c->peers_open) return demcz_fail(c, DEMCZ_ERR_INVALID, "set_archive: the peers hold mappings of the current archive; create a new context");
    if (c->dZ) { cudaFree(c->dZ); c->dZ = nullptr; }
    c->zcap = cap;
    // archive + segment flags written by the peers (demcz_internal.h)
    CU(c, cudaMalloc((void**)&c->dZ, demcz_alloc_floats(c) * sizeof(float)));
    CU(c, cudaMemsetAsync(c->dZ + demcz_flag_off_floats(c), 0, (demcz_alloc_floats(c) - demcz_flag_off_floats(c)) * sizeof(float), c->stream));
    c->M = M0;
    c->appends = 0;
    c->vis_waited = 0;
    double* stage = nullptr;
    CU(c, cudaMalloc((void**)&stage, (size_t)M0 * o.d * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(stage, Z, (size_t)M0 * o.d * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        import_rows_kernel<<<grid1d(M0), 256, 0, c->stream>>>(stage, M0, M0, o.d, c->ZS, c->dZ);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(stage);
    CU(c, e);
    c->archive_set = true;
    return DEMCZ_OK;
}

static int eval_state_logp(demcz_ctx* c)
{
    if (c->lik_kind == DEMCZ_LIK_MVN) {
        CU(c, launch_eval(c, c->dX, c->o.N, c->ldx, c->dlogp));
    } else if (c->lik_kind == DEMCZ_LIK_LINREG) {
        int rc = demcz_linreg_eval(c, c->dX, c->o.N, c->ldx, nullptr);
        if (rc != DEMCZ_OK) return rc;
    } else {
        return demcz_fail(c, DEMCZ_ERR_INVALID, "no likelihood set");
    }
    c->state_lp_valid = true;
    return DEMCZ_OK;
}

extern "C" int demcz_set_state(demcz_ctx* c, const double* X, const double* logp)
{
    CHECK_CTX(c);
    if (!X) return demcz_fail(c, DEMCZ_ERR_INVALID, "set_state: X is null");
    const int N = c->o.N, d = c->o.d;
    double* stage = c->d_stage;
    CU(c, cudaMemcpyAsync(stage, X, (size_t)N * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    import_soa_kernel<<<grid1d(N), 256, 0, c->stream>>>(stage, N, d, c->ldx, c->dX);
    CU(c, cudaGetLastError());
    if (logp) {
        CU(c, cudaMemcpyAsync(stage + (size_t)N * d, logp, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (c->lik_kind == DEMCZ_LIK_LINREG) {
            CU(c, cudaMemcpyAsync(demcz_linreg_lp(c), stage + (size_t)N * d, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        } else {
            f64_to_f32_kernel<<<grid1d(N), 256, 0, c->stream>>>(stage + (size_t)N * d, N, c->dlogp);
            CU(c, cudaGetLastError());
        }
    }
    CU(c, cudaStreamSynchronize(c->stream));      // the caller's buffers are borrowed for the call only
    c->state_set = true;
    c->state_lp_valid = logp != nullptr;
    return DEMCZ_OK;
}

extern "C" int demcz_set_temperature_table(demcz_ctx* c, const double* T, int64_t len)
{
    CHECK_CTX(c);
    if (!T || len < c->o.Ngeneration) return demcz_fail(c, DEMCZ_ERR_INVALID, "set_temperature_table: need %d entries", c->o.Ngeneration);
    std::vector<float> Tf((size_t)len);
    for (int64_t i = 0; i < len; ++i) Tf[i] = (float)T[i];
    if (c->dT) { cudaFree(c->dT); c->dT = nullptr; }
    CU(c, cudaMalloc((void**)&c->dT, (size_t)(len ? len : 1) * sizeof(float)));
    CU(c, cudaMemcpy(c->dT, Tf.data(), (size_t)len * sizeof(float), cudaMemcpyHostToDevice));
    c->hT = Tf;
    return DEMCZ_OK;
}

// ------------------------------------------------------------------------------------------------
// the generation loop
// ------------------------------------------------------------------------------------------------
static int init_state_if_needed(demcz_ctx* c)
{
    const demcz_opts& o = c->o;
    if (!c->state_set) {
        if (o.init_mode == DEMCZ_INIT_REF_SERIAL_ZEROS) {           // SURVEY Q1 (src/demcz.jl:11,15)
            CU(c, cudaMemsetAsync(c->dX, 0, (size_t)c->ldx * o.d * sizeof(float), c->stream));
        } else {                                                    // last N rows of Z (src/demcz.jl:14 comment; :113)
            if (c->M < o.N_global) return demcz_fail(c, DEMCZ_ERR_INVALID, "archive has %lld rows < N=%lld chains", (long long)c->M, (long long)o.N_global);
            if (c->M > c->zcap) return demcz_fail(c, DEMCZ_ERR_INVALID, "init from a wrapped archive is not supported");
            rows_to_soa_kernel<<<grid1d(o.N), 256, 0, c->stream>>>(c->dZ, c->M - o.N_global + o.chain_offset, c->zcap, o.N, o.d,
                                                                   c->ZS, c->ldx, c->dX);
            CU(c, cudaGetLastError());
        }
        c->state_set = true;
        c->state_lp_valid = false;
    }
    if (!c->state_lp_valid) return eval_state_logp(c);
    return DEMCZ_OK;
}

// The rows a proposal may draw from, as (count, physical row of logical row 0).
//  * overlap sync mode: the newest appended segment is not drawable yet (it is still being exchanged under this
//    K-block's compute) — on every rank and for every world size, so results do not depend on the sharding;
//  * ring buffer about to wrap: the oldest rows that can be overwritten while this K-block's threads are still
//    gathering are excluded: the N_global rows of this K-block's own append and, in overlap mode, N_global more — a
//    peer may run one K-block ahead and push its NEXT segment into this replica while this rank is still drawing
//    (it cannot get further: its K-block after that needs the segment this rank appends at the end of this one).
//    Applied for every world size, so a wrapped run is still independent of the sharding and bit-reproducible.
// NCCL transport: also makes the compute stream wait (event) until the newest drawable segment has been gathered; the
// peer-push transport needs no stream operation — the gathering kernels poll the flag words (demcz_comm_need).
// The rule itself, free of any device state (exported for the CPU model check of the exchange protocol,
// tests/test_exchange_protocol_model.py): rows appended so far M, ring capacity, rows per segment, sync mode, segments
// appended so far on this archive, appends this context will still make -> (drawable rows, physical row of logical row 0).
extern "C" int demcz_archive_view_rule(int64_t M, int64_t zcap, int64_t N_global, int32_t overlap, int64_t appends, int64_t remaining,
                                       int64_t* m_out, int64_t* z0_out)
{
    if (!m_out || !z0_out || zcap <= 0 || N_global <= 0 || M < 0) return DEMCZ_ERR_INVALID;
    const int64_t pend = (overlap && appends > 0) ? N_global : 0;
    const int64_t ahead = std::max<int64_t>(0, std::min<int64_t>(overlap ? 2 : 1, remaining));
    const int64_t T = M + ahead * N_global;                   // rows [M, T) may be written while this K-block draws
    int64_t m, z0;
    if (T > zcap) {                                           // they land on the oldest rows: keep off them
        z0 = (T - zcap) % zcap;
        m = zcap - ahead * N_global - pend;
    } else if (M > zcap) {                                    // wrapped ring, no further append
        z0 = M % zcap;
        m = zcap - pend;
    } else {
        z0 = 0;
        m = M - pend;
    }
    *m_out = m;
    *z0_out = z0;
    return m >= 2 ? DEMCZ_OK : DEMCZ_ERR_INVALID;
}

int demcz_archive_view(demcz_ctx* c, uint32_t* Mdraw, uint32_t* zstart)
{
    const demcz_opts& o = c->o;
    const bool overlap = o.sync_mode == DEMCZ_SYNC_OVERLAP;
    int64_t m, z0;
    const int ok = demcz_archive_view_rule(c->M, c->zcap, o.N_global, overlap ? 1 : 0, c->appends, (int64_t)o.Ngeneration / o.K - c->appends_run, &m, &z0);
    if (ok != DEMCZ_OK) return demcz_fail(c, DEMCZ_ERR_INVALID, "demcz_run: fewer than 2 drawable archive rows (M=%lld, capacity=%lld)", (long long)c->M, (long long)c->zcap);
    *Mdraw = (uint32_t)m;
    *zstart = (uint32_t)z0;
    if (o.world > 1 && !c->peers_open && overlap && c->appends >= 2 && c->vis_waited < c->appends - 1) {
        // segments 0 .. appends-2 become drawable: the side stream's all-gather of them must be done
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_vis[(c->appends - 2) & 1], 0));
        c->vis_waited = c->appends - 1;
    }
    if (o.world > 1 && c->peers_open && !c->inkernel_poll) {
        // peer transport without in-kernel polling: a one-warp wait kernel in front of the gathering launch — and, in
        // overlap mode, the rank's OWN push of that segment must be done as well: this makes the next step kernel and the
        // next push kernel start together, so the push kernel's CTAs get their SM slots before the step kernel has filled
        // the GPU (without it the push starves behind a full wave: 8 GPUs 45.8 instead of ~70 G chain-updates/s)
        const int64_t need = overlap ? c->appends - 1 : c->appends;
        if (need > c->vis_waited) {
            if (overlap && c->appends >= 2) CU(c, cudaStreamWaitEvent(c->stream, c->ev_vis[(c->appends - 2) & 1], 0));
            int rc = demcz_comm_wait_segment(c, need);
            if (rc != DEMCZ_OK) return rc;
            c->vis_waited = need;
        }
    }
    return DEMCZ_OK;
}

// After the kernel that wrote this rank's rows of the next segment has been launched: exchange, advance M.
int demcz_archive_appended(demcz_ctx* c)
{
    const demcz_opts& o = c->o;
    const int64_t row0 = c->M % c->zcap;
    if (o.world > 1) {
        if (o.sync_mode == DEMCZ_SYNC_OVERLAP) {
            const int slot = (int)(c->appends & 1);
            if (c->last_append_by_step_kernel) {         // the side stream waits on the word the kernel's last CTA writes
                int rc = demcz_comm_wait_done_word(c, (unsigned int)(c->appends + 1));
                if (rc != DEMCZ_OK) return rc;
            } else {
                CU(c, cudaEventRecord(c->ev_app[slot], c->stream));
                for (int i = 0; i < std::max(1, c->n_xs); ++i)
                    CU(c, cudaStreamWaitEvent(c->xs[i] ? c->xs[i] : c->comm_stream, c->ev_app[slot], 0));
            }
            int rc = demcz_comm_exchange(c, row0, true);
            if (rc != DEMCZ_OK) return rc;
            CU(c, cudaEventRecord(c->ev_vis[slot], c->comm_stream));
        } else {
            int rc = demcz_comm_exchange(c, row0, false);
            if (rc != DEMCZ_OK) return rc;
        }
    }
    c->M += o.N_global;
    c->appends += 1;
    c->appends_run += 1;
    return DEMCZ_OK;
}

// History with thinning: generation grel (1-based) is stored iff grel % thin == 0, at index grel / thin - 1.  A launch
// covering generations (g, g + n_gens] therefore stores indices [g / thin, (g + n_gens) / thin).
int demcz_hist_begin(demcz_ctx* c, int64_t g, int n_gens, float** hist, void** lphist)
{
    const demcz_opts& o = c->o;
    *hist = nullptr; *lphist = nullptr;
    const int64_t G_st = (int64_t)o.Ngeneration / c->thin;              // stored generations of the whole run
    if (!o.store_chain || G_st <= 0) return DEMCZ_OK;
    const bool lin = c->lik_kind == DEMCZ_LIK_LINREG;
    if (o.store_chain == 1) {
        *hist = c->dhist;
        *lphist = lin ? (void*)demcz_linreg_lphist(c) : (void*)c->dlphist;
        return DEMCZ_OK;
    }
    // ---- streamed history ----------------------------------------------------------------------------------
    const size_t dn = (size_t)o.d * o.N;
    if (!c->hs_hx) {                                    // first launch: allocate (the likelihood kind is known by now)
        c->hs_lp_elem = lin ? sizeof(double) : sizeof(float);
        c->hs_slice_gens = std::max<int64_t>(1, std::min<int64_t>(o.K, o.Ngeneration));
        for (int i = 0; i < 2; ++i) {
            CU(c, cudaMalloc((void**)&c->hs_dx[i], (size_t)c->hs_slice_gens * dn * sizeof(float)));
            CU(c, cudaMalloc(&c->hs_dlp[i], (size_t)c->hs_slice_gens * o.N * c->hs_lp_elem));
            CU(c, cudaEventCreateWithFlags(&c->hs_ev_k[i], cudaEventDisableTiming));
            CU(c, cudaEventCreateWithFlags(&c->hs_ev_c[i], cudaEventDisableTiming));
        }
        CU(c, cudaHostAlloc((void**)&c->hs_hx, (size_t)G_st * dn * sizeof(float), cudaHostAllocDefault));
        CU(c, cudaHostAlloc(&c->hs_hlp, (size_t)G_st * o.N * c->hs_lp_elem, cudaHostAllocDefault));
        CU(c, cudaStreamCreateWithFlags(&c->hs_stream, cudaStreamNonBlocking));
    }
    if (n_gens > c->hs_slice_gens) return demcz_fail(c, DEMCZ_ERR_INVALID, "history slice of %d generations > %lld", n_gens, (long long)c->hs_slice_gens);
    const int p = (int)(c->hs_launches & 1);
    if (c->hs_launches >= 2) CU(c, cudaStreamWaitEvent(c->stream, c->hs_ev_c[p], 0));     // the slice's previous content is on the host
    const int64_t h0 = g / c->thin;                                                       // first stored index of this launch
    *hist = c->hs_dx[p] - (size_t)h0 * dn;                                                // kernels index with the absolute stored index
    *lphist = (char*)c->hs_dlp[p] - (size_t)h0 * o.N * c->hs_lp_elem;
    return DEMCZ_OK;
}

int demcz_hist_end(demcz_ctx* c, int64_t g, int n_gens)
{
    const demcz_opts& o = c->o;
    if (o.store_chain != 2 || !c->hs_hx) return DEMCZ_OK;
    const size_t dn = (size_t)o.d * o.N;
    const int p = (int)(c->hs_launches & 1);
    const int64_t h0 = g / c->thin, h1 = (g + n_gens) / c->thin;                          // stored indices [h0, h1)
    CU(c, cudaEventRecord(c->hs_ev_k[p], c->stream));
    if (h1 > h0) {
        CU(c, cudaStreamWaitEvent(c->hs_stream, c->hs_ev_k[p], 0));
        CU(c, cudaMemcpyAsync(c->hs_hx + (size_t)h0 * dn, c->hs_dx[p], (size_t)(h1 - h0) * dn * sizeof(float), cudaMemcpyDeviceToHost, c->hs_stream));
        CU(c, cudaMemcpyAsync((char*)c->hs_hlp + (size_t)h0 * o.N * c->hs_lp_elem, c->hs_dlp[p], (size_t)(h1 - h0) * o.N * c->hs_lp_elem,
                              cudaMemcpyDeviceToHost, c->hs_stream));
    }
    CU(c, cudaEventRecord(c->hs_ev_c[p], c->hs_stream));
    c->hs_launches += 1;
    return DEMCZ_OK;
}

static cudaEvent_t pool_event(demcz_ctx* c, size_t i)
{
    while (c->ev_pool.size() <= i) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        c->ev_pool.push_back(ev);
    }
    return c->ev_pool[i];
}

static int start_rhat_window(demcz_ctx* c)
{
    const demcz_opts& o = c->o;
    const size_t nx = (size_t)c->ldx * o.d;
    CU(c, cudaMemsetAsync(c->dS1[0], 0, nx * sizeof(float), c->stream));
    CU(c, cudaMemsetAsync(c->dS1[1], 0, nx * sizeof(float), c->stream));
    CU(c, cudaMemsetAsync(c->dS2, 0, (size_t)2 * o.d * sizeof(double), c->stream));
    CU(c, cudaMemsetAsync(c->d_scratch, 0, (size_t)o.d * sizeof(double), c->stream));
    colsum_kernel<<<dim3(std::min(grid1d(o.N), 64), o.d), 256, 0, c->stream>>>(c->dX, o.N, c->ldx, c->d_scratch);
    CU(c, cudaGetLastError());
    // shift = this rank's population mean now: conditioning of the FP32 variance sums only (any value near
    // the chains does).  The host needs it as a kernel parameter -> one small D2H per Rhat window.
    double sums[64];
    CU(c, cudaMemcpyAsync(sums, c->d_scratch, (size_t)o.d * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 64; ++k) c->shift[k] = k < o.d ? (float)(sums[k] / o.N) : 0.f;
    CU(c, cudaMemcpyAsync(c->dshift, c->shift, (size_t)o.d * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return DEMCZ_OK;
}

// Rhat_gelman on the finished window (src/utils.jl:2-20) from the streaming accumulators.
static int finish_rhat_window(demcz_ctx* c, int n)
{
    const demcz_opts& o = c->o;
    const int d = o.d;
    double* part = c->d_scratch;                       // [3][d] partials, then [2][d] S2 copied behind
    CU(c, cudaMemsetAsync(part, 0, (size_t)3 * d * sizeof(double), c->stream));
    rhat_partial_kernel<<<dim3(std::min(grid1d(o.N), 128), d, 2), 256, 0, c->stream>>>(c->dS1[0], c->dS1[1], o.N, c->ldx, d,
                                                                                        c->dshift, 1.0 / n, part);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(part + 3 * d, c->dS2, (size_t)2 * d * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    int rc = demcz_comm_allreduce_f64(c, part, 5 * d);
    if (rc != DEMCZ_OK) return rc;
    std::vector<double> h((size_t)5 * d);
    CU(c, cudaMemcpyAsync(h.data(), part, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const double m = 2.0 * (double)o.N_global;          // m = Npop*2  (:5)
    double mx = 0.0;
    for (int k = 0; k < d; ++k) {
        const double sum_mean = h[k], sum_mean2 = h[d + k], sum_s1sq = h[2 * d + k];
        const double s2 = h[3 * d + k] + h[4 * d + k];
        const double W = (s2 - sum_s1sq / n) / ((n - 1.0) * m);                     // :14-15
        const double avg = sum_mean / m;                                            // :10
        const double B = (double)n / (m - 1.0) * (sum_mean2 - m * avg * avg);       // :13
        const double varhat = (n - 1.0) / n * W + B / n;                            // :16
        c->rhat[k] = sqrt(varhat / W);                                              // :18
        if (!(c->rhat[k] <= mx)) mx = c->rhat[k];                                   // NaN propagates
    }
    c->rhat_max = mx;
    c->rhat_valid = true;
    return DEMCZ_OK;
}

extern "C" int demcz_run(demcz_ctx* c, int64_t max_generations, demcz_run_info* info)
{
    CHECK_CTX(c);
    const demcz_opts& o = c->o;
    if (c->lik_kind == DEMCZ_LIK_NONE) return demcz_fail(c, DEMCZ_ERR_INVALID, "demcz_run: no likelihood set");
    if (!c->archive_set) return demcz_fail(c, DEMCZ_ERR_INVALID, "demcz_run: no archive set");
    if (o.world > 1 && !c->comm) return demcz_fail(c, DEMCZ_ERR_INVALID, "demcz_run: world=%d but demcz_comm_init was not called", o.world);
    int rc = init_state_if_needed(c);
    if (rc != DEMCZ_OK) return rc;
    rc = demcz_moments_begin(c);
    if (rc != DEMCZ_OK) return rc;

    if (!c->ev_run[0]) { CU(c, cudaEventCreate(&c->ev_run[0])); CU(c, cudaEventCreate(&c->ev_run[1])); }
    CU(c, cudaEventRecord(c->ev_run[0], c->stream));

    const int64_t g_begin = c->g_done;
    int64_t g_end = o.Ngeneration;
    if (max_generations > 0 && g_begin + max_generations < g_end) g_end = g_begin + max_generations;
    const int every = o.autostop_every, nhalf = every / 2;
    size_t n_launch = 0, n_launch_untimed = 0;     // step-kernel launches (bracketed by events under profile=1) / snooker launches
    c->xchg_used = 0;
    const unsigned long long lin_launch0 = demcz_linreg_launches(c);
    if (c->lik_kind == DEMCZ_LIK_LINREG) demcz_linreg_profile_ms(c, nullptr);      // drop brackets of earlier calls
    bool stopped = false;

    int64_t g = g_begin;
    while (g < g_end && !stopped) {
        // ---- next boundary ---------------------------------------------------------------------
        // small-N persistent launches: one CTA holds every chain, the append is a CTA barrier inside the kernel, so the
        // launch need not end at a K boundary (it still ends at every host decision: Rhat halves, adaptγ, snooker)
        const bool pers = c->persistent_ok && c->lik_kind == DEMCZ_LIK_MVN;
        int64_t gb = pers ? g_end : std::min<int64_t>(g_end, (g / o.K + 1) * (int64_t)o.K);            // archive append
        bool acc = false;
        int half = 0;
        if (o.autostop) {
            const int64_t w0 = g / every * every;
            const int64_t rel = g - w0;                                                 // generations of the window done
            if (rel == 0) { rc = start_rhat_window(c); if (rc != DEMCZ_OK) return rc; }
            if (rel < nhalf)            { gb = std::min(gb, w0 + nhalf);     acc = true; half = 0; }
            else if (rel < 2 * nhalf)   { gb = std::min(gb, w0 + 2 * nhalf); acc = true; half = 1; }
            else                        { gb = std::min(gb, w0 + every); }
        }
        if (o.anneal && o.adapt) gb = std::min<int64_t>(gb, (g / o.adapt_every + 1) * (int64_t)o.adapt_every);
        // snooker generations (opts.snooker_prob) run in their own launch: stop the fused launch in front of the next one
        bool snk = false;
        if (o.snooker_prob > 0.0) {
            if (demcz_is_snooker_gen(c, o.gen_offset + g + 1)) { snk = true; gb = g + 1; }
            else for (int64_t t = g + 1; t < gb; ++t) if (demcz_is_snooker_gen(c, o.gen_offset + t + 1)) { gb = t; break; }
        }
        if (pers) {
            // no ring wrap inside a launch, at most 2^15 generations per launch (32-bit proposal counters)
            gb = std::min<int64_t>(gb, g + 32768);
            const int64_t room = (c->zcap - c->M) / o.N_global - 2;              // appends that fit well before the ring wraps (the view's wrap rules never apply)
            const int64_t last = (g / o.K + std::max<int64_t>(room, 0)) * (int64_t)o.K;     // last generation whose append still fits
            if (last <= g) gb = std::min<int64_t>(gb, (g / o.K + 1) * (int64_t)o.K);       // ring about to wrap: one K-block per launch
            else gb = std::min<int64_t>(gb, last);
        }
        const int n_gens = (int)(gb - g);
        const int64_t n_appends = gb / o.K - g / o.K;                            // appends made by this launch (pers: possibly several)
        const bool pers_launch = pers && !snk && c->M + (n_appends + 2) * o.N_global <= c->zcap;
        const bool append = pers_launch ? n_appends > 0 : (gb % o.K) == 0;

        if (c->lik_kind == DEMCZ_LIK_LINREG) {
            rc = demcz_linreg_generations(c, g, n_gens, append, acc, half, snk);
            if (rc != DEMCZ_OK) return rc;
        } else if (snk) {
            SnookerArgs a;
            memset(&a, 0, sizeof a);
            a.X = c->dX; a.logp = c->dlogp; a.N = o.N; a.ldx = c->ldx; a.d = o.d;
            a.chain_offset = (uint32_t)o.chain_offset;
            a.Z = c->dZ; a.ZS = c->ZS; a.zcap = (uint32_t)c->zcap;
            rc = demcz_archive_view(c, &a.Mdraw, &a.zstart);
            if (rc != DEMCZ_OK) return rc;
            if (a.Mdraw < 3) return demcz_fail(c, DEMCZ_ERR_INVALID, "demcz_run: the snooker update needs 3 drawable archive rows");
            a.gen = o.gen_offset + g + 1; a.grel = g + 1;
            a.keys = c->keys;
            a.gamma_s = (float)o.snooker_gamma;
            a.Ttab = o.anneal ? c->dT : nullptr;
            a.nblocks = o.Nblocks;
            { void* lph = nullptr; float* h = nullptr; rc = demcz_hist_begin(c, g, 1, &h, &lph); if (rc != DEMCZ_OK) return rc;
              const bool keep = h != nullptr && (g + 1) % c->thin == 0;
              a.hist = keep ? h : nullptr; a.lphist = (float*)lph; a.hidx = (g + 1) / c->thin - 1; }
            a.tr_r1 = c->tr_r1; a.tr_r2 = c->tr_r2; a.tr_u = c->tr_u; a.tr_xprop = c->tr_xprop; a.tr_lpprop = c->tr_lpprop; a.tr_acc = c->tr_acc;
            a.chain_changed = c->d_chain_changed; a.win_changed = c->d_counters + 0; a.accepted_total = c->d_counters + 1;
            a.count_run = (g == 0) ? 0 : 1;
            a.count_win = (o.anneal && o.adapt && g % o.adapt_every == 0) ? 0 : 1;
            if (acc) { a.S1 = c->dS1[half]; a.S2 = c->dS2 + (size_t)half * o.d; for (int k = 0; k < 32; ++k) a.shift[k] = c->shift[k]; }
            a.Zw = append ? c->dZ : nullptr;
            a.append_row0 = (uint32_t)(c->M % c->zcap);
            a.flags = c->dZ ? demcz_flags(c) : nullptr; a.world = o.world; a.my_rank = o.rank; a.need = demcz_comm_need(c);
            c->last_append_by_step_kernel = false;
            CU(c, launch_snooker(c, a));
            ++n_launch_untimed;
            rc = demcz_hist_end(c, g, 1);
            if (rc != DEMCZ_OK) return rc;
            if (o.moments) { rc = demcz_moments_add_state(c); if (rc != DEMCZ_OK) return rc; }
        } else {
            // ---- launch ------------------------------------------------------------------------
            StepArgs a;
            memset(&a, 0, sizeof a);
            a.X = c->dX; a.logp = c->dlogp; a.N = o.N; a.ldx = c->ldx; a.d = o.d;
            a.chain_offset = (uint32_t)o.chain_offset;
            a.Z = c->dZ;
            a.zcap = (uint32_t)c->zcap;
            rc = demcz_archive_view(c, &a.Mdraw, &a.zstart);
            if (rc != DEMCZ_OK) return rc;
            a.gen0 = o.gen_offset + g + 1;
            a.grel0 = g + 1;
            a.n_gens = n_gens;
            a.keys = c->keys;
            a.nblocks = o.Nblocks;
            for (int b = 0; b < o.Nblocks; ++b) {
                uint32_t mask = 0;
                for (int t = c->block_ptr[b]; t < c->block_ptr[b + 1]; ++t) mask |= 1u << c->block_idx[t];
                const int blocklen = c->block_ptr[b + 1] - c->block_ptr[b];
                a.blocks[b].mask = mask;
                a.blocks[b].cscale = (float)(blocklen == 1 ? c->gamma : c->gamma / sqrt(2.0 * blocklen));   // src/demcz.jl:183-187
            }
            for (int k = 0; k < o.d; ++k) (&a.eps2[0].x)[k] = (float)c->eps[k];
            a.Ttab = o.anneal ? c->dT : nullptr;
            a.Zw = append ? c->dZ : nullptr;
            a.append_row0 = (uint32_t)(c->M % c->zcap);
            { void* lph = nullptr; rc = demcz_hist_begin(c, g, n_gens, &a.hist, &lph); if (rc != DEMCZ_OK) return rc; a.lphist = (float*)lph; }
            a.tr_r1 = c->tr_r1; a.tr_r2 = c->tr_r2; a.tr_u = c->tr_u; a.tr_xprop = c->tr_xprop;
            a.tr_lpprop = c->tr_lpprop; a.tr_acc = c->tr_acc;
            a.chain_changed = c->d_chain_changed;
            a.win_changed = c->d_counters + 0;
            a.accepted_total = c->d_counters + 1;
            a.run_skip_gi = (g == 0) ? 0 : -1;
            a.win_skip_gi = (o.anneal && o.adapt && g % o.adapt_every == 0) ? 0 : -1;
            if (acc) { a.S1 = c->dS1[half]; a.S2 = c->dS2 + (size_t)half * o.d; memcpy(a.shift2, c->shift, sizeof a.shift2); }

            const bool full = o.Nblocks == 1 && a.blocks[0].mask == (o.d == 32 ? 0xFFFFFFFFu : ((1u << o.d) - 1u));
            a.thin = c->thin; a.hist_phase = (int)(g % c->thin); a.hist_idx0 = (unsigned int)(g / c->thin);
            a.flags = c->dZ ? demcz_flags(c) : nullptr; a.world = o.world; a.my_rank = o.rank;
            a.need = demcz_comm_need(c);
            if (o.moments) { a.mom = c->d_mom; memcpy(a.mom_shift2, c->mom_shift, sizeof a.mom_shift2); }
            if (pers_launch) {
                a.pers_K = o.K; a.pers_kphase = (int)(g % o.K);
                a.pers_delay = (o.sync_mode == DEMCZ_SYNC_OVERLAP && c->appends == 0) ? 1 : 0;
                // overlap mode: the view handed over by demcz_archive_view already hides the newest segment; every further
                // append of the launch reveals the one before it
            }
            c->last_append_by_step_kernel = false;
            if (append && o.world > 1 && c->memop_wait && o.sync_mode == DEMCZ_SYNC_OVERLAP) {
                a.done_word = (unsigned int*)(c->d_barrier + 4);
                a.done_ctr = (unsigned int*)(c->d_barrier + 5);
                a.done_seq = (unsigned int)(c->appends + 1);
                c->last_append_by_step_kernel = true;
            }
            if (o.profile) CU(c, cudaEventRecord(pool_event(c, 2 * n_launch), c->stream));
            CU(c, demcz_launch_step(c, a, acc, full));
            if (o.profile) CU(c, cudaEventRecord(pool_event(c, 2 * n_launch + 1), c->stream));
            ++n_launch;
            rc = demcz_hist_end(c, g, n_gens);
            if (rc != DEMCZ_OK) return rc;
            if (o.moments) c->mom_gens += n_gens;          // accumulated inside the step kernel (MOM instantiation)
        }
        c->proposals += snk ? (int64_t)o.N : (int64_t)n_gens * o.Nblocks * o.N;        // a snooker generation is one proposal per chain

        // ---- archive append + exchange (src/demcz.jl:88-91; sync mode :129-137) -----------------
        if (append) {
            for (int64_t t = 0; t < (pers_launch ? n_appends : 1); ++t) {
                rc = demcz_archive_appended(c);
                if (rc != DEMCZ_OK) return rc;
            }
        }
        g = gb;
        c->g_done = g;

        // ---- adaptγ (src/demcz_anneal.jl:48-57) --------------------------------------------------
        if (o.anneal && o.adapt && g % o.adapt_every == 0) {
            rc = demcz_comm_allreduce_u64(c, c->d_counters + 0, 1);
            if (rc != DEMCZ_OK) return rc;
            unsigned long long cnt = 0;
            CU(c, cudaMemcpyAsync(&cnt, c->d_counters + 0, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            CU(c, cudaMemsetAsync(c->d_counters + 0, 0, sizeof cnt, c->stream));
            const double ar = (double)cnt / ((double)o.N_global * o.adapt_every);
            if (ar < 0.1) c->gamma = std::max(o.min_gamma, c->gamma * 0.5);
            else if (ar > 0.5) c->gamma = std::min(o.max_gamma, c->gamma * 1.5);
        }
        // ---- Rhat autostop (src/demcz.jl:39-55) --------------------------------------------------
        if (o.autostop && !o.anneal && g % every == 0) {
            rc = finish_rhat_window(c, nhalf);
            if (rc != DEMCZ_OK) return rc;
            if (c->rhat_max < o.autostop_Rhat) stopped = true;
        }
    }
    if (o.world > 1) {                                                       // drain the exchange: the archive is complete when the call returns
        for (int i = 0; i < 8; ++i) if (c->xs[i]) CU(c, cudaStreamSynchronize(c->xs[i]));      // this rank's own transfers
        if (o.sync_mode == DEMCZ_SYNC_OVERLAP && c->appends > 0) CU(c, cudaStreamWaitEvent(c->stream, c->ev_vis[(c->appends - 1) & 1], 0));
        rc = demcz_comm_wait_segment(c, c->appends);
        if (rc != DEMCZ_OK) return rc;
    }
    CU(c, cudaEventRecord(c->ev_run[1], c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->comm_stream) CU(c, cudaStreamSynchronize(c->comm_stream));
    if (c->hs_stream) CU(c, cudaStreamSynchronize(c->hs_stream));                 // streamed history is complete on the host
    if (stopped) c->stopped_by_rhat = true;

    if (info) {
        memset(info, 0, sizeof *info);
        info->generations_done = c->g_done - g_begin;
        info->generations_total = c->g_done;
        info->stopped_by_rhat = stopped ? 1 : 0;
        info->rhat_valid = c->rhat_valid ? 1 : 0;
        info->rhat_max = c->rhat_valid ? c->rhat_max : NAN;
        info->final_gamma = c->gamma;
        info->M = c->M;
        info->proposals = c->proposals;
        unsigned long long acc_total = 0;
        CU(c, cudaMemcpy(&acc_total, c->d_counters + 1, sizeof acc_total, cudaMemcpyDeviceToHost));
        info->accepted = (int64_t)acc_total;
        info->step_launches = (int64_t)(n_launch + n_launch_untimed) + (int64_t)(demcz_linreg_launches(c) - lin_launch0);
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, c->ev_run[0], c->ev_run[1]));
        info->device_ms = ms;
        if (o.profile) {
            double tot = 0.0;
            for (size_t i = 0; i < n_launch; ++i) {
                float t = 0.f;
                CU(c, cudaEventElapsedTime(&t, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]));
                tot += t;
            }
            if (c->lik_kind == DEMCZ_LIK_LINREG) tot = demcz_linreg_profile_ms(c, nullptr);
            info->step_kernel_ms = tot;
            double xt = 0.0;
            for (size_t i = 0; i + 1 < c->xchg_used; i += 2) {
                float t = 0.f;
                if (cudaEventElapsedTime(&t, c->ev_xchg[i], c->ev_xchg[i + 1]) == cudaSuccess) xt += t;
            }
            info->exchange_ms = xt;
            info->exchanges = (int64_t)(c->xchg_used / 2);
        }
    }
    return DEMCZ_OK;
}

extern "C" int demcz_get_kernel_info(demcz_ctx* c, char* buf, int64_t len)
{
    if (!c) return DEMCZ_ERR_INVALID;
    std::string s = demcz_linreg_info(c);
    s += c->persistent_ok ? "step_persistent=1;" : "step_persistent=0;";
    if (c->o.world > 1) {
        s += !c->peers_open ? "exchange=nccl_allgather;" : c->push_engine == 1 ? "exchange=copy_engine;" : "exchange=push_kernel;";
        s += c->memop_wait ? "memop_wait=1;" : "memop_wait=0;";
        s += c->memop_write ? "memop_write=1;" : "memop_write=0;";
    }
    if (buf && len > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)len - 1);
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int)s.size() + 1;
}

extern "C" int demcz_set_profile(demcz_ctx* c, int32_t on)
{
    CHECK_CTX(c);
    c->o.profile = on ? 1 : 0;
    return DEMCZ_OK;
}

// ------------------------------------------------------------------------------------------------
// ABI: results
// ------------------------------------------------------------------------------------------------
static int d2h_as_f64(demcz_ctx* c, const float* src, size_t n, double* dst)
{
    double* stage = nullptr;
    CU(c, cudaMalloc((void**)&stage, n * sizeof(double)));
    f32_to_f64_kernel<<<grid1d((long long)n), 256, 0, c->stream>>>(src, (long long)n, stage);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, stage, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(stage);
    CU(c, e);
    return DEMCZ_OK;
}

extern "C" int demcz_get_chain(demcz_ctx* c, double* chain, double* log_obj)
{
    CHECK_CTX(c);
    const size_t G = (size_t)(c->g_done / c->thin), N = c->o.N, d = c->o.d;      // stored generations
    if (!c->o.store_chain) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_chain: store_chain was 0");
    if (G == 0) return DEMCZ_OK;
    if (c->o.store_chain == 2) {                        // streamed: FP32 (FP64 log-objective for linreg) pinned host copy -> caller's Float64
        if (!c->hs_hx) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_chain: no history yet");
        CU(c, cudaStreamSynchronize(c->hs_stream));
        if (chain) for (size_t t = 0; t < G * d * N; ++t) chain[t] = (double)c->hs_hx[t];
        if (log_obj) {
            if (c->hs_lp_elem == sizeof(double)) memcpy(log_obj, c->hs_hlp, G * N * sizeof(double));
            else for (size_t t = 0; t < G * N; ++t) log_obj[t] = (double)((const float*)c->hs_hlp)[t];
        }
        return DEMCZ_OK;
    }
    if (!c->dhist) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_chain: store_chain was 0");
    // device history is [g][k][chain] in FP32 == Julia's column-major N x d x G
    if (chain) { int rc = d2h_as_f64(c, c->dhist, G * d * N, chain); if (rc) return rc; }
    if (log_obj && c->lik_kind == DEMCZ_LIK_LINREG) {
        CU(c, cudaMemcpyAsync(log_obj, demcz_linreg_lphist(c), G * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    } else if (log_obj) { int rc = d2h_as_f64(c, c->dlphist, G * N, log_obj); if (rc) return rc; }
    return DEMCZ_OK;
}

extern "C" int demcz_get_state(demcz_ctx* c, double* X, double* logp)
{
    CHECK_CTX(c);
    if (!c->state_set && !c->archive_set) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_state: no state and no archive");
    { int rc = init_state_if_needed(c); if (rc) return rc; }   // also evaluates logp after set_state(X, NULL)
    const int N = c->o.N, d = c->o.d;
    double* stage = c->d_stage;
    if (X) {
        export_soa_kernel<<<grid1d(N), 256, 0, c->stream>>>(c->dX, N, d, c->ldx, stage);
        CU(c, cudaGetLastError());
        CU(c, cudaMemcpyAsync(X, stage, (size_t)N * d * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    if (logp && c->lik_kind == DEMCZ_LIK_LINREG) {
        CU(c, cudaMemcpyAsync(logp, demcz_linreg_lp(c), (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else if (logp) {
        f32_to_f64_kernel<<<grid1d(N), 256, 0, c->stream>>>(c->dlogp, N, stage + (size_t)N * d);
        CU(c, cudaGetLastError());
        CU(c, cudaMemcpyAsync(logp, stage + (size_t)N * d, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return DEMCZ_OK;
}

extern "C" int demcz_get_archive(demcz_ctx* c, double* Z, int64_t* M)
{
    CHECK_CTX(c);
    if (!c->archive_set) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_archive: no archive");
    const int64_t rows = std::min<int64_t>(c->M, c->zcap);
    if (M) *M = rows;
    if (!Z) return DEMCZ_OK;
    double* stage = nullptr;
    CU(c, cudaMalloc((void**)&stage, (size_t)rows * c->o.d * sizeof(double)));
    const int64_t zstart = c->M > c->zcap ? c->M % c->zcap : 0;
    export_rows_kernel<<<grid1d(rows), 256, 0, c->stream>>>(c->dZ, rows, zstart, c->zcap, c->o.d, c->ZS, stage);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(Z, stage, (size_t)rows * c->o.d * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(stage);
    CU(c, e);
    return DEMCZ_OK;
}

extern "C" int demcz_get_moments(demcz_ctx* c, double* mean, double* cov)
{
    CHECK_CTX(c);
    const int d = c->o.d, N = c->o.N;
    double* sums = c->d_scratch;            // [d] sums, [d] mean, [d*d] cov
    double* dmean = sums + d;
    double* dcov = dmean + d;
    CU(c, cudaMemsetAsync(sums, 0, (size_t)(2 * d + d * d) * sizeof(double), c->stream));
    colsum_kernel<<<dim3(std::min(grid1d(N), 64), d), 256, 0, c->stream>>>(c->dX, N, c->ldx, sums);
    int rc = demcz_comm_allreduce_f64(c, sums, d);
    if (rc) return rc;
    shift_from_sums_kernel<<<1, 64, 0, c->stream>>>(sums, d, 1.0 / (double)c->o.N_global, nullptr, dmean);
    cov_kernel<<<dim3(std::min(grid1d(N), 64), d * d), 256, 0, c->stream>>>(c->dX, N, c->ldx, d, dmean, dcov);
    CU(c, cudaGetLastError());
    rc = demcz_comm_allreduce_f64(c, dcov, d * d);
    if (rc) return rc;
    std::vector<double> h((size_t)d + d * d);
    CU(c, cudaMemcpyAsync(h.data(), dmean, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (mean) for (int k = 0; k < d; ++k) mean[k] = h[k];
    if (cov) for (int t = 0; t < d * d; ++t) cov[t] = h[d + t] / (double)c->o.N_global;
    return DEMCZ_OK;
}

extern "C" int demcz_get_archive_moments(demcz_ctx* c, int64_t first_row, int64_t n_rows, double* mean, double* cov)
{
    CHECK_CTX(c);
    if (!c->archive_set) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_archive_moments: no archive");
    const int64_t rows = std::min<int64_t>(c->M, c->zcap);
    if (first_row < 0 || n_rows < 1 || first_row + n_rows > rows)
        return demcz_fail(c, DEMCZ_ERR_INVALID, "get_archive_moments: rows [%lld, %lld) outside the %lld valid rows", (long long)first_row,
                          (long long)(first_row + n_rows), (long long)rows);
    const int d = c->o.d;
    const int64_t zstart = c->M > c->zcap ? c->M % c->zcap : 0;
    double* sums = c->d_scratch;            // [d] sums, [d] mean, [d*d] cov
    double* dmean = sums + d;
    double* dcov = dmean + d;
    CU(c, cudaMemsetAsync(sums, 0, (size_t)(2 * d + d * d) * sizeof(double), c->stream));
    const int blocks = (int)std::min<int64_t>((n_rows + 255) / 256, 256);
    rows_colsum_kernel<<<dim3(blocks, d), 256, 0, c->stream>>>(c->dZ, first_row, n_rows, zstart, c->zcap, c->ZS, sums);
    shift_from_sums_kernel<<<1, 64, 0, c->stream>>>(sums, d, 1.0 / (double)n_rows, nullptr, dmean);
    rows_cov_kernel<<<dim3(blocks, d * d), 256, 0, c->stream>>>(c->dZ, first_row, n_rows, zstart, c->zcap, c->ZS, d, dmean, dcov);
    CU(c, cudaGetLastError());
    std::vector<double> h((size_t)d + d * d);
    CU(c, cudaMemcpyAsync(h.data(), dmean, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (mean) for (int k = 0; k < d; ++k) mean[k] = h[k];
    if (cov) for (int t = 0; t < d * d; ++t) cov[t] = h[d + t] / (double)n_rows;
    return DEMCZ_OK;
}

// ------------------------------------------------------------------------------------------------
// streaming moments of the full history (opts.moments): flatten_chain + mean_cov_chain (src/utils.jl:22-32, 96-106)
// ------------------------------------------------------------------------------------------------
int demcz_moments_begin(demcz_ctx* c)
{
    const demcz_opts& o = c->o;
    if (!o.moments || c->mom_shift_set) return DEMCZ_OK;
    if (!c->d_mom) {
        CU(c, cudaMalloc((void**)&c->d_mom, (64 + 64 * 64) * sizeof(double)));
        CU(c, cudaMemsetAsync(c->d_mom, 0, (64 + 64 * 64) * sizeof(double), c->stream));
    }
    // shift = population mean of the state the run starts from (conditioning of the FP32 products only; the same on
    // every rank: all-reduced), fixed for the life of the accumulators
    CU(c, cudaMemsetAsync(c->d_scratch, 0, (size_t)o.d * sizeof(double), c->stream));
    colsum_kernel<<<dim3(std::min(grid1d(o.N), 64), o.d), 256, 0, c->stream>>>(c->dX, o.N, c->ldx, c->d_scratch);
    CU(c, cudaGetLastError());
    int rc = demcz_comm_allreduce_f64(c, c->d_scratch, o.d);
    if (rc != DEMCZ_OK) return rc;
    double sums[64];
    CU(c, cudaMemcpyAsync(sums, c->d_scratch, (size_t)o.d * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 64; ++k) c->mom_shift[k] = k < o.d ? (float)(sums[k] / (double)o.N_global) : 0.f;
    if (!c->dmomshift) CU(c, cudaMalloc((void**)&c->dmomshift, 64 * sizeof(float)));
    CU(c, cudaMemcpyAsync(c->dmomshift, c->mom_shift, 64 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->mom_shift_set = true;
    return DEMCZ_OK;
}

int demcz_moments_add_state(demcz_ctx* c)
{
    const demcz_opts& o = c->o;
    state_moments_kernel<<<dim3(std::min(grid1d(o.N), 32), o.d * o.d + o.d), 256, 0, c->stream>>>(c->dX, o.N, c->ldx, o.d, c->dmomshift, c->d_mom);
    CU(c, cudaGetLastError());
    c->mom_gens += 1;
    return DEMCZ_OK;
}

int demcz_moments_get(demcz_ctx* c, double* mean, double* cov, int64_t* n, int32_t reset)
{
    const demcz_opts& o = c->o;
    const int d = o.d;
    if (n) *n = c->mom_gens * o.N_global;
    if (!c->d_mom || c->mom_gens == 0) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_history_moments: no generation accumulated yet");
    double* tmp = nullptr;                                       // all-reduce a copy: the accumulators stay per-rank
    CU(c, cudaMalloc((void**)&tmp, (64 + 64 * 64) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(tmp, c->d_mom, (64 + 64 * 64) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    int rc = DEMCZ_OK;
    if (e == cudaSuccess) rc = demcz_comm_allreduce_f64(c, tmp, 64 + 64 * 64);
    std::vector<double> h(64 + 64 * 64);
    if (e == cudaSuccess && rc == DEMCZ_OK) e = cudaMemcpyAsync(h.data(), tmp, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    if (rc != DEMCZ_OK) return rc;
    CU(c, e);
    const double inv = 1.0 / ((double)c->mom_gens * (double)o.N_global);
    for (int k = 0; k < d; ++k) if (mean) mean[k] = (double)c->mom_shift[k] + h[k] * inv;
    if (cov)
        for (int k1 = 0; k1 < d; ++k1)
            for (int k2 = 0; k2 < d; ++k2) {
                const int a = std::min(k1, k2), b = std::max(k1, k2);                 // upper triangle is accumulated
                cov[k1 + (size_t)d * k2] = h[64 + a * 64 + b] * inv - (h[k1] * inv) * (h[k2] * inv);
            }
    if (reset) {
        CU(c, cudaMemsetAsync(c->d_mom, 0, (64 + 64 * 64) * sizeof(double), c->stream));
        c->mom_gens = 0;
    }
    return DEMCZ_OK;
}

extern "C" int demcz_get_history_moments(demcz_ctx* c, double* mean, double* cov, int64_t* n, int32_t reset)
{
    CHECK_CTX(c);
    if (!c->o.moments) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_history_moments: opts.moments was 0");
    return demcz_moments_get(c, mean, cov, n, reset);
}

extern "C" int demcz_get_archive_checksum(demcz_ctx* c, int64_t first_row, int64_t n_rows, uint64_t* checksum)
{
    CHECK_CTX(c);
    if (!c->archive_set || !checksum) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_archive_checksum: no archive / null output");
    const int64_t rows = std::min<int64_t>(c->M, c->zcap);
    if (first_row < 0 || n_rows < 0 || first_row + n_rows > rows)
        return demcz_fail(c, DEMCZ_ERR_INVALID, "get_archive_checksum: rows [%lld, %lld) outside the %lld valid rows", (long long)first_row,
                          (long long)(first_row + n_rows), (long long)rows);
    const int64_t zstart = c->M > c->zcap ? c->M % c->zcap : 0;
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(c->d_scratch);
    CU(c, cudaMemsetAsync(acc, 0, sizeof *acc, c->stream));
    if (n_rows > 0) {
        const int blocks = (int)std::min<int64_t>((n_rows + 255) / 256, 148 * 8);
        rows_checksum_kernel<<<blocks, 256, 0, c->stream>>>(c->dZ, first_row, n_rows, zstart, c->zcap, c->ZS, c->o.d, acc);
        CU(c, cudaGetLastError());
    }
    unsigned long long h = 0;
    CU(c, cudaMemcpyAsync(&h, acc, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    *checksum = h;
    return DEMCZ_OK;
}

extern "C" int demcz_get_rhat(demcz_ctx* c, double* rhat)
{
    CHECK_CTX(c);
    if (!rhat) return DEMCZ_ERR_INVALID;
    for (int k = 0; k < c->o.d; ++k) rhat[k] = c->rhat[k];
    return c->rhat_valid ? DEMCZ_OK : demcz_fail(c, DEMCZ_ERR_INVALID, "get_rhat: no Rhat window finished yet");
}

extern "C" int demcz_get_accept_counts(demcz_ctx* c, int64_t* per_chain)
{
    CHECK_CTX(c);
    if (!per_chain) return DEMCZ_ERR_INVALID;
    std::vector<int> h((size_t)c->o.N);
    CU(c, cudaMemcpy(h.data(), c->d_chain_changed, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); ++i) per_chain[i] = h[i];
    return DEMCZ_OK;
}

extern "C" int demcz_eval_logdensity(demcz_ctx* c, const double* X, int64_t n, double* out)
{
    CHECK_CTX(c);
    if (!X || !out || n < 1) return demcz_fail(c, DEMCZ_ERR_INVALID, "eval_logdensity: bad arguments");
    if (c->lik_kind == DEMCZ_LIK_NONE) return demcz_fail(c, DEMCZ_ERR_INVALID, "eval_logdensity: no likelihood set");
    const int d = c->o.d;
    const int64_t ld = (n + 31) / 32 * 32;
    double* stage = nullptr;
    float* xf = nullptr;
    CU(c, cudaMalloc((void**)&stage, (size_t)n * (d + 1) * sizeof(double)));
    cudaError_t e = cudaMalloc((void**)&xf, (size_t)ld * (d + 1) * sizeof(float));
    int rc = DEMCZ_OK;
    if (e == cudaSuccess) e = cudaMemsetAsync(xf, 0, (size_t)ld * (d + 1) * sizeof(float), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(stage, X, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        import_soa_kernel<<<grid1d(n), 256, 0, c->stream>>>(stage, n, d, ld, xf);
        float* of = xf + (size_t)ld * d;
        if (c->lik_kind == DEMCZ_LIK_MVN) {
            e = launch_eval(c, xf, (int)n, (int)ld, of);
            if (e == cudaSuccess) {
                f32_to_f64_kernel<<<grid1d(n), 256, 0, c->stream>>>(of, n, stage + (size_t)n * d);
                e = cudaGetLastError();
            }
        } else {
            rc = demcz_linreg_eval(c, xf, (int)n, (int)ld, stage + (size_t)n * d);
        }
    }
    if (e == cudaSuccess && rc == DEMCZ_OK) e = cudaMemcpyAsync(out, stage + (size_t)n * d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(stage);
    if (xf) cudaFree(xf);
    if (rc != DEMCZ_OK) return rc;
    CU(c, e);
    return DEMCZ_OK;
}

extern "C" int demcz_get_trace(demcz_ctx* c, uint32_t* r1, uint32_t* r2, float* u, float* xprop, float* lp_prop,
                               uint8_t* accepted)
{
    CHECK_CTX(c);
    if (!c->o.trace || !c->tr_r1) return demcz_fail(c, DEMCZ_ERR_INVALID, "get_trace: trace was 0");
    const size_t np = (size_t)c->g_done * c->o.Nblocks * c->o.N;
    if (np == 0) return DEMCZ_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    if (r1) CU(c, cudaMemcpy(r1, c->tr_r1, np * 4, cudaMemcpyDeviceToHost));
    if (r2) CU(c, cudaMemcpy(r2, c->tr_r2, np * 4, cudaMemcpyDeviceToHost));
    if (u) CU(c, cudaMemcpy(u, c->tr_u, np * 4, cudaMemcpyDeviceToHost));
    if (lp_prop) CU(c, cudaMemcpy(lp_prop, c->tr_lpprop, np * 4, cudaMemcpyDeviceToHost));
    if (accepted) CU(c, cudaMemcpy(accepted, c->tr_acc, np, cudaMemcpyDeviceToHost));
    if (xprop) CU(c, cudaMemcpy(xprop, c->tr_xprop, np * c->o.d * 4, cudaMemcpyDeviceToHost));
    return DEMCZ_OK;
}
