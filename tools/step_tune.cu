// step_tune.cu — times structural variants (OPT bits x resident blocks) of the fused DEMCz step kernel on the
// C4 shape (d=20, 131072 chains, 10 generations per launch, archive >> L2).  Development tool, not product:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o step_tune tools/step_tune.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../demc.jl_b200/csrc/demcz_step.cuh"
using namespace demcz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int LD = 20;

template <int MINB, int OPT, int WARPS = 4>
float run_variant(StepArgs a, const std::vector<float>& Wfull, const std::vector<float>& mu, int reps, double* checksum)
{
    using MP = MvnParams<LD>;
    MvnParams<LD> P;
    memset(&P, 0, sizeof P);
    auto W = [&](int i, int j) { return j <= i ? Wfull[i * LD + j] : 0.f; };
    for (int k = 0; k < LD / 2; ++k) P.mu2[k] = make_float2(mu[2 * k], mu[2 * k + 1]);
    if (OPT & OPT_JMAJOR) {
        for (int j = 0; j < LD; ++j)
            for (int i2 = j / 2; i2 < LD / 2; ++i2) P.W[MP::wj_off(j) + i2 - j / 2] = make_float2(W(2 * i2, j), W(2 * i2 + 1, j));
    } else {
        for (int i2 = 0; i2 < LD / 2; ++i2)
            for (int j = 0; j < 2 * i2 + 2; ++j) P.W[i2 * (i2 + 1) + j] = make_float2(W(2 * i2, j), W(2 * i2 + 1, j));
    }
    P.c0 = -3.f;
    auto kern = demcz_step_mvn_kernel<LD, true, true, WARPS, MINB, OPT>;
    const size_t smem = (size_t)WARPS * 2 * StageLayout<LD>::FLOATS * sizeof(float);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
    const int grid = (a.N + WARPS * 32 - 1) / (WARPS * 32);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // same starting state for every variant
    CK(cudaMemset(a.S1, 0, (size_t)a.ldx * LD * 4));
    for (int w = 0; w < 3; ++w) { a.gen0 += 10; a.grel0 += 10; kern<<<grid, WARPS * 32, smem>>>(a, P); }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) { a.gen0 += 10; a.grel0 += 10; kern<<<grid, WARPS * 32, smem>>>(a, P); }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<float> lp(a.N);
    CK(cudaMemcpy(lp.data(), a.logp, a.N * 4, cudaMemcpyDeviceToHost));
    double cs = 0; for (float v : lp) cs += v;
    *checksum = cs;
    printf("OPT=%2d (Wsmem=%d jmajor=%d idxvec=%d s2smem=%d) WARPS=%d MINB=%d occ=%d blocks/SM N=%d : %8.2f us/launch  checksum %.6e\n", OPT,
           !!(OPT & 1), !!(OPT & 2), !!(OPT & 4), !!(OPT & 8), WARPS, MINB, occ, a.N, 1e3f * ms / reps, cs);
    fflush(stdout);
    return ms / reps;
}

// Upper bound probe: nothing but the random row fetch (C 16-byte pieces of `U` independent random rows per thread in
// flight), no Philox, no staging, no likelihood.  rows/s here is what the HBM system delivers for this access pattern.
template <int C, int U>
__global__ void gather_probe_kernel(const float* __restrict__ Z, uint32_t M, int ZS, int rounds, float* __restrict__ out)
{
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float4 v[U][C];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            h = h * 1664525u + 1013904223u;
            const uint32_t row = __umulhi(h ^ (h >> 15), M);
            const float4* p = reinterpret_cast<const float4*>(Z + (size_t)row * ZS);
#pragma unroll
            for (int c = 0; c < C; ++c) v[u][c] = __ldcg(p + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < C; ++c) acc += v[u][c].x + v[u][c].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int C, int U>
void run_probe(const float* Z, uint32_t M, int ZS, float* out, const char* what)
{
    const int threads = 256, blocks = 148 * 8, rounds = 16;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather_probe_kernel<C, U><<<blocks, threads>>>(Z, M, ZS, rounds, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) gather_probe_kernel<C, U><<<blocks, threads>>>(Z, M, ZS, rounds, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double rows = 5.0 * blocks * threads * rounds * U;
    printf("gather probe %-28s: %7.2f G rows/s  = %6.0f GB/s of 128-B line fills, %6.0f GB/s of row payload (%d B)\n", what,
           rows / (ms * 1e-3) / 1e9, rows * 128 / (ms * 1e-3) / 1e9, rows * C * 16 / (ms * 1e-3) / 1e9, C * 16);
}

int main(int argc, char** argv)
{
    const int N = 131072, d = 20, ldx = N;
    const size_t M0 = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)(8u << 20);
    const int ZS = zslot(LD);
    if (argc > 2) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[2]))); }
    { size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", g); }
    srand(1);
    std::vector<float> W(LD * LD, 0.f), mu(LD, 0.f);
    for (int i = 0; i < d; ++i) { mu[i] = 0.1f * i; for (int j = 0; j <= i; ++j) W[i * LD + j] = (i == j ? 2.0f : 0.3f * ((rand() % 200) / 100.f - 1.f)); }
    float *Z, *X, *logp, *S1; double* S2; int* cc; unsigned long long* ctr;
    CK(cudaMalloc(&Z, M0 * ZS * 4)); CK(cudaMalloc(&X, (size_t)ldx * LD * 4)); CK(cudaMalloc(&logp, N * 4));
    CK(cudaMalloc(&S1, (size_t)ldx * LD * 4)); CK(cudaMalloc(&S2, 64 * 8)); CK(cudaMalloc(&cc, N * 4)); CK(cudaMalloc(&ctr, 64));
    CK(cudaMemset(S2, 0, 64 * 8)); CK(cudaMemset(cc, 0, N * 4)); CK(cudaMemset(ctr, 0, 64));
    {
        std::vector<float> h(M0 * ZS);
        for (size_t i = 0; i < h.size(); ++i) h[i] = ((i % ZS) < (size_t)d) ? mu[i % ZS] + 0.5f * ((rand() % 2001) / 1000.f - 1.f) : 0.f;
        CK(cudaMemcpy(Z, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    std::vector<float> hx((size_t)ldx * LD), hl(N, -1e3f);
    for (int k = 0; k < LD; ++k) for (int i = 0; i < N; ++i) hx[(size_t)k * ldx + i] = k < d ? mu[k] + 0.3f * ((rand() % 2001) / 1000.f - 1.f) : 0.f;

    StepArgs a;
    memset(&a, 0, sizeof a);
    a.X = X; a.logp = logp; a.N = N; a.ldx = ldx; a.d = d; a.chain_offset = 0;
    a.Z = Z; a.Mdraw = (uint32_t)M0; a.zstart = 0; a.zcap = (uint32_t)M0;
    a.gen0 = 1; a.grel0 = 1; a.n_gens = 10; a.keys = make_philox_keys(31953150ull); a.nblocks = 1;
    a.blocks[0].mask = (1u << d) - 1; a.blocks[0].cscale = 2.38f / sqrtf(2.f * d);
    for (int k = 0; k < d; ++k) (&a.eps2[0].x)[k] = 1e-5f;
    a.chain_changed = cc; a.win_changed = ctr; a.accepted_total = ctr + 1; a.run_skip_gi = -1; a.win_skip_gi = -1;
    a.S1 = S1; a.S2 = S2;
    for (int k = 0; k < d; ++k) (&a.shift2[0].x)[k] = mu[k];

    auto reset = [&]() { CK(cudaMemcpy(X, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(logp, hl.data(), N * 4, cudaMemcpyHostToDevice)); };
    run_probe<5, 4>(Z, (uint32_t)M0, ZS, S1, "80-B rows, 4 rows in flight");
    run_probe<5, 8>(Z, (uint32_t)M0, ZS, S1, "80-B rows, 8 rows in flight");
    run_probe<8, 8>(Z, (uint32_t)M0, ZS, S1, "128-B rows, 8 rows in flight");
    run_probe<1, 16>(Z, (uint32_t)M0, ZS, S1, "16-B piece, 16 in flight");
    double cs;
    const int reps = 20;
#define RUN(MINB, OPT) reset(); run_variant<MINB, OPT>(a, W, mu, reps, &cs);
    RUN(4, 0)
    { StepArgs b = a; b.Zw = Z; b.append_row0 = 1000; reset(); printf("with append: "); run_variant<4, 0>(b, W, mu, reps, &cs); }
    { StepArgs b = a; b.Mdraw = (uint32_t)(M0 / 2); reset(); printf("draw from half: "); run_variant<4, 0>(b, W, mu, reps, &cs); }
    { StepArgs b = a; b.hist = nullptr; b.S1 = S1; reset(); printf("accept-heavy state: "); for (int k = 0; k < 20; ++k) (&b.eps2[0].x)[k] = 0.5f; run_variant<4, 0>(b, W, mu, reps, &cs); }
    return 0;
}
