// push_bench.cu — what can SM-issued stores into a peer GPU's memory sustain over NVLink on this box?
// Development tool (not product): one process, devices 0 and 1 with peer access.  Emulates the egress of the
// 8-GPU C4 archive exchange (131072 rows of 80 payload bytes in 128-byte slots, to P = 7 destinations) and
// times store patterns in isolation and next to a DRAM-saturating random gather on both devices.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/push_bench tools/push_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Dsts { float* p[8]; };

// pattern A (what demcz_comm.cu does): consecutive lanes copy consecutive 16-byte pieces, C pieces per row,
// rows in slots of ZS floats at the source; destination slot stride DS floats (DS = ZS: in place; DS = 4*C: dense).
template <int U>
__global__ void __launch_bounds__(512)
push_coop(const float* __restrict__ src, Dsts dst, int np, int nrows, int ZS, int DS, int C)
{
    const long long total = (long long)nrows * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * U) {
        float4 v[U];
        size_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long t = t0 + u * stride;
            const long long row = t / C;
            const int piece = (int)(t - row * C);
            off[u] = (size_t)row * DS + piece * 4;
            if (t < total) v[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)row * ZS + piece * 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t0 + u * stride < total)
                for (int r = 0; r < np; ++r) *reinterpret_cast<float4*>(dst.p[r] + off[u]) = v[u];
    }
}

// pattern B: lane = row (what a fused step-kernel epilogue would do without staging): C 16-byte stores at slot stride
__global__ void __launch_bounds__(512)
push_lane_per_row(const float* __restrict__ src, Dsts dst, int np, int nrows, int ZS, int C)
{
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (long long)gridDim.x * blockDim.x) {
        float4 v[8];
        for (int c = 0; c < C; ++c) v[c] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)row * ZS) + c);
        for (int r = 0; r < np; ++r)
            for (int c = 0; c < C; ++c) reinterpret_cast<float4*>(dst.p[r] + (size_t)row * ZS)[c] = v[c];
    }
}

// DRAM hog: random 128-byte line reads (the step kernel's gather), runs for `iters` rounds
__global__ void hog(const float4* __restrict__ buf, unsigned long long lines, int iters, float* sink)
{
    unsigned long long s = 12345;   // same sequence in the 8 lanes of a group: they read the 8 pieces of one line
    float acc = 0.f;
    const int piece = threadIdx.x & 7, grp = threadIdx.x >> 3;
    for (int i = 0; i < iters; ++i) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {           // four independent random lines in flight per 8-lane group
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            unsigned long long h = (s >> 3) ^ (unsigned long long)(grp + 64 * blockIdx.x) * 0xD1B54A32D192ED03ull;
            h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
            v[u] = __ldcg(buf + (h % lines) * 8 + piece);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].w;
    }
    if (acc == 1.2345f) *sink = acc;
}

struct Side {
    int dev, peer;
    float* src;
    Dsts dst;
    float4* hogbuf;
    float* sink;
    cudaStream_t s, hs;
    cudaEvent_t e0, e1;
};

int main(int argc, char** argv)
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    const int nrows = 131072, ZS = 32, P = 7;
    const unsigned long long hog_lines = (1ull << 30) / 128;    // 1 GB
    Side sd[2];
    for (int i = 0; i < 2; ++i) {
        sd[i].dev = i; sd[i].peer = 1 - i;
        CK(cudaSetDevice(i));
        int can = 0; CK(cudaDeviceCanAccessPeer(&can, i, 1 - i));
        if (!can) { printf("no peer access %d->%d\n", i, 1 - i); return 0; }
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        CK(cudaMalloc(&sd[i].src, (size_t)nrows * ZS * 4));
        CK(cudaMemset(sd[i].src, 1, (size_t)nrows * ZS * 4));
        CK(cudaMalloc(&sd[i].hogbuf, hog_lines * 128));
        CK(cudaMemset(sd[i].hogbuf, 0, hog_lines * 128));
        CK(cudaMalloc(&sd[i].sink, 4));
        CK(cudaStreamCreateWithFlags(&sd[i].s, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&sd[i].hs, cudaStreamNonBlocking));
        CK(cudaEventCreate(&sd[i].e0)); CK(cudaEventCreate(&sd[i].e1));
    }
    for (int i = 0; i < 2; ++i) {          // destinations of side i live on the peer
        CK(cudaSetDevice(1 - i));
        for (int r = 0; r < P; ++r) { CK(cudaMalloc(&sd[i].dst.p[r], (size_t)nrows * ZS * 4)); CK(cudaMemset(sd[i].dst.p[r], 0, (size_t)nrows * ZS * 4)); }
    }
    for (int i = 0; i < 2; ++i) { CK(cudaSetDevice(i)); CK(cudaDeviceSynchronize()); }

    const int reps = 10;
    // mode: 0 = device 0 pushes alone; 1 = both push (bidirectional); 2 = both push while both run the DRAM hog
    auto run = [&](const char* name, int mode, int np, double payload_bytes_per_dst, auto launch) {
        float best[2] = {1e30f, 1e30f};
        for (int rep = 0; rep < reps; ++rep) {
            const int nside = mode == 0 ? 1 : 2;
            if (mode == 2)
                for (int i = 0; i < 2; ++i) { CK(cudaSetDevice(i)); hog<<<148 * 2, 512, 0, sd[i].hs>>>(sd[i].hogbuf, hog_lines, 400, sd[i].sink); }
            for (int i = 0; i < nside; ++i) {
                CK(cudaSetDevice(i));
                CK(cudaEventRecord(sd[i].e0, sd[i].s));
                launch(sd[i]);
                CK(cudaGetLastError());
                CK(cudaEventRecord(sd[i].e1, sd[i].s));
            }
            for (int i = 0; i < 2; ++i) { CK(cudaSetDevice(i)); CK(cudaDeviceSynchronize()); }
            for (int i = 0; i < nside; ++i) { float ms; CK(cudaEventElapsedTime(&ms, sd[i].e0, sd[i].e1)); if (ms < best[i]) best[i] = ms; }
        }
        const double gb = payload_bytes_per_dst * np / 1e9;
        printf("%-58s mode=%d np=%d : %8.1f us  %7.1f GB/s payload", name, mode, np, 1e3 * best[0], gb / (best[0] * 1e-3));
        if (mode) printf("   | reverse dir %8.1f us %7.1f GB/s", 1e3 * best[1], gb / (best[1] * 1e-3));
        printf("\n");
        fflush(stdout);
    };

    // how long does the hog alone take (to see that it covers the push)?
    {
        CK(cudaSetDevice(0));
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        hog<<<148 * 2, 512, 0, sd[0].hs>>>(sd[0].hogbuf, hog_lines, 400, sd[0].sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, sd[0].hs));
        hog<<<148 * 2, 512, 0, sd[0].hs>>>(sd[0].hogbuf, hog_lines, 400, sd[0].sink);
        CK(cudaEventRecord(b, sd[0].hs));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        const double lines = 148.0 * 2 * 64 * 400 * 4;
        printf("hog alone: %.1f us, %.1f G lines/s = %.0f GB/s of 128-byte fills\n", 1e3 * ms, lines / (ms * 1e-3) / 1e9, lines * 128 / (ms * 1e-3) / 1e9);
    }

    const int grids[] = {16, 32, 64, 148, 296};
    for (int mode = 0; mode < 3; ++mode) {
        for (int np : {1, 7}) {
            for (int g : grids) {
                char nm[128];
                snprintf(nm, sizeof nm, "coop 80B in 128B slots, grid %3d x 256, U=8", g);
                run(nm, mode, np, nrows * 80.0, [&](Side& s) { push_coop<8><<<g, 256, 0, s.s>>>(s.src, s.dst, np, nrows, ZS, ZS, 5); });
            }
            for (int g : {32, 148}) {
                char nm[128];
                snprintf(nm, sizeof nm, "coop 80B in 128B slots, grid %3d x 512, U=4", g);
                run(nm, mode, np, nrows * 80.0, [&](Side& s) { push_coop<4><<<g, 512, 0, s.s>>>(s.src, s.dst, np, nrows, ZS, ZS, 5); });
                snprintf(nm, sizeof nm, "coop 96B in 128B slots, grid %3d x 256, U=8", g);
                run(nm, mode, np, nrows * 96.0, [&](Side& s) { push_coop<8><<<g, 256, 0, s.s>>>(s.src, s.dst, np, nrows, ZS, ZS, 6); });
                snprintf(nm, sizeof nm, "coop full 128B slots, grid %3d x 256, U=8", g);
                run(nm, mode, np, nrows * 128.0, [&](Side& s) { push_coop<8><<<g, 256, 0, s.s>>>(s.src, s.dst, np, nrows, ZS, ZS, 8); });
                snprintf(nm, sizeof nm, "coop 80B rows -> dense destination, grid %3d x 256, U=8", g);
                run(nm, mode, np, nrows * 80.0, [&](Side& s) { push_coop<8><<<g, 256, 0, s.s>>>(s.src, s.dst, np, nrows, ZS, 20, 5); });
                snprintf(nm, sizeof nm, "lane-per-row 5 x 16B at 128B stride, grid %3d x 256", g);
                run(nm, mode, np, nrows * 80.0, [&](Side& s) { push_lane_per_row<<<g, 256, 0, s.s>>>(s.src, s.dst, np, nrows, ZS, 5); });
            }
            // copy engine for reference: np copies of the padded slice
            run("cudaMemcpyPeerAsync of the padded slice (128B slots)", mode, np, nrows * 128.0, [&](Side& s) {
                for (int r = 0; r < np; ++r) cudaMemcpyPeerAsync(s.dst.p[r], s.peer, s.src, s.dev, (size_t)nrows * ZS * 4, s.s);
            });
        }
    }
    return 0;
}
