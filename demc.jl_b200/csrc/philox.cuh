// philox.cuh — Philox4x32-10 counter-based RNG and the draw mappings of the DEMCz proposal.
// Stream layout (shared bit-for-bit with oracle/demcz_oracle.c):
//   key = (seed_lo, seed_hi);  ctr = (global chain id, gen lo32, (block << 16) | call, gen hi32)
//   call 0     -> (w_r1, w_r2, w_accept, spare)
//   call 1 + q -> jitter normals of columns 4q..4q+3 (two Box-Muller pairs)
// Replaces the global MersenneTwister draws of src/demcz.jl:177,179,184,186,198 (reference).
#pragma once
#include <stdint.h>

namespace demcz {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct u32x4 { uint32_t x, y, z, w; };

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += kPhiloxW0;   // uniform across the warp: folded into constants by the compiler
        k1 += kPhiloxW1;
    }
    return u32x4{c0, c1, c2, c3};
}

// ((w >> 9) + 0.5) * 2^-23 : in (0,1), exact in FP32 (23-bit integer + 1/2)
__device__ __forceinline__ float uniform01(uint32_t w)
{
    return fmaf((float)(w >> 9), 1.1920928955078125e-07f, 5.9604644775390625e-08f);
}

// i1 = rand(1:M); deleteat!(set,i1); i2 = rand(set)  (src/demcz.jl:176-179), 0-based
__device__ __forceinline__ void draw_pair(uint32_t w1, uint32_t w2, uint32_t M, uint32_t& r1, uint32_t& r2)
{
    r1 = __umulhi(w1, M);
    const uint32_t j = __umulhi(w2, M - 1u);
    r2 = j + (j >= r1 ? 1u : 0u);
}

// z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1).  The angle is folded to
// (-pi, pi) (cos/sin(t + pi) = -cos/-sin t) where the MUFU approximations are most accurate.
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1)
{
    const float u1 = uniform01(wa), u2 = uniform01(wb);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.2831853071795864769f * (u2 - 0.5f), &s, &c);
    z0 = -r * c;
    z1 = -r * s;
}

}  // namespace demcz
