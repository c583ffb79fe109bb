// philox.cuh — Philox4x32-10 counter-based RNG and the draw mappings of the DEMCz proposal.
// Stream layout (shared bit-for-bit with oracle/demcz_oracle.c):
//   key = (seed_lo, seed_hi);  ctr = (global chain id, gen lo32, (block << 16) | call, gen hi32)
//   call 0     -> (w_r1, w_r2, w_accept, spare)
//   call 1 + q -> jitter normals of columns 8q..8q+7: word j -> one Box-Muller pair (16-bit
//                 uniforms from the high / low half-word) -> columns 8q+2j, 8q+2j+1
// Replaces the global MersenneTwister draws of src/demcz.jl:177,179,184,186,198 (reference).
#pragma once
#include <stdint.h>

namespace demcz {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct u32x4 { uint32_t x, y, z, w; };

// The key schedule (key + r * Weyl constant) depends only on the seed: the host expands it once and
// the kernels take the ten round keys as constant-bank operands of the XORs.
struct PhiloxKeys { uint32_t k0[10], k1[10]; };

inline PhiloxKeys make_philox_keys(uint64_t seed)
{
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { k.k0[r] = a; k.k1[r] = b; a += kPhiloxW0; b += kPhiloxW1; }
    return k;
}

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        c0 = hi1 ^ c1 ^ K.k0[r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ K.k1[r];
        c3 = lo0;
    }
    return u32x4{c0, c1, c2, c3};
}

// host copy (generation-level decisions such as "is this a snooker generation" are taken by the host from the same stream)
inline u32x4 philox4x32_10_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K)
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0, p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r], n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    return u32x4{c0, c1, c2, c3};
}

// Snooker update (ter Braak & Vrugt 2008), stream layout: block field 0xFFFF.
//   generation-level choice: ctr = (0xFFFFFFFF, gen lo, 0xFFFF0000, gen hi), word 0 -> u; snooker iff u < snooker_prob
//   per chain: call 0xFFFF0000 -> (w_z, w_z1, w_z2, w_accept), call 0xFFFF0001 -> word 0: jump size gamma_s = 1.2 + u
// z, z1, z2 are three DISTINCT archive rows: z = mulhi(w_z, M); z1 = j + (j >= z), j = mulhi(w_z1, M-1);
// z2 = j2 stepped over min(z,z1) and max(z,z1), j2 = mulhi(w_z2, M-2).
constexpr uint32_t kSnookerCall = 0xFFFF0000u;
__host__ __device__ __forceinline__ void draw_triple(uint32_t z_hi32, uint32_t z1_hi32, uint32_t z2_hi32, uint32_t& z, uint32_t& z1, uint32_t& z2)
{
    // arguments are already the mulhi results: z in [0,M), z1 in [0,M-1), z2 in [0,M-2)
    z = z_hi32;
    z1 = z1_hi32 + (z1_hi32 >= z ? 1u : 0u);
    const uint32_t lo = z < z1 ? z : z1, hi = z < z1 ? z1 : z;
    z2 = z2_hi32;
    if (z2 >= lo) ++z2;
    if (z2 >= hi) ++z2;
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

// One Philox word -> two N(0,1) draws (the eps jitter of src/demcz.jl:184/186):
//   u1 = ((w >> 16) + 0.5) 2^-16, u2 = ((w & 0xffff) + 0.5) 2^-16   (exact in FP32, in (0,1))
//   z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1)
// The angle is folded to (-pi, pi) (cos/sin(t + pi) = -cos/-sin t) where the MUFU approximations
// are most accurate.
__device__ __forceinline__ float2 box_muller16(uint32_t w)
{
    const float u1 = fmaf((float)(w >> 16), 1.52587890625e-05f, 7.62939453125e-06f);
    const float t = fmaf((float)(w & 0xFFFFu), 1.52587890625e-05f, 7.62939453125e-06f - 0.5f);   // u2 - 1/2
    float l2, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));               // sqrt(-2 ln u1)
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(6.2831853071795864769f * t));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(6.2831853071795864769f * t));
    return make_float2(-r * c, -r * s);
}

}  // namespace demcz
