// mcl_rng.cuh -- Philox4x32-10 and the SFU helpers shared by the native kernels (sm_100a).
#pragma once
#include <stdint.h>

namespace mcl {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
// counter domains (top 4 bits of counter word 3)
constexpr uint32_t DOM_STEP = 0u, DOM_SCALAR = 1u, DOM_SEED_E = 2u, DOM_SEED_H = 3u, DOM_STEP1 = 4u, DOM_SEL = 5u;

struct RoundKeys { uint32_t k[20]; };

__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, const RoundKeys &K)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)PHILOX_M0 * c0;
        unsigned long long p1 = (unsigned long long)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
}

__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float warp_min_f32(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v)
{
    uint32_t r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
// u32 -> uniform on [2^-24, 1 - 2^-24] (exact): one LEA.HI + one FADD, no conversion instruction
__device__ __forceinline__ float u01(uint32_t r) { return __uint_as_float((r >> 9) | 0x3f800000u) - 0.99999994f; }

// splitmix-style spreading of the user seed into the two Philox key words, then the ten round keys
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline RoundKeys make_round_keys(uint64_t seed)
{
    RoundKeys K;
    uint64_t s = mix64(seed);
    uint32_t k0 = (uint32_t)s, k1 = (uint32_t)(s >> 32);
    for (int r = 0; r < 10; r++) { K.k[2 * r] = k0; K.k[2 * r + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
    return K;
}

}  // namespace mcl
