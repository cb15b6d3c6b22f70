// Philox4x32-10 counter-based generator (Salmon et al., SC'11), the engine's native stream (kernel K1).
// Replaces the reference's global numpy MT19937 stream (esloco.py:100) -- statistically, not bit for bit;
// the replay entry points bypass it.  Counter layout used by the engine:
//   c0 = draw index n within the iteration, c1 = auxiliary index (blocked-hit group / target row),
//   c2 = iteration id, c3 = (combination << 4) | slot;  key = 64-bit seed.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ESL_HD __host__ __device__ __forceinline__
#else
#define ESL_HD inline
#endif

struct Philox4 {
    uint32_t x, y, z, w;
};

enum : uint32_t { ESL_SLOT_PLACE = 0u, ESL_SLOT_BLOCK = 1u, ESL_SLOT_SCALE = 2u };

ESL_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// Round keys of one Philox key, computed once per kernel so the ten key bumps are not redone for every draw.
struct PhiloxKeys {
    uint32_t a[10], b[10];
};
ESL_HD PhiloxKeys philox_keys(uint32_t k0, uint32_t k1)
{
    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) { k.a[r] = k0 + (uint32_t)r * 0x9E3779B9u; k.b[r] = k1 + (uint32_t)r * 0xBB67AE85u; }
    return k;
}
ESL_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.a[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.b[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return Philox4{c0, c1, c2, c3};
}

// 53-bit uniform in [0,1) from two words, the same construction numpy's legacy stream uses.
ESL_HD double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
