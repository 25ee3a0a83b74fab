// Counter-based Philox4x32-10 streams (Salmon et al. 2011, "Parallel random numbers: as easy as 1, 2, 3").
// Replaces the reference's per-thread MT19937 (Random.cpp:41-126): every photon packet owns the stream
// (key = simulation seed, counter = (global packet index, draw block)), so results do not depend on how
// packets are mapped to threads, CTAs or GPUs.
#pragma once
#include <cstdint>

namespace skg
{

struct Philox
{
    uint32_t key0, key1;
    uint32_t c0, c1, c2, c3;    // c0,c1 = packet index; c2 = block counter; c3 = stream kind
    uint32_t out[4];
    int have;

    __device__ __forceinline__ void init(uint64_t seed, uint64_t packet, uint32_t kind = 0)
    {
        key0 = (uint32_t)seed; key1 = (uint32_t)(seed >> 32);
        c0 = (uint32_t)packet; c1 = (uint32_t)(packet >> 32); c2 = 0; c3 = kind;
        have = 0;
    }

    __device__ __forceinline__ void round(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t k0, uint32_t k1)
    {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, a), lo0 = M0 * a;
        uint32_t hi1 = __umulhi(M1, c), lo1 = M1 * c;
        a = hi1 ^ b ^ k0; b = lo1; c = hi0 ^ d ^ k1; d = lo0;
    }

    __device__ __forceinline__ void refill()
    {
        uint32_t a = c0, b = c1, c = c2, d = c3, k0 = key0, k1 = key1;
#pragma unroll
        for (int i = 0; i < 10; i++)
        {
            round(a, b, c, d, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = a; out[1] = b; out[2] = c; out[3] = d;
        c2++;
        have = 2;
    }

    // uniform deviate in the OPEN interval (0,1), 53 random bits (Random::uniform excludes 0 and 1 too,
    // Random.cpp:89-126)
    __device__ __forceinline__ double uniform()
    {
        if (have == 0) refill();
        have--;
        uint64_t bits = ((uint64_t)out[2 * have + 1] << 32) | out[2 * have];
        return ((double)(bits >> 12) + 0.5) * (1.0 / 4503599627370496.0);     // 2^-52
    }
};

}   // namespace skg
