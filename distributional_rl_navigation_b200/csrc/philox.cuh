// Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), written
// from the published round function: the device-side random streams of the replay sampler and of the acting kernel
// (taus, epsilon-greedy).  Counter-based: draw (key, counter) is a pure function, so a launch is reproducible and needs
// no generator state in memory.  The reference draws these numbers from python `random` / torch's CPU generator
// (replay_buffer.py:47, agent.py:200-203, model.py:149); its streams are mirrored on the HOST by the single-env path,
// the vectorised device path uses this generator (parity harnesses inject indices / taus instead).
#pragma once
#include <stdint.h>

namespace philox {

struct u4 { uint32_t x, y, z, w; };

__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ inline u4 philox4x32_10(u4 ctr, uint32_t k0, uint32_t k1)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = u4{hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0};
        k0 += W0; k1 += W1;
    }
    return ctr;
}

// uniform in [0, 1): 24 random bits, like torch.rand for float32 (a multiple of 2^-24, never 1.0)
__host__ __device__ inline float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

}  // namespace philox
