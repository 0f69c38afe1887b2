// Counter-based random numbers shared by the simulation kernels (ssm_sim.cu) and the bootstrap kernel
// (ssm_bootstrap.cu): draws depend on (key, counter) only, never on the grid shape.
#pragma once
#include "ssm_common.cuh"

namespace ssm {

// ---- Philox4x32-10 (Salmon et al., SC'11) -----------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    SSM_DEV static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    SSM_DEV void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
        uint32_t c[4] = {c0, c1, c2, c3};
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(c, a, b);
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = c[i];
    }
};

// uniform in (0, 1) from 64 random bits (53-bit mantissa, never 0)
SSM_DEV double u01(uint32_t lo, uint32_t hi) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace ssm
