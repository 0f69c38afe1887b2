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

// per-trajectory (or per-sample) generator: normals by Box-Muller, Gamma by Marsaglia-Tsang, Gaussian / Student draws
struct Rng {
    Philox ph;
    uint32_t t_lo, t_hi;
    // two standard normals for (step, stream, call) by Box-Muller
    SSM_DEV void normal2(uint32_t step, uint32_t stream, uint32_t call, double &z0, double &z1) const {
        uint32_t r[4];
        ph.gen(t_lo, t_hi, step, (stream << 16) | call, r);
        const double u = u01(r[0], r[1]), v = u01(r[2], r[3]);
        const double rad = sqrt(-2.0 * log(u));
        double s, c;
        sincospi(2.0 * v, &s, &c);
        z0 = rad * c;
        z1 = rad * s;
    }
    template <int DIM>
    SSM_DEV void normals(uint32_t step, uint32_t stream, double (&z)[DIM]) const {
#pragma unroll
        for (int i = 0; i < DIM; i += 2) {
            double a, b;
            normal2(step, stream, i / 2, a, b);
            z[i] = a;
            if (i + 1 < DIM) z[i + 1] = b;
        }
    }
    // Gamma(shape a >= 1, scale 1) by Marsaglia-Tsang; calls >= 64 are reserved for it
    SSM_DEV double gamma(uint32_t step, uint32_t stream, double a) const {
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (uint32_t it = 0; it < 64; ++it) {
            uint32_t r[4];
            double x, unused;
            normal2(step, stream, 64 + 2 * it, x, unused);
            ph.gen(t_lo, t_hi, step, (stream << 16) | (65 + 2 * it), r);
            const double u = u01(r[0], r[1]);
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
        }
        return d;
    }
    // DIM-variate draw  F z  (Gaussian) or  F z / sqrt(g), g ~ Gamma(nu/2, 2/nu)  (Student, utils.py:380-382)
    template <int DIM>
    SSM_DEV void draw(uint32_t step, uint32_t stream, const double *F, double dof, double (&o)[DIM]) const {
        double z[DIM];
        normals<DIM>(step, stream, z);
        double sc = 1.0;
        if (dof > 0.0) sc = rsqrt(gamma(step, stream, 0.5 * dof) * (2.0 / dof));
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; ++j) s = fma(F[i * DIM + j], z[j], s);
            o[i] = s * sc;
        }
    }
};

}  // namespace ssm
