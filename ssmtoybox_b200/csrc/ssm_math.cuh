// Lean fp64 exp and atan2 for the fused forward pass.
//
// The reference evaluates np.exp / np.arctan2 inside dyn_fcn / meas_fcn (ssmod.py:548-566 reentry drag, :1227-1252 radar
// bearing); the forward pass of the 5-D models calls them 16 + 5 times per trajectory-step.  CUDA's libm versions cost
// 59 / 141 issued instructions per call including the call sequence -- 22 / 39 of them UMOV halves of 64-bit immediates --
// and the marginal cost of a call was measured at exactly its share of the issued instructions (tools: SSM_DUP_EXP /
// SSM_DUP_ATAN2: 16 exp = 2.3 ms, 5 atan2 = 1.6 ms of a 14.1 ms pass).  The versions below keep libm's algorithms
// (Cody-Waite reduction + degree-11 polynomial; one division + odd polynomial) but
//   * read their coefficients from a __constant__ table (one LDCU.128 per two coefficients),
//   * leave everything that is not a finite, normal-range argument to ONE rare branch into the out-of-line libm routine
//     (NaN, infinities, zeros, |x| >= 700: bit-identical special-case behaviour by construction),
//   * divide with a single-precision reciprocal seed + two Newton steps + one residual correction,
//   * reduce atan to |t| <= tan(pi/8) with the division folded in ((mn - mx) / (mn + mx)), so 13 coefficients suffice.
// 26 / ~60 instructions inline.  Accuracy (tests/test_gpu_parity.py::test_math_probe, against numpy): exp <= 1 ulp,
// atan2 <= 2 ulp -- the bounds libm's own routines are specified to.  Coefficients: tools/gen_math_coeffs.py (mpmath).
#pragma once
#include <math.h>

namespace ssm {

struct MathTab {
    double l2e, nln2_hi, nln2_lo, pad0;
    double ec[10];  // (exp(r) - 1 - r) / r^2, degree 9: polynomial error 0.14 ulp on |r| <= ln2 / 2
    double ac[13];  // (atan(sqrt s) / sqrt s - 1) / s, degree 12: 0.03 ulp on s <= tan(pi/8)^2
    double tan_pio8;
    double pio4_hi, pio4_lo, pio2_hi, pio2_lo, pi_hi, pi_lo;
};

#ifndef SSM_LEAN_TAB
#define SSM_LEAN_TAB 1  // 1: coefficients from the __constant__ table; 0: as immediates (two UMOV / MOV per coefficient)
#endif
#define SSM_MATH_TAB_INIT                                                                                                              \
    {1.44269504088896339e+00, -6.93147180559945286e-01, -2.31904681384629956e-17, 0.0,                                                 \
     {5.00000000000000111e-01, 1.66666666666666685e-01, 4.16666666666241636e-02, 8.33333333333006500e-03, 1.38888889171967186e-03,     \
      1.98412698630405450e-04, 2.48015213223686919e-05, 2.75572684803100238e-06, 2.76200758799834784e-07, 2.51003758325613201e-08},    \
     {-3.33333333333333315e-01, 1.99999999999999928e-01, -1.42857142857115177e-01, 1.11111111107547289e-01, -9.09090906700999041e-02,  \
      7.69230673600017345e-02, -6.66664202005516626e-02, 5.88192525319285939e-02, -5.25804155429774114e-02, 4.71939503002718613e-02,   \
      -4.10443626575408421e-02, 3.06357041129463811e-02, -1.39182292910010330e-02},                                                    \
     4.14213562373095034e-01, 7.85398163397448279e-01, 3.06161699786838302e-17, 1.57079632679489656e+00, 6.12323399573676604e-17,      \
     3.14159265358979312e+00, 1.22464679914735321e-16}
#if SSM_LEAN_TAB
static __constant__ MathTab ssm_math_tab = SSM_MATH_TAB_INIT;
#define SSM_MATH_TAB_REF const MathTab &T = ssm_math_tab
#else
#define SSM_MATH_TAB_REF constexpr MathTab T = SSM_MATH_TAB_INIT
#endif

// The slow paths take the lean value as an (unused, but opaquely consumed) argument: without that use the optimiser sinks
// the whole lean computation into the not-taken side of the patch branch, where it no longer dominates the next call
// and nothing is merged.
static __device__ __noinline__ double exp_libm(double x, double lean) {
    asm volatile("" ::"d"(lean));
    return exp(x);
}
static __device__ __noinline__ double atan2_libm(double y, double x, double lean) {
    asm volatile("" ::"d"(lean));
    return atan2(y, x);
}

__device__ __forceinline__ double lean_exp(double x) {
    // The lean value is computed unconditionally and patched afterwards, so that the straight-line part dominates every
    // later use: identical calls at structurally equal sigma points are merged by common-subexpression elimination
    // (5 instead of 11 bearings per step), which an early return into the slow path would prevent.
    SSM_MATH_TAB_REF;
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: the low word of x log2(e) + magic is round(x log2 e)
    double t = fma(x, T.l2e, magic);
    const int k = __double2loint(t);
    t -= magic;
    double r = fma(t, T.nln2_hi, x);  // exact (Cody-Waite)
    r = fma(t, T.nln2_lo, r);
    double p = T.ec[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) p = fma(p, r, T.ec[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));  // |k| <= 1010: the result stays normal
    if (!(fabs(x) < 700.0)) p = exp_libm(x, p);  // overflow, gradual underflow, NaN: libm
    return p;
}

__device__ __forceinline__ double lean_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const double sum = ax + ay;
    SSM_MATH_TAB_REF;
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax, mn = swap ? ax : ay;
    // atan(mn / mx) = pi/4 + atan((mn - mx) / (mn + mx)) above tan(pi/8): one division either way, |t| <= tan(pi/8)
    const bool big = mn > T.tan_pio8 * mx;
    const double num = big ? mn - mx : mn, den = big ? mn + mx : mx;
    // single-precision seed (2^-22) + two Newton steps.  Not the fp64 approximation instruction: that needs inline PTX, and
    // the optimiser does not merge inline asm across the sigma points (11 instead of 5 copies of this routine per step).
    double rc = (double)__fdividef(1.0f, (float)den);
    double e = fma(-den, rc, 1.0);
    rc = fma(rc, e, rc);
    e = fma(-den, rc, 1.0);
    rc = fma(rc, e, rc);
    double t = num * rc;
    t = fma(fma(-den, t, num), rc, t);  // residual correction: t = num / den to 0.5 ulp
    const double s = t * t;
    double p = T.ac[12];
#pragma unroll
    for (int i = 11; i >= 0; --i) p = fma(p, s, T.ac[i]);
    double r = fma(t * s, p, t);
    if (big) r = T.pio4_hi + (r + T.pio4_lo);
    if (swap) r = T.pio2_hi - (r - T.pio2_lo);
    if (x < 0.0) r = T.pi_hi - (r - T.pi_lo);
    r = copysign(r, y);
    if (!(sum > 1e-30 && sum < 1e30)) r = atan2_libm(y, x, r);  // both zero, infinities, NaN, magnitudes outside the seed's range: libm
    return r;
}

}  // namespace ssm
