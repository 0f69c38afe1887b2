// Shared device helpers: packed-symmetric small matrices held in registers, Cholesky and
// triangular solves, streaming loads/stores of the [component][step][trajectory] layout.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ssm_b200.h"
#include "ssm_math.cuh"

#define SSM_DEV __device__ __forceinline__

namespace ssm {

// index into a packed lower-triangular array, r >= c
SSM_DEV constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }
SSM_DEV constexpr int sym(int r, int c) { return r >= c ? tri(r, c) : tri(c, r); }
template <int D>
struct TriSize {
    static constexpr int value = D * (D + 1) / 2;
};

// streaming (evict-first) global accesses: every bulk array is touched exactly once per pass
SSM_DEV double ld_stream(const double *p) { return __ldcs(p); }
#ifdef SSM_ST_PLAIN
SSM_DEV void st_stream(double *p, double v) { *p = v; }
#else
SSM_DEV void st_stream(double *p, double v) { __stcs(p, v); }
#endif

// Row pointer of a bulk array: base + (k * ld + t), computed once per array and step and made opaque to the
// optimiser, which otherwise re-associates base + (rk + c * cs) into a 64-bit add plus a 64-bit scaled add (LEA
// pair) per access.  With the row pointer pinned every access is `q + c * cs` = one add with a uniform operand.
template <class T>
SSM_DEV T *row_ptr(T *base, long long rk) {
    T *q = base + rk;
    asm volatile("" : "+l"(q));
    return q;
}

// The component stride as a 32-bit unsigned value (models with dx > 1: n_steps * ld < 2^32 is checked at launch, and
// 2^32 doubles per component are 34 GB): the byte offset c * cs * 8 is ONE IMAD.WIDE.U32 with an immediate, against
// IMAD.WIDE.U32 + IMAD + IADD for a 64-bit stride -- 87 addresses per reentry step.
#ifndef SSM_NARROW_STRIDE
#define SSM_NARROW_STRIDE 1
#endif
template <bool NARROW>
struct CompStride {
    long long v;
    SSM_DEV explicit CompStride(long long s) : v(s) {}
    SSM_DEV long long operator()(int c) const { return c * v; }
};
template <>
struct CompStride<true> {
    unsigned v;
    SSM_DEV explicit CompStride(long long s) : v((unsigned)s) {}
    SSM_DEV size_t operator()(int c) const { return (size_t)(unsigned)c * v; }
};
template <int DX>
struct NarrowStride {
    static constexpr bool value = SSM_NARROW_STRIDE != 0 && DX > 1;
};
// launch-time check of the 32-bit component stride
template <int DX>
inline bool stride_fits(long long n_steps, long long ld) { return !NarrowStride<DX>::value || n_steps * ld < (1LL << 32); }

// Stream-ordered scratch memory (scheduler workspace, partial statistics rows).  The default memory pool returns
// freed blocks to the driver at the next synchronisation (release threshold 0), so a caller that synchronises between
// calls pays a driver allocation of tens of MB in front of every launch (measured as +-20 % run-to-run noise of the
// forward pass).  The first scratch allocation on a device raises the threshold: freed scratch stays cached.
inline cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured[dev] = true;
    }
    return cudaMallocAsync(p, bytes, s);
}

SSM_DEV double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// Out-of-line fp64 math.  The fused forward pass for the 5-D models is one straight-line body; with
// libm inlined at every use (16 exp, ~20 sqrt, ~25 divisions, 5 atan2 per step) it is ~144 KB of SASS and
// the SMs run instruction-fetch bound at ~1.9 IPC (profiles/: sm__icc_request_hit_rate 58-75 %,
// stall_no_instruction dominant at saturation).  Calling one shared copy of each routine keeps the
// body inside the instruction cache; scalar arguments and results stay in registers, and the results
// are bit-identical to the inlined calls.
#ifndef SSM_INLINE_MATH
#define SSM_MATH_FN static __device__ __noinline__
#else
#define SSM_MATH_FN static __device__ __forceinline__
#endif
// exp stays libm's.  Measured alternatives on the reentry forward pass (16 calls per step, 64 instructions each of which
// 22 are UMOV halves of immediates; 125 000 x 500, filter only / with predictive moments, libm 15.75 / 20.28 ms):
// own 2^k exp(r) with the coefficients in a constant-bank table (37 instructions per call, -430 per step) 16.14 / 20.45;
// own Estrin-scheme evaluation with immediates (5-deep instead of 14-deep DFMA chain) 15.92 / 20.31.  Neither the issue
// slots nor the chain latency of exp are what bounds the kernel; both were within 1 ulp of numpy.
// Round 2: sqrt / rsqrt (17-21 instructions each, 31 calls per reentry step) are inlined again while exp / atan2 /
// division stay out of line: a call costs ~10 marshalling instructions plus CALL / RET (28 % of the samples on these two
// routines were `branch_resolving`), which is more than half of such a small body, and 31 x 20 instructions do not hurt the
// instruction cache the way the 144 KB all-inline body did.  Measured on the reentry forward pass, 125 000 x 500:
// 15.64 -> 14.66 ms filter only, 20.11 -> 19.50 ms with predictive moments (SSM_INLINE_SMALL_MATH=0 restores the calls).
#ifndef SSM_INLINE_SMALL_MATH
#define SSM_INLINE_SMALL_MATH 1
#endif
#ifndef SSM_INLINE_EXP
#define SSM_INLINE_EXP 0
#endif
#if SSM_INLINE_SMALL_MATH
#define SSM_SMALL_MATH_FN static __device__ __forceinline__
#else
#define SSM_SMALL_MATH_FN SSM_MATH_FN
#endif
// Round 2: exp and atan2 are the lean routines of ssm_math.cuh (SSM_LEAN_MATH: 1 = inline, 2 = one out-of-line copy each,
// 0 = libm's as before).
#ifndef SSM_LEAN_MATH
#define SSM_LEAN_MATH 0
#endif
#if SSM_LEAN_MATH == 1
static __device__ __forceinline__ double m_exp(double x) { return lean_exp(x); }
#elif SSM_LEAN_MATH == 2
static __device__ __noinline__ double m_exp(double x) { return lean_exp(x); }
#elif SSM_INLINE_EXP
static __device__ __forceinline__ double m_exp(double x) { return exp(x); }
#else
SSM_MATH_FN double m_exp(double x) { return exp(x); }
#endif
// Developer switch SSM_DUP_{EXP,ATAN2,SQRT,RSQRT}: evaluate the routine a second time on a perturbed argument and fold
// the result in with weight 0.0 (not removable: 0 * NaN) -- the time added is the marginal cost of that routine's calls.
#if defined(SSM_DUP_EXP) || defined(SSM_DUP_ATAN2) || defined(SSM_DUP_SQRT) || defined(SSM_DUP_RSQRT)
#define SSM_DUP(fn, r, ...) ((r) + 0.0 * fn(__VA_ARGS__))
#endif
SSM_SMALL_MATH_FN double m_sqrt(double x) { return sqrt(x); }
SSM_MATH_FN double m_rcp(double x) { return 1.0 / x; }
SSM_SMALL_MATH_FN double m_rsqrt(double x) { return rsqrt(x); }
SSM_MATH_FN double m_div(double a, double b) { return a / b; }
// Two results per call (16-byte struct: returned in registers).  sincos: 84 instructions inline, 28 of them UMOV halves of
// immediates, 11 copies per coordinated-turn step = 15 % of that loop body -- one shared copy instead.  div2: a / b and
// c / b by the same denominator share the reciprocal seed and its Newton steps; both quotients are the IEEE ones.
struct Pair2 {
    double u, v;
};
SSM_MATH_FN Pair2 m_sincos(double x) {
    Pair2 r;
    sincos(x, &r.u, &r.v);
    return r;
}
SSM_MATH_FN Pair2 m_div2(double a, double c, double b) {
    Pair2 r;
    r.u = a / b;
    r.v = c / b;
    return r;
}
#if SSM_LEAN_MATH == 1
static __device__ __forceinline__ double m_atan2(double y, double x) { return lean_atan2(y, x); }
#elif SSM_LEAN_MATH == 2
static __device__ __noinline__ double m_atan2(double y, double x) { return lean_atan2(y, x); }
#else
SSM_MATH_FN double m_atan2(double y, double x) { return atan2(y, x); }
#endif

// Lower Cholesky factor of a symmetric matrix given by its packed lower triangle.
// Mirrors dpotrf('L') as called by numpy.linalg.cholesky (mtran.py:139, bqmtran.py:98): only the
// lower triangle is read and a pivot <= 0 is a failure (LinAlgError).  A NaN pivot is NOT a failure:
// the OpenBLAS potrf behind numpy 2.3 tests `ajj <= 0` only, so NaNs propagate silently until
// scipy's check_finite in cho_factor raises ValueError (measured on the golden case
// c3_reentry_gpq_fail); the kernels reproduce that sequence.  Returns false on failure.
template <int D>
SSM_DEV bool chol_lower(const double (&A)[TriSize<D>::value], double (&L)[TriSize<D>::value], double *inv_diag = nullptr) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = A[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[tri(j, k)], L[tri(j, k)], s);
        ok = ok && !(s <= 0.0);
        // one out-of-line call per pivot: r = 1/sqrt(s) (<= 1 ulp), L_jj = s r, 1/L_jj = r.  LAPACK computes
        // sqrt and divides; the difference is ~1 ulp of L, far below the 1e-9 parity tolerance.
        const double inv = m_rsqrt(s);
        const double d = s * inv;
        L[tri(j, j)] = d;
        if (inv_diag) inv_diag[j] = inv;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double t = A[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[tri(i, k)], L[tri(j, k)], t);
            L[tri(i, j)] = t * inv;
        }
    }
    return ok;
}

// X = (S^-1 C)^T for SPD S (packed lower, E x E) and C (E x D): the reference's
// cho_solve(cho_factor(S), C).T (ssinf.py:321, 342).  K is D x E.  Returns false if S is not PD.
template <int E, int D>
SSM_DEV bool spd_gain(const double (&S)[TriSize<E>::value], const double (&C)[E][D], double (&K)[D][E],
                      double (&Ls)[TriSize<E>::value]) {
    double inv[E];
    const bool ok = chol_lower<E>(S, Ls, inv);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        double z[E];
#pragma unroll
        for (int i = 0; i < E; ++i) {  // forward substitution L z = C[:, d]
            double t = C[i][d];
#pragma unroll
            for (int k = 0; k < i; ++k) t = fma(-Ls[tri(i, k)], z[k], t);
            z[i] = t * inv[i];
        }
#pragma unroll
        for (int i = E - 1; i >= 0; --i) {  // back substitution L^T x = z
            double t = z[i];
#pragma unroll
            for (int k = i + 1; k < E; ++k) t = fma(-Ls[tri(k, i)], K[d][k], t);
            K[d][i] = t * inv[i];
        }
    }
    return ok;
}

SSM_DEV bool finite_d(double v) { return isfinite(v); }

}  // namespace ssm
