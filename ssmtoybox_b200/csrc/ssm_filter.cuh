// K2: fused forward pass.  One thread owns one trajectory for the whole time loop; the filter
// state (mean, packed covariance), the Cholesky factor, the sigma points and the function
// evaluations live in registers; quadrature weights are read from the kernel-parameter constant
// bank (fast path): LDCU into uniform registers, which the DFMAs take as their weight operand (sm_100
// has no constant-bank operand form for fp64); measurements are read and moments written with
// coalesced streaming accesses over the trajectory axis.
//
// Replaces, per trajectory and time step (file:line in /root/reference/ssmtoybox):
//   StateSpaceInference.forward_pass            ssinf.py:66-118
//   GaussianInference._time_update / _measurement_update      ssinf.py:254-323
//   StudentianInference._time_update / _measurement_update    ssinf.py:634-736
//   SigmaPointTransform.apply                   mtran.py:105-149
//   BQTransform.apply (+ _mean/_covariance/_cross_covariance)  bq/bqmtran.py:60-223
//   StudentTProcessTransform._covariance        bq/bqmtran.py:394-415 (bq/bqmod.py:1132-1160)
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "ssm_models.cuh"
#include "ssm_scores.cuh"

namespace ssm {

enum { PTS_AXIS_C = 0, PTS_AXIS = 1, PTS_GENERIC = 2 };
// Internal transform kind of the fast path (never part of the C ABI): a BQ transform on a [0 | cI | -cI] point set whose
// weights are bitwise invariant under the coordinate reflections x_j -> -x_j (weights_reflective()).  See the BQR branch of
// moment_transform for what that buys.
#define SSM_TF_BQR 1001
#ifndef SSM_WEIGHT_VIEWS
#define SSM_WEIGHT_VIEWS 0
#endif
#ifndef SSM_CROSS_COLS_MIN_E
#define SSM_CROSS_COLS_MIN_E 99  // output dimension from which the BQ cross-covariance is formed column by column (measured slower)
#endif
#ifndef SSM_SYM_COLUMNS
#define SSM_SYM_COLUMNS 1
#endif
#ifndef SSM_SYM_WC
#define SSM_SYM_WC 1  // fast path: half-row form of fx Wc fx^T (symmetric Wc); 0 = dense rows as in round 1
#endif
constexpr int GEN_CAP = 64;  // capacity of the runtime-N (generic point set) path with the function values kept per thread
constexpr int GEN_CAP_STREAM = 4096;  // sigma-point rules beyond GEN_CAP points: two streaming passes, nothing stored

// ------------------------------------------------------------------------------------------------
// transform parameters, fast path: everything by value inside the kernel parameter block
// ------------------------------------------------------------------------------------------------
template <int D, int E, int NCAP, int KIND, int PTS>
struct TfConst {
    static constexpr int NW = (KIND == SSM_TF_SP) ? 1 : (KIND == SSM_TF_BQR ? D + 1 : NCAP);
    static constexpr int NK = (KIND == SSM_TF_TP) ? NCAP : 1;
    static constexpr int DU = (PTS == PTS_GENERIC) ? D : 1;
    static constexpr int DC = (KIND == SSM_TF_SP) ? 1 : D;
    int n;
    int tp_full;
    double c;  // axis point scale
    double tp_a, tp_b;  // nu - 2, 1 / (nu - 2 + N)
    double wm_[NCAP];
    double wc_[NCAP];
    double Wc_[NW][NCAP];
    double wch_[NW];  // 0.5 * Wc(j, j): diagonal of the half-row form (see SYMW)
    double Wcc_[DC][NCAP];
    double mv_[E][E];
    double iK_[NK][NCAP];
    double U_[DU][NCAP];
    int zero_[8];  // always 0, but only known at run time: see row_view
    // The weights as seen from output row a of an unrolled loop: the same table behind an offset the compiler cannot
    // fold.  Without it the loads of every weight W(i, j) are merged across the E unrolled rows: ~130 weights stay live
    // in uniform registers, get copied to vector registers (IMAD.U32 R, RZ, RZ, UR) and spilled; with it each use is one
    // LDCU.64 c[0x0][UR + imm] next to its DFMA.
    SSM_DEV const TfConst &row_view(int a) const {
#if SSM_WEIGHT_VIEWS
        return *(const TfConst *)((const char *)this + (size_t)((a + 1) & zero_[a & 7]) * 16);
#else
        return *this;
#endif
    }
    SSM_DEV double wm(int i) const { return wm_[i]; }
    SSM_DEV double wc(int i) const { return wc_[i]; }
    SSM_DEV double Wc(int i, int j) const { return Wc_[i][j]; }
    SSM_DEV double Wch(int j) const { return wch_[j]; }
    SSM_DEV double Wcc(int d, int i) const { return Wcc_[d][i]; }
    SSM_DEV double mv(int a, int b) const { return mv_[a][b]; }
    SSM_DEV double iK(int i, int j) const { return iK_[i][j]; }
    SSM_DEV double U(int d, int i) const { return U_[d][i]; }
    // The fast path is launched for bitwise SYMMETRIC covariance weights only (wc_symmetric(); every weight set the
    // reference produces is: bq/bqmod.py:519-521, 988-990 average Wc with its transpose), which lets the dense form
    // fx Wc fx^T run on the upper triangle of Wc: 605 instead of 770 DFMA for the 5-D dynamics transform.
    static constexpr bool SYMW = SSM_SYM_WC != 0;
};

// generic path: weights in global memory (uniform addresses -> one broadcast transaction per warp)
template <int D, int E>
struct TfGlobal {
    int n;
    int tp_full;
    double c;
    double tp_a, tp_b;
    const double *wm_, *wc_, *Wc_, *Wcc_, *iK_, *U_;
    double mv_[E][E];
    SSM_DEV const TfGlobal &row_view(int) const { return *this; }
    SSM_DEV double wm(int i) const { return __ldg(wm_ + i); }
    SSM_DEV double wc(int i) const { return __ldg(wc_ + i); }
    SSM_DEV double Wc(int i, int j) const { return __ldg(Wc_ + i * n + j); }
    SSM_DEV double Wch(int j) const { return 0.5 * __ldg(Wc_ + j * n + j); }
    static constexpr bool SYMW = false;  // any Wc, symmetric or not: dense rows
    SSM_DEV double Wcc(int d, int i) const { return __ldg(Wcc_ + d * n + i); }
    SSM_DEV double mv(int a, int b) const { return mv_[a][b]; }
    SSM_DEV double iK(int i, int j) const { return __ldg(iK_ + i * n + j); }
    SSM_DEV double U(int d, int i) const { return __ldg(U_ + d * n + i); }
};

// ------------------------------------------------------------------------------------------------
// sigma point i from mean m and Cholesky factor L:  x = m + L u_i   (mtran.py:139, bqmtran.py:101)
// Axis sets (UT / SR / fully-symmetric degree 3): u_i = +-c e_j, so x = m +- c L[:, j]; rows above j
// are structurally equal to m, which lets the compiler share model sub-expressions between points.
// The product and the sum are rounded separately, like numpy's  mean + L.dot(U).
// ------------------------------------------------------------------------------------------------
template <int D, int PTS, class Tf>
SSM_DEV void sigma_point(const Tf &tf, int i, const double (&m)[D], const double (&L)[TriSize<D>::value], double (&x)[D]) {
    if constexpr (PTS == PTS_GENERIC) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j <= r; ++j) s = fma(L[tri(r, j)], tf.U(j, i), s);
            x[r] = __dadd_rn(m[r], s);
        }
    } else {
        const int base = (PTS == PTS_AXIS_C) ? 1 : 0;
        if (PTS == PTS_AXIS_C && i == 0) {
#pragma unroll
            for (int r = 0; r < D; ++r) x[r] = m[r];
            return;
        }
        const int j = (i - base) % D;
        const double cs = ((i - base) < D) ? tf.c : -tf.c;
#pragma unroll
        for (int r = 0; r < D; ++r) x[r] = (r >= j) ? __dadd_rn(m[r], __dmul_rn(cs, L[tri(r, j)])) : m[r];
    }
}

// ------------------------------------------------------------------------------------------------
// one moment transform.  F(x, out) evaluates the (noise-free) model function.
// Outputs: mf (E), Cf packed lower (E), Cfx (E x D) when want_cross.  Returns false when the input
// covariance is not positive definite.
// ------------------------------------------------------------------------------------------------
template <int E, int NCAP, int SMT>
struct FxStore {  // shared memory, stride SMT doubles between elements of one thread
    double *p;
    SSM_DEV explicit FxStore(double *q) : p(q) {}
    SSM_DEV double get(int a, int i) const { return p[(a * NCAP + i) * SMT]; }
    SSM_DEV void set(int a, int i, double v) { p[(a * NCAP + i) * SMT] = v; }
};
template <int E, int NCAP>
struct FxStore<E, NCAP, 0> {  // registers
    double v_[E][NCAP];
    SSM_DEV explicit FxStore(double *) {}
    SSM_DEV double get(int a, int i) const { return v_[a][i]; }
    SSM_DEV void set(int a, int i, double v) { v_[a][i] = v; }
};

// SMT = 0: function evaluations fx (E x N) in registers.  SMT = CTA size: fx lives in shared memory, element
// (a, i) of this thread at sfx[(a * N + i) * SMT] (sfx already offset by threadIdx.x, conflict-free 8-byte
// accesses).  The arithmetic and its order are identical in both variants; the shared-memory variant frees
// ~2 E N registers per thread, which buys a higher occupancy for the 5-D models.
// The cross-covariance is handed out element by element through `sink(a, r, Cov(f_a, x_r))` so that the
// caller decides where it lives: the measurement transform keeps it in registers for the gain, the dynamics
// transform streams it straight to HBM (no 25-double live array between the transform and the stores).
// EXACT: mean and centred cross-covariance sums with separately rounded products (no FMA).  The reference's models
// with non-additive noise rely on EXACT cancellation between mirrored sigma points: z = 0.05 r x^2 at the +r / -r
// points must sum to a mean of exactly 0 and a state cross-covariance of exactly 0, because the measurement covariance
// is ~m^4 (1e-60 while the mean is still at rounding level) and K = Pxy / Py turns any 1e-17 residue of a fused
// multiply-add into O(1) garbage.  numpy's small dot products round every product before adding, so they cancel.
// Sigma-point transform over a LARGE generic point set (Gauss-Hermite product rules: 3^5 = 243 points for the 5-D
// models at the default degree, mtran.py:309-360): the function values do not fit a thread, so the rule is walked
// twice -- mean first, then the centred covariance and cross-covariance with the function re-evaluated at every
// point.  Loop nests and FMA order are those of the stored variant below, so the results are identical to it.
template <int D, int E, bool EXACT, class Tf, class F, class Sink>
SSM_DEV void sigma_point_transform_streamed(const Tf &tf, const int n, const double (&m)[D], const double (&L)[TriSize<D>::value], F f,
                                            double (&mf)[E], double (&Cf)[TriSize<E>::value], const bool want_cross, Sink sink) {
#pragma unroll
    for (int a = 0; a < E; ++a) mf[a] = 0.0;
    for (int i = 0; i < n; ++i) {
        double x[D], o[E];
        sigma_point<D, PTS_GENERIC>(tf, i, m, L, x);
        f(x, o);
        const double w = tf.wm(i);
#pragma unroll
        for (int a = 0; a < E; ++a) mf[a] = EXACT ? __dadd_rn(mf[a], __dmul_rn(o[a], w)) : fma(o[a], w, mf[a]);
    }
#pragma unroll
    for (int a = 0; a < TriSize<E>::value; ++a) Cf[a] = 0.0;
    double Cfx[E][D];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int r = 0; r < D; ++r) Cfx[a][r] = 0.0;
    for (int i = 0; i < n; ++i) {
        double x[D], o[E];
        sigma_point<D, PTS_GENERIC>(tf, i, m, L, x);
        f(x, o);
        const double w = tf.wc(i);
#pragma unroll
        for (int a = 0; a < E; ++a) o[a] -= mf[a];
#pragma unroll
        for (int a = 0; a < E; ++a) {
            const double t = o[a] * w;
#pragma unroll
            for (int b = 0; b <= a; ++b) Cf[tri(a, b)] = fma(t, o[b], Cf[tri(a, b)]);
        }
        if (want_cross) {
#pragma unroll
            for (int r = 0; r < D; ++r) {
                const double dxr = x[r] - m[r];
#pragma unroll
                for (int a = 0; a < E; ++a)
                    Cfx[a][r] = EXACT ? __dadd_rn(Cfx[a][r], __dmul_rn(__dmul_rn(o[a], w), dxr)) : fma(o[a] * w, dxr, Cfx[a][r]);
            }
        }
    }
    if (want_cross) {
#pragma unroll
        for (int a = 0; a < E; ++a)
#pragma unroll
            for (int r = 0; r < D; ++r) sink(a, r, Cfx[a][r]);
    }
}

template <int D, int E, int PTS, int NPTS, int KIND, int SMT, bool EXACT, class Tf, class F, class Sink>
SSM_DEV bool moment_transform(const Tf &tf, const double (&m)[D], const double (&P)[TriSize<D>::value], F f,
                              double (&mf)[E], double (&Cf)[TriSize<E>::value], const bool want_cross, Sink sink,
                              double *sfx) {
    constexpr int NCAP = (NPTS > 0) ? NPTS : GEN_CAP;
    const int n = (NPTS > 0) ? NPTS : tf.n;
    double L[TriSize<D>::value];
    const bool ok = chol_lower<D>(P, L);
    if constexpr (KIND == SSM_TF_SP && PTS == PTS_GENERIC) {
        if (n > NCAP) {
            sigma_point_transform_streamed<D, E, EXACT>(tf, n, m, L, f, mf, Cf, want_cross, sink);
            return ok;
        }
    }

    FxStore<E, NCAP, SMT> fxs(sfx);
#define fx(a, i) fxs.get(a, i)
#pragma unroll
    for (int i = 0; i < n; ++i) {
        double x[D], o[E];
        sigma_point<D, PTS>(tf, i, m, L, x);
        f(x, o);
#pragma unroll
        for (int a = 0; a < E; ++a) fxs.set(a, i, o[a]);
    }
    if constexpr (KIND == SSM_TF_BQR) {
        // Reflection-symmetric weights.  The point set [0 | c e_j | -c e_j] is mapped onto itself by each coordinate
        // reflection x_j -> -x_j, and so is everything the weights are made of (a kernel that depends on x_j - x'_j
        // through its square, a symmetric integration density): in exact arithmetic
        //   wm(j+) = wm(j-),   Wc = P Wc P^T for every reflection P,   Wcc(d, .) is odd under reflection d and even under
        //   all others, i.e. zero except for  Wcc(d, d+) = -Wcc(d, d-).
        // With  g = [f_0, f_1+ + f_1-, ..., f_D+ + f_D-]  and  a_j = f_j+ - f_j-  the three sums of bqmtran.py:175-223 are
        //   mean  = wm_0 g_0 + sum_j wm_j g_j                                              (D + 1 instead of 2 D + 1 terms)
        //   fx Wc fx^T = g S g^T + sum_j alpha_j a_j a_j^T,   S (D+1 x D+1), alpha_j = (Wc(j+,j+) - Wc(j+,j-)) / 2
        //   fx Wcc^T = [Wcc(d, d+) a_d]_d                                                   (E D instead of E N D terms)
        // -- for the 5-D reentry dynamics ~565 instead of ~1 040 multiply-adds per transform and 37 instead of 132
        // weights.  The host launches this instantiation only for weight sets that HAVE the invariance bit for bit
        // (weights_reflective(): the package's own weights, which are symmetrised when they are built; weights assigned
        // from a reference run carry its rounding noise in the entries that are zero here and take the dense path).
        static_assert(PTS == PTS_AXIS_C && NPTS == 2 * D + 1 && SMT == 0, "reflection-symmetric path: [0 | cI | -cI] points, registers");
        constexpr int M = D + 1;
        double g[E][M], av[E][D];
#pragma unroll
        for (int a = 0; a < E; ++a) {
            g[a][0] = fx(a, 0);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                g[a][1 + j] = fx(a, 1 + j) + fx(a, 1 + D + j);
                av[a][j] = fx(a, 1 + j) - fx(a, 1 + D + j);
            }
        }
#pragma unroll
        for (int a = 0; a < E; ++a) mf[a] = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const double w = tf.wm(i);
#pragma unroll
            for (int a = 0; a < E; ++a) mf[a] = fma(g[a][i], w, mf[a]);
        }
        if (want_cross) {
#pragma unroll
            for (int a = 0; a < E; ++a) {
                double T[D];
#pragma unroll
                for (int d = 0; d < D; ++d) T[d] = av[a][d] * tf.Wcc(d, 0);
#pragma unroll
                for (int r = 0; r < D; ++r) {  // (T L^T)[a][r] = sum_{d<=r} T[a][d] L[r][d]                bqmtran.py:223
                    double s = 0.0;
#pragma unroll
                    for (int d = 0; d <= r; ++d) s = fma(T[d], L[tri(r, d)], s);
                    sink(a, r, s);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < TriSize<E>::value; ++a) Cf[a] = 0.0;
        // g S g^T column by column on half rows (as the SYMW branch below, over D + 1 columns)
#pragma unroll
        for (int j = 0; j < M; ++j) {
            double hc[E];
            {
                const double w = tf.Wch(j);
#pragma unroll
                for (int a = 0; a < E; ++a) hc[a] = g[a][j] * w;
            }
#pragma unroll
            for (int i = 0; i < j; ++i) {
                const double w = tf.Wc(j, i);
#pragma unroll
                for (int a = 0; a < E; ++a) hc[a] = fma(g[a][i], w, hc[a]);
            }
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b < E; ++b) Cf[sym(a, b)] = fma(hc[a], g[b][j], Cf[sym(a, b)]);
        }
#pragma unroll
        for (int a = 0; a < E; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Cf[tri(a, b)] = fma(-mf[a], mf[b], (a == b) ? Cf[tri(a, b)] + Cf[tri(a, b)] : Cf[tri(a, b)]);
        // the odd parts are differences of neighbouring function values: small, and added after the cancellation
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double w = tf.wc(j);
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const double t = av[a][j] * w;
#pragma unroll
                for (int b = 0; b <= a; ++b) Cf[tri(a, b)] = fma(t, av[b][j], Cf[tri(a, b)]);
            }
        }
#pragma unroll
        for (int a = 0; a < E; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Cf[tri(a, b)] += tf.mv(a, b);
        return ok;
    } else {
    // mean_f = fx . wm                                          mtran.py:143, bqmtran.py:175
    // point-major: one weight, E consecutive uses (same sums, same order over i as a row-major loop)
#pragma unroll
    for (int a = 0; a < E; ++a) mf[a] = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
        const double w = tf.wm(i);
#pragma unroll
        for (int a = 0; a < E; ++a) mf[a] = EXACT ? __dadd_rn(mf[a], __dmul_rn(fx(a, i), w)) : fma(fx(a, i), w, mf[a]);
    }
#pragma unroll
    for (int a = 0; a < TriSize<E>::value; ++a) Cf[a] = 0.0;

    if (KIND == SSM_TF_SP) {
        // centred form with diagonal weights                    mtran.py:145-148
#pragma unroll
        for (int a = 0; a < E; ++a)
#pragma unroll
            for (int i = 0; i < n; ++i) fxs.set(a, i, fx(a, i) - mf[a]);
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const double w = tf.wc(i);
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const double t = fx(a, i) * w;
#pragma unroll
                for (int b = 0; b <= a; ++b) Cf[tri(a, b)] = fma(t, fx(b, i), Cf[tri(a, b)]);
            }
        }
        if (want_cross) {
            double Cfx[E][D];
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int r = 0; r < D; ++r) Cfx[a][r] = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                if (PTS == PTS_AXIS_C && i == 0) continue;  // x_0 - m == 0
                const double w = tf.wc(i);
                double x[D];
                sigma_point<D, PTS>(tf, i, m, L, x);
                const int j0 = (PTS == PTS_GENERIC) ? 0 : (i - (PTS == PTS_AXIS_C ? 1 : 0)) % D;
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    if (r < j0) continue;  // structurally zero
                    const double dxr = x[r] - m[r];  // (x - mean), mtran.py:148
#pragma unroll
                    for (int a = 0; a < E; ++a)
                        Cfx[a][r] = EXACT ? __dadd_rn(Cfx[a][r], __dmul_rn(__dmul_rn(fx(a, i), w), dxr)) : fma(fx(a, i) * w, dxr, Cfx[a][r]);
                }
            }
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int r = 0; r < D; ++r) sink(a, r, Cfx[a][r]);
        }
    } else {
        // un-centred form with dense weights                    bqmtran.py:198-199, 223
        if (want_cross) {
            if constexpr (Tf::SYMW && E >= SSM_CROSS_COLS_MIN_E) {
                // Opt-in, measured slower.  Column r of the cross-covariance at a time:  Cov(f, x_r) = fx (L Wcc)[r, :]^T,
                // the product re-associated as fx (Wcc^T L^T) (bqmtran.py:223 evaluates (fx Wcc^T) L^T): N D (D + 1) / 2 +
                // E N D multiply-adds instead of E N D + E D (D + 1) / 2 (440 instead of 350 for the reentry dynamics), every
                // weight Wcc(d, i) used by ONE instruction right after its load instead of once per unrolled row.  Reentry
                // forward pass, 125 000 x 500: 13.51 / 18.66 ms against 12.72 / 17.46 ms row by row.
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    double Mr[NCAP];
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        double s = 0.0;
#pragma unroll
                        for (int d = 0; d <= r; ++d) s = fma(L[tri(r, d)], tf.Wcc(d, i), s);
                        Mr[i] = s;
                    }
#pragma unroll
                    for (int a = 0; a < E; ++a) {
                        double s = 0.0;
#pragma unroll
                        for (int i = 0; i < n; ++i) s = fma(fx(a, i), Mr[i], s);
                        sink(a, r, s);
                    }
                }
            } else {
                // G rows at a time (same sums, same order over i for every G).  Measured on the reentry forward pass with
                // predictive moments: G = 1 17.5 ms, G = 2 20.5 ms, G = 5 18.6 ms
                constexpr int G = 1;
#pragma unroll
                for (int a0 = 0; a0 < E; a0 += G) {
                    double T[G][D];
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int d = 0; d < D; ++d) T[g][d] = 0.0;
#pragma unroll
                    for (int i = 0; i < n; ++i) {
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const double w = tf.Wcc(d, i);
#pragma unroll
                            for (int g = 0; g < G; ++g) T[g][d] = fma(fx(a0 + g, i), w, T[g][d]);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int r = 0; r < D; ++r) {  // (T L^T)[a][r] = sum_{d<=r} T[a][d] L[r][d]
                            double s = 0.0;
#pragma unroll
                            for (int d = 0; d <= r; ++d) s = fma(T[g][d], L[tri(r, d)], s);
                            sink(a0 + g, r, s);
                        }
                }
            }
        }
        if constexpr (Tf::SYMW) {
            // Symmetric Wc:  fx_a Wc fx_b^T = h_a . fx_b + h_b . fx_a  with the half rows
            //   h_a[j] = Wc(j, j) / 2 * fx(a, j) + sum_{i < j} fx(a, i) Wc(i, j)
            // (each off-diagonal weight multiplies fx(a, i) fx(b, j) + fx(a, j) fx(b, i) once instead of twice):
            // E N (N + 1) / 2 + E^2 N instead of E N^2 + E (E + 1) N / 2 multiply-adds, same products, same weights.
#if SSM_SYM_COLUMNS
            // Column by column: every weight Wc(i, j) is loaded once and used by the E rows in consecutive instructions,
            // then it is dead -- with the row-by-row order below the compiler merges the E uses of a weight across the
            // unrolled rows, keeps ~70 weights live in the 63 uniform registers and spills them (R2UR.FILL + LDL).
#pragma unroll
            for (int j = 0; j < n; ++j) {
                double hc[E];
                {
                    const double w = tf.Wch(j);
#pragma unroll
                    for (int a = 0; a < E; ++a) hc[a] = fx(a, j) * w;
                }
#pragma unroll
                for (int i = 0; i < j; ++i) {
                    const double w = tf.Wc(j, i);  // = Wc(i, j), contiguous along the row
#pragma unroll
                    for (int a = 0; a < E; ++a) hc[a] = fma(fx(a, i), w, hc[a]);
                }
#pragma unroll
                for (int a = 0; a < E; ++a)
#pragma unroll
                    for (int b = 0; b < E; ++b) Cf[sym(a, b)] = fma(hc[a], fx(b, j), Cf[sym(a, b)]);  // diagonal: half of it
            }
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) Cf[tri(a, b)] = fma(-mf[a], mf[b], (a == b) ? Cf[tri(a, b)] + Cf[tri(a, b)] : Cf[tri(a, b)]);
#else
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const Tf &tw = tf.row_view(a);
                double h[NCAP];
#pragma unroll
                for (int j = 0; j < n; ++j) h[j] = fx(a, j) * tw.Wch(j);
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const double v = fx(a, i);
#pragma unroll
                    for (int j = i + 1; j < n; ++j) h[j] = fma(v, tw.Wc(i, j), h[j]);
                }
#pragma unroll
                for (int b = 0; b < E; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < n; ++j) s = fma(h[j], fx(b, j), s);
                    if (b == a) Cf[tri(a, a)] = s + s;
                    else Cf[sym(a, b)] += s;
                }
            }
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) Cf[tri(a, b)] = fma(-mf[a], mf[b], Cf[tri(a, b)]);
#endif
        } else {
#pragma unroll
        for (int a = 0; a < E; ++a) {
            const Tf &tw = tf.row_view(a);
            double row[NCAP];
#pragma unroll
            for (int j = 0; j < n; ++j) row[j] = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const double v = fx(a, i);
#pragma unroll
                for (int j = 0; j < n; ++j) row[j] = fma(v, tw.Wc(i, j), row[j]);
            }
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < n; ++j) s = fma(row[j], fx(b, j), s);
                Cf[tri(a, b)] = s - mf[a] * mf[b];
            }
        }
        }
        if (KIND == SSM_TF_TP) {
            // data-dependent model variance  mv (nu - 2 + fx iK fx^T) / (nu - 2 + N)   bqmod.py:1155-1160
            const double mv0 = tf.mv(0, 0);
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const Tf &tw = tf.row_view(a);
                double row[NCAP];
#pragma unroll
                for (int j = 0; j < n; ++j) row[j] = 0.0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const double v = fx(a, i);
#pragma unroll
                    for (int j = 0; j < n; ++j) row[j] = fma(v, tw.iK(i, j), row[j]);
                }
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    if (!tf.tp_full && b != a) continue;
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < n; ++j) s = fma(row[j], fx(b, j), s);
                    Cf[tri(a, b)] += ((tf.tp_a + s) * tf.tp_b) * mv0;
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) Cf[tri(a, b)] += tf.mv(a, b);
        }
    }
    return ok;
    }   // KIND != SSM_TF_BQR
#undef fx
}

// ------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------
struct FilterBuffers {
    const double *y;
    double *fi_mean, *fi_cov, *pr_mean, *pr_cov, *pr_xx;
    const double *init_mean, *init_cov;
    double *last_mean, *last_cov;
    const int32_t *t_offset;
    int32_t *status;
    long long n_traj, ld;
    int n_steps, k0;
    // ticket scheduler (multi-wave launches): work item = (block of THREADS trajectories, chunk of time steps)
    int *sched;        // [0] ticket counter, [1 + blk] chunks completed for trajectory block blk; NULL = plain mode
    double *state;     // carried filter state between chunks: [blk][component][thread]
    int chunk;         // time steps per work item
    int n_blocks;      // trajectory blocks
    // time window [k_lo, k_hi) of the n_steps slots to process (ssm_filter_window); resume: trajectories whose
    // status is already non-zero (failed in an earlier window) stay failed and keep their status
    int k_lo, k_hi, resume;
    // in-kernel phase-1 scoring of the FILTERED moments (filter_kernel<..., SCORE = true>, ssm_filter_scores): truth,
    // partial statistics rows [trajectory block][step of the window][ScoreRow::WP], per-trajectory time-sums of the
    // squared error (dx, ld), and per scored unit d' P^-1 d (n_steps, ld) and d = x - m (dx, n_steps, ld), nullable
    const double *x_truth;
    double *partial, *rmse_acc, *quad, *dres;
    // Dyn::time_term(k0 + k) per time slot k (models with HasTimeTerm, launches without per-trajectory time offsets), or NULL
    const double *time_tab;
    // ssm_filter_window_lower: of the symmetric outputs fi_cov / pr_cov only the lower triangles (column <= row) are
    // written -- for consumers that read nothing else (the score-only smoother): 20 of the 85 stores of a 5-D step less
    int lower_only;
};

template <class Dyn>
__global__ void time_tab_kernel(double *tab, int k0, int k_lo, int k_hi) {
    const int k = k_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (HasTimeTerm<Dyn>::value) {
        if (k < k_hi) tab[k] = Dyn::time_term((double)(k0 + k));
    }
}
// fills the table behind the launch's stream; returns NULL (in-line evaluation) when it does not apply
template <class Dyn>
inline double *make_time_tab(const FilterBuffers &b, cudaStream_t s) {
    if (!HasTimeTerm<Dyn>::value || b.t_offset || b.k_hi <= b.k_lo) return nullptr;
    double *tab = nullptr;
    if (scratch_alloc((void **)&tab, (size_t)b.n_steps * sizeof(double), s) != cudaSuccess) return nullptr;
    time_tab_kernel<Dyn><<<(b.k_hi - b.k_lo + 127) / 128, 128, 0, s>>>(tab, b.k0, b.k_lo, b.k_hi);
    return tab;
}

// NaN-fill of the outputs of failed trajectories from their failing step on, launched behind every forward-pass
// kernel (ssm_abi.cu).  All lanes of a warp walk the time steps together, so the failed lanes of a warp write the same
// rows in the same iteration; warps without a failed trajectory exit after reading their status word.  (Filling the
// tail from inside the forward pass cost the headline kernel 5-9 % through register allocation; a failed thread filling
// its own tail at the end of the forward pass issued 85 scattered 8-byte stores per step in a serial loop -- with the
// 35 % failures of the reference's TPQ weights on the coordinated-turn model that was 25 ms on top of a 35 ms pass.)
int filter_nan_fill(const FilterBuffers &b, int dx, cudaStream_t stream);

template <int DX, int DY, class TfD, class TfO>
struct FilterPar {
    TfD tf_dyn;
    TfO tf_obs;
    double dyn_par[4], obs_par[8];
    double m0[DX];
    double P0[TriSize<DX>::value];
    double GQG[TriSize<DX>::value];
    double R[TriSize<DY>::value];
    double dof, x0_dof, q_dof, r_dof, s0;  // Student family
    int fixed_dof;
    // non-additive noise: moments of the noise part of the augmented vector [x; noise] (ssinf.py:271-272, 282-283)
    double q_mean[5], q_cov[TriSize<5>::value], r_mean[DY];
    FilterBuffers b;
};

// [m; nm], blockdiag(P, Nc) in packed lower storage: the augmented moments of a non-additive model
template <int DX, int DN>
SSM_DEV void augment(const double (&m)[DX], const double (&P)[TriSize<DX>::value], const double *nm, const double *Nc,
                     double (&ma)[DX + DN], double (&Pa)[TriSize<DX + DN>::value]) {
#pragma unroll
    for (int i = 0; i < DX; ++i) ma[i] = m[i];
#pragma unroll
    for (int i = 0; i < DN; ++i) ma[DX + i] = nm[i];
#pragma unroll
    for (int r = 0; r < DX + DN; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c)
            Pa[tri(r, c)] = (r < DX) ? P[tri(r, c)] : (c < DX ? 0.0 : Nc[tri(r - DX, c - DX)]);
}

// Element (c, k, t) of a bulk array = base + rk + c * cs with rk = k * ld + t (per thread, once per step) and the
// kernel-uniform component stride cs = n_steps * ld: one 64-bit add per access instead of a 64-bit multiply chain.
template <int C, class CS>
SSM_DEV void store_vec(double *base, const CS &cs, long long rk, const double (&v)[C]) {
    if (!base) return;
    double *q = row_ptr(base, rk);
#pragma unroll
    for (int c = 0; c < C; ++c) st_stream(q + cs(c), v[c]);
}
template <int D, bool LOWER, class CS>
SSM_DEV void store_sym_impl(double *q, const CS &cs, const double (&P)[TriSize<D>::value]) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < (LOWER ? r + 1 : D); ++c) st_stream(q + cs(r * D + c), P[sym(r, c)]);
}
// lower_only (kernel-uniform): two straight store sequences behind ONE branch
template <int D, class CS>
SSM_DEV void store_sym(double *base, const CS &cs, long long rk, const double (&P)[TriSize<D>::value], const bool lower_only = false) {
    if (!base) return;
    double *q = row_ptr(base, rk);
    if (lower_only) store_sym_impl<D, true>(q, cs, P);
    else store_sym_impl<D, false>(q, cs, P);
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
#ifndef SSM_SYNC_STEPS
#define SSM_SYNC_STEPS 1
#endif
#ifndef SSM_TICKET_SCHED
#define SSM_TICKET_SCHED 1
#endif
#ifndef SSM_TICKET_CHUNK
#define SSM_TICKET_CHUNK 25
#endif
#ifndef SSM_SMEM_FX_MIN_DX
#define SSM_SMEM_FX_MIN_DX 99  // state dimension from which the function evaluations move to shared memory
                               // (measured on B200: registers win for dx = 5, 21.2 vs 27.1 ms; kept as an option)
#endif
// SCORE: the filtered moments of every step are scored against x_truth while they are in registers (squared error, error
// outer product, NLL, d' P^-1 d: score_step of ssm_scores.cuh, reduced over the CTA once per step like the smoother's
// in-kernel scoring), so a filter-only Monte-Carlo run keeps no per-trajectory moment arrays at all
// (research/gpq/icinco_demo.py:115-125 keeps none either).  Separate instantiation: the plain kernel is unchanged.
template <class Dyn, class Obs, int PTS, int NPTS, int KIND, int FAMILY, class Par, int THREADS, int MINB, bool SMEM_FX, bool SCORE = false>
__global__ void __launch_bounds__(THREADS, MINB) filter_kernel(const __grid_constant__ Par p) {
    constexpr bool SYNC_STEPS = SSM_SYNC_STEPS != 0;
    constexpr int SMT = SMEM_FX ? THREADS : 0;
    extern __shared__ double ssm_dyn_smem[];
    double *sfx = SMEM_FX ? ssm_dyn_smem + threadIdx.x : nullptr;
    constexpr int DX = Dyn::DX, DY = Obs::DY;
    constexpr int TX = TriSize<DX>::value, TY = TriSize<DY>::value;
    constexpr bool NA = !Dyn::ADDITIVE || !Obs::ADDITIVE;  // non-additive noise somewhere: exact-cancellation sums
    // lower-triangle-only stores (ssm_filter_window_lower) are honoured by the compact-sum instantiation; in the others the
    // extra branch around the stores cost the full-store launches 4 % (dense-sum reentry pass 17.5 -> 18.2 ms) and the
    // skipped stores bought them nothing, so they always write the full matrices
    constexpr bool LOWER_OK = (KIND == SSM_TF_BQR);
    const FilterBuffers &b = p.b;
    // model parameters as values: a load through the parameter block is repeated after every (rare-path) call, and the
    // re-loaded value is a new one to common-subexpression elimination
    double dpar[4], opar[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) dpar[i] = p.dyn_par[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) opar[i] = p.obs_par[i];
    const int N = b.n_steps;
    const long long ld = b.ld;
    const CompStride<NarrowStride<DX>::value> cs((long long)N * ld);  // component stride of the [component][step][trajectory] arrays
    constexpr int NSTATE = DX + TX + 2;
    constexpr int WS = ScoreRow<DX>::WP;
    static_assert(!SCORE || THREADS == SC_THREADS, "the CTA reduction of the scores is laid out for SC_THREADS threads");
    static_assert(!SCORE || FAMILY == SSM_FAMILY_GAUSS, "in-kernel scoring reads the filtered COVARIANCE (Gaussian family)");
    __shared__ double s_score[SCORE ? BlockReduce<WS>::SIZE : 1];
    __shared__ int s_ticket;
    // Ticket scheduling.  A trajectory is a 500-step serial recursion, so a plain launch is quantised in waves of
    // (resident CTAs x THREADS) whole trajectories: 125 000 trajectories = 2.2 waves cost 3 (measured -14 %).  With
    // a scheduler workspace the grid is persistent and CTAs draw (trajectory block, time chunk) items from an
    // atomic ticket counter in chunk-major order; the filter state crosses chunks through global memory.  Item
    // (blk, kc) waits for done[blk] >= kc, which a CTA that drew an EARLIER ticket -- hence already running --
    // publishes, so the wait cannot deadlock.  Arithmetic is unchanged: results are bitwise identical.
    const bool ticketed = b.sched != nullptr;
    const int n_chunks = ticketed ? (b.k_hi - b.k_lo + b.chunk - 1) / b.chunk : 1;
    const long long n_items = ticketed ? (long long)b.n_blocks * n_chunks : 0;
  for (;;) {
    long long blk = blockIdx.x;
    int kc = 0;
    if (ticketed) {
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(b.sched, 1);
        __syncthreads();
        const long long tk = s_ticket;
        if (tk >= n_items) break;
        kc = (int)(tk / b.n_blocks);
        blk = tk % b.n_blocks;
    }
    const int k_begin = b.k_lo + (ticketed ? kc * b.chunk : 0);
    const int k_end = ticketed ? min(b.k_hi, k_begin + b.chunk) : b.k_hi;
    const long long t_raw = blk * blockDim.x + threadIdx.x;
    const bool active = t_raw < b.n_traj;
    const long long t = active ? t_raw : b.n_traj - 1;  // idle lanes shadow the last trajectory, never store

    double m[DX], P[TX];  // filtered mean and covariance (Student family: scale matrix x_smat_fi)
    int fail = active ? 0 : -1, kfail = 0;
    if (kc > 0) {
        if (threadIdx.x == 0) {
            int spins = 0;
            while (atomicAdd(b.sched + 1 + blk, 0) < kc) { __nanosleep(200); ++spins; }
            if (spins) atomicAdd(b.sched + 1 + b.n_blocks, spins);  // diagnostic: total polls that had to wait
            __threadfence();
        }
        __syncthreads();
        const double *st = b.state + (blk * NSTATE) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int a = 0; a < DX; ++a) m[a] = __ldcg(st + (long long)a * blockDim.x);
#pragma unroll
        for (int a = 0; a < TX; ++a) P[a] = __ldcg(st + (long long)(DX + a) * blockDim.x);
        fail = (int)__ldcg(st + (long long)(DX + TX) * blockDim.x);
        kfail = (int)__ldcg(st + (long long)(DX + TX + 1) * blockDim.x);
    } else if (b.init_mean) {
#pragma unroll
        for (int a = 0; a < DX; ++a) m[a] = b.init_mean[(long long)a * ld + t];
#pragma unroll
        for (int r = 0; r < DX; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) P[tri(r, c)] = b.init_cov[(long long)(r * DX + c) * ld + t];
    } else {
#pragma unroll
        for (int a = 0; a < DX; ++a) m[a] = p.m0[a];
#pragma unroll
        for (int a = 0; a < TX; ++a) P[a] = (FAMILY == SSM_FAMILY_STUDENT) ? p.s0 * p.P0[a] : p.P0[a];
    }
    if (kc == 0 && b.resume && active && b.status[t] != 0) { fail = b.status[t] & 0xff; kfail = b.k_lo; }
    const double tbase = (double)(b.k0 + (b.t_offset ? b.t_offset[t] : 0));

    double ynext[DY];
#pragma unroll
    for (int a = 0; a < DY; ++a) ynext[a] = ld_stream(b.y + cs(a) + ((long long)k_begin * ld + t));
    double se_acc[SCORE ? DX : 1];
    if (SCORE) {   // per-trajectory time-sum of the squared error: continued across time chunks and windows
#pragma unroll
        for (int a = 0; a < DX; ++a) se_acc[a] = (b.rmse_acc && k_begin > 0) ? __ldcg(b.rmse_acc + (long long)a * ld + t) : 0.0;
    }

    for (int k = k_begin; k < k_end; ++k) {
        // The fully unrolled step body is ~140 KB of SASS, far beyond the instruction caches.  Re-aligning
        // the warps of the CTA once per step makes them stream the body together, so one instruction fetch
        // from L2 serves all of them instead of one per warp (profiles/: stall_no_inst, fetch-bound).
        if (SYNC_STEPS) __syncthreads();
        double sv[SCORE ? WS : 1];
        if (SCORE) {
#pragma unroll
            for (int i = 0; i < WS; ++i) sv[i] = 0.0;
        }
      // ONE exit of the step body: with SCORE every thread of the CTA must reach the reduction behind it
      do {
        if (fail) break;
        const long long rk = (long long)k * ld + t;  // row offset of step k (see store_vec)
        double yk[DY];
#pragma unroll
        for (int a = 0; a < DY; ++a) yk[a] = ynext[a];
        if (k + 1 < k_end) {
#pragma unroll
            for (int a = 0; a < DY; ++a) ynext[a] = ld_stream(b.y + cs(a) + (rk + ld));
        }
        const double time = tbase + (double)k;  // the reference passes time = k - 1, k 1-based (ssinf.py:104)
        constexpr bool TT = HasTimeTerm<Dyn>::value;
        double tt = 0.0;
        if constexpr (TT) tt = b.time_tab ? __ldg(b.time_tab + k) : Dyn::time_term(time);

        double scale = 1.0;
        if (FAMILY == SSM_FAMILY_STUDENT) {  // ssinf.py:650-660
            double dof_pr = p.dof;
            if (p.fixed_dof) dof_pr = fmin(fmin(p.x0_dof + (double)k * DY, p.q_dof), p.r_dof);
            scale = (dof_pr - 2.0) / dof_pr;
        }

        // ---- time update: predictive state moments (ssinf.py:276-279 / 669-676) ----------------
        double mp[DX], Pp[TX];
        const bool want_xx = b.pr_xx != nullptr;
        double *q_xx = b.pr_xx ? row_ptr(b.pr_xx, rk) : nullptr;
        bool ok;
        if constexpr (Dyn::ADDITIVE) {
            ok = moment_transform<DX, DX, PTS, NPTS, KIND, SMT, NA>(
                p.tf_dyn, m, P,
                [&](const double (&x)[DX], double (&o)[DX]) {
                    const double q0[Dyn::DQ] = {};
                    if constexpr (TT) Dyn::template f_tt<false>(dpar, x, q0, tt, o);
                    else Dyn::template f<false>(dpar, x, q0, time, o);
                },
                mp, Pp, want_xx,
                [&](int a, int c, double v) { st_stream(q_xx + cs(a * DX + c), v); },  // Cov(x_k, x_{k-1})[a][c] -> pr_xx_cov[a][c][k][t]
                sfx);
        } else {
            // non-additive process noise: transform of the augmented vector [x; q], cross-covariance trimmed to its
            // first dx columns (ssinf.py:271-272, 294)
            constexpr int DD = DX + Dyn::DQ;
            double ma[DD], Pa[TriSize<DD>::value];
            augment<DX, Dyn::DQ>(m, P, p.q_mean, p.q_cov, ma, Pa);
            ok = moment_transform<DD, DX, PTS, NPTS, KIND, SMT, NA>(
                p.tf_dyn, ma, Pa,
                [&](const double (&xq)[DD], double (&o)[DX]) {
                    double x[DX], q[Dyn::DQ];
#pragma unroll
                    for (int i = 0; i < DX; ++i) x[i] = xq[i];
#pragma unroll
                    for (int i = 0; i < Dyn::DQ; ++i) q[i] = xq[DX + i];
                    if constexpr (TT) Dyn::template f_tt<true>(dpar, x, q, tt, o);
                    else Dyn::template f<true>(dpar, x, q, time, o);
                },
                mp, Pp, want_xx,
                [&](int a, int c, double v) {
                    if (c < DX) st_stream(q_xx + cs(a * DX + c), v);
                },
                sfx);
        }
        if (!ok) { fail = SSM_FAIL_CHOL_DYN; kfail = k; break; }
        if (FAMILY == SSM_FAMILY_STUDENT) {
            if (b.pr_cov) {
                double Cp[TX];
#pragma unroll
                for (int a = 0; a < TX; ++a) Cp[a] = Dyn::ADDITIVE ? Pp[a] + p.GQG[a] : Pp[a];  // x_cov_pr, ssinf.py:674-675 (additive noise only)
                store_sym<DX>(b.pr_cov, cs, rk, Cp, LOWER_OK && b.lower_only != 0);
            }
#pragma unroll
            for (int a = 0; a < TX; ++a) Pp[a] = Dyn::ADDITIVE ? fma(scale, Pp[a], p.s0 * p.GQG[a]) : scale * Pp[a];  // x_smat_pr, ssinf.py:672, 674-676
        } else {
            if (Dyn::ADDITIVE) {
#pragma unroll
                for (int a = 0; a < TX; ++a) Pp[a] += p.GQG[a];  // ssinf.py:278-279
            }
            store_sym<DX>(b.pr_cov, cs, rk, Pp, LOWER_OK && b.lower_only != 0);
        }
        store_vec<DX>(b.pr_mean, cs, rk, mp);

        // ---- predictive measurement moments (ssinf.py:287-291 / 684-693) ------------------------
        double my[DY], Sy[TY], Syx[DY][DX];
        if constexpr (Obs::ADDITIVE) {
            ok = moment_transform<DX, DY, PTS, NPTS, KIND, SMT, NA>(
                p.tf_obs, mp, Pp,
                [&](const double (&x)[DX], double (&o)[DY]) {
                    const double r0[DY] = {};
                    Obs::template h<false>(opar, x, r0, time, o);
                },
                my, Sy, true,
                [&](int a, int c, double v) { Syx[a][c] = v; },
                sfx);
        } else {
            // non-additive measurement noise: [x; r] (ssinf.py:282-283), cross-covariance trimmed (:293)
            constexpr int DO = DX + DY;
            double ma[DO], Pa[TriSize<DO>::value];
            augment<DX, DY>(mp, Pp, p.r_mean, p.R, ma, Pa);
            ok = moment_transform<DO, DY, PTS, NPTS, KIND, SMT, NA>(
                p.tf_obs, ma, Pa,
                [&](const double (&xr)[DO], double (&o)[DY]) {
                    double x[DX], r[DY];
#pragma unroll
                    for (int i = 0; i < DX; ++i) x[i] = xr[i];
#pragma unroll
                    for (int i = 0; i < DY; ++i) r[i] = xr[DX + i];
                    Obs::template h<true>(opar, x, r, time, o);
                },
                my, Sy, true,
                [&](int a, int c, double v) {
                    if (c < DX) Syx[a][c] = v;
                },
                sfx);
        }
        if (!ok) { fail = SSM_FAIL_CHOL_OBS; kfail = k; break; }
        if (FAMILY == SSM_FAMILY_STUDENT) {
#pragma unroll
            for (int a = 0; a < TY; ++a) Sy[a] = Obs::ADDITIVE ? fma(scale, Sy[a], p.s0 * p.R[a]) : scale * Sy[a];  // ssinf.py:687-693
#pragma unroll
            for (int a = 0; a < DY; ++a)
#pragma unroll
                for (int d = 0; d < DX; ++d) Syx[a][d] *= scale;
        } else if (Obs::ADDITIVE) {
#pragma unroll
            for (int a = 0; a < TY; ++a) Sy[a] += p.R[a];  // ssinf.py:290-291
        }

        // ---- measurement update (ssinf.py:321-323 / 724-736) -----------------------------------
        bool fin = true;
#pragma unroll
        for (int a = 0; a < TY; ++a) fin = fin && finite_d(Sy[a]);
#pragma unroll
        for (int a = 0; a < DY; ++a)
#pragma unroll
            for (int d = 0; d < DX; ++d) fin = fin && finite_d(Syx[a][d]);
        if (!fin) { fail = SSM_FAIL_NONFINITE_GAIN; kfail = k; break; }
        double K[DX][DY], Ls[TY];
        ok = spd_gain<DY, DX>(Sy, Syx, K, Ls);
        if (!ok) { fail = SSM_FAIL_CHOL_GAIN; kfail = k; break; }
        double e[DY];
#pragma unroll
        for (int a = 0; a < DY; ++a) e[a] = yk[a] - my[a];
#pragma unroll
        for (int d = 0; d < DX; ++d) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DY; ++a) s = fma(K[d][a], e[a], s);
            m[d] = mp[d] + s;
        }
        {
            double KS[DX][DY];  // gain . Sy
#pragma unroll
            for (int d = 0; d < DX; ++d)
#pragma unroll
                for (int a = 0; a < DY; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < DY; ++c) s = fma(K[d][c], Sy[sym(c, a)], s);
                    KS[d][a] = s;
                }
#pragma unroll
            for (int r = 0; r < DX; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < DY; ++a) s = fma(KS[r][a], K[c][a], s);
                    P[tri(r, c)] = Pp[tri(r, c)] - s;
                }
        }
        store_vec<DX>(b.fi_mean, cs, rk, m);
        store_sym<DX>(b.fi_cov, cs, rk, P, LOWER_OK && b.lower_only != 0);  // Student: x_cov_fi = x_smat_pr - K Sy K^T (ssinf.py:727)
        if (FAMILY == SSM_FAMILY_STUDENT) {
            // delta = chol(Sy)^-1 e ; x_smat_fi = (dof + delta'delta) / (dof + dy) x_cov_fi   ssinf.py:731-733
            double dd = 0.0, z[DY];
#pragma unroll
            for (int i = 0; i < DY; ++i) {
                double s = e[i];
#pragma unroll
                for (int c = 0; c < i; ++c) s = fma(-Ls[tri(i, c)], z[c], s);
                z[i] = m_div(s, Ls[tri(i, i)]);
                dd = fma(z[i], z[i], dd);
            }
            const double sc = (p.dof + dd) / (p.dof + (double)DY);
#pragma unroll
            for (int a = 0; a < TX; ++a) P[a] *= sc;
        }
        if constexpr (SCORE) {
            if (active) {   // score (m, P) of step k: utils.py:18-148 via score_step
                double d[DX], se[DX], qf;
                const double *qx = row_ptr(b.x_truth, rk);
#pragma unroll
                for (int a = 0; a < DX; ++a) d[a] = ld_stream(qx + cs(a)) - m[a];
                score_step<DX>(d, P, sv, se, &qf);
                if (b.quad) st_stream(b.quad + rk, qf);
                if (b.dres) {
                    double *qd = row_ptr(b.dres, rk);
#pragma unroll
                    for (int a = 0; a < DX; ++a) st_stream(qd + cs(a), d[a]);
                }
#pragma unroll
                for (int a = 0; a < DX; ++a) se_acc[a] += se[a];
            }
        }
      } while (0);
        if constexpr (SCORE)
            block_reduce_store<WS>(sv, s_score, k, b.partial + ((long long)blk * (b.k_hi - b.k_lo) + (k - b.k_lo)) * WS);
    }
    if constexpr (SCORE) {
        if (b.rmse_acc && active) {
#pragma unroll
            for (int a = 0; a < DX; ++a) __stcg(b.rmse_acc + (long long)a * ld + t, (fail && k_end >= b.k_hi) ? qnan() : se_acc[a]);
        }
    }

    if (k_end < b.k_hi) {
        // hand the state to whichever CTA draws the next chunk of this trajectory block
        double *st = b.state + (blk * NSTATE) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int a = 0; a < DX; ++a) __stcg(st + (long long)a * blockDim.x, m[a]);
#pragma unroll
        for (int a = 0; a < TX; ++a) __stcg(st + (long long)(DX + a) * blockDim.x, P[a]);
        __stcg(st + (long long)(DX + TX) * blockDim.x, (double)fail);
        __stcg(st + (long long)(DX + TX + 1) * blockDim.x, (double)kfail);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(b.sched + 1 + blk, kc + 1);
    } else if (active) {
        if (fail) {
            // the NaN rows [kfail, k_hi) of the bulk outputs are written by filter_nan_fill() after this kernel
    #pragma unroll
            for (int a = 0; a < DX; ++a) m[a] = qnan();
    #pragma unroll
            for (int a = 0; a < TX; ++a) P[a] = qnan();
        }
        if (b.last_mean) {
    #pragma unroll
            for (int a = 0; a < DX; ++a) b.last_mean[(long long)a * ld + t] = m[a];
        }
        if (b.last_cov) {
    #pragma unroll
            for (int r = 0; r < DX; ++r)
    #pragma unroll
                for (int c = 0; c < DX; ++c) b.last_cov[(long long)(r * DX + c) * ld + t] = P[sym(r, c)];
        }
        if (!(b.resume && b.status[t] != 0)) b.status[t] = fail ? (((kfail + 1) << 8) | fail) : 0;
    }
    if (!ticketed) break;
  }
}

// ------------------------------------------------------------------------------------------------
// host side: lowering of ssm_desc into the parameter block and launch
// ------------------------------------------------------------------------------------------------
struct HostTfInfo {
    int pts;      // PTS_*
    double c;     // axis scale
};

// classify a unit point set: [0 | cI | -cI], [cI | -cI] or generic
inline HostTfInfo classify_points(const ssm_transform &tf) {
    const int D = tf.dim_in, N = tf.n_pts;
    HostTfInfo r{PTS_GENERIC, 0.0};
    auto U = [&](int d, int i) { return tf.points[d * N + i]; };
    for (int base = 0; base <= 1; ++base) {
        if (N != 2 * D + base) continue;
        const double c = U(0, base);
        bool ok = c > 0.0;
        for (int d = 0; d < D && ok; ++d)
            for (int i = 0; i < N && ok; ++i) {
                double want = 0.0;
                if (i >= base && i - base < D && i - base == d) want = c;
                if (i - base >= D && i - base - D == d) want = -c;
                ok = (U(d, i) == want);
            }
        if (ok) { r.pts = base ? PTS_AXIS_C : PTS_AXIS; r.c = c; return r; }
    }
    return r;
}

// Wc == Wc^T bit for bit (what the half-row form of the fast path needs; the reference's weights always are)
inline bool wc_symmetric(const ssm_transform &tf) {
    if (tf.kind == SSM_TF_SP || !SSM_SYM_WC) return true;
    const int N = tf.n_pts;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j)
            if (!(tf.Wc[i * N + j] == tf.Wc[j * N + i])) return false;
    return true;
}

// Weights of a BQ transform on [0 | cI | -cI] points that are invariant under every coordinate reflection, bit for bit
// (the BQR instantiation reads one representative of each class of equal weights).  SSM_REFL=0 forces the dense path.
inline bool weights_reflective(const ssm_transform &tf, const HostTfInfo &info) {
    const char *env = getenv("SSM_REFL");
    if (env && atoi(env) == 0) return false;
    if (tf.kind != SSM_TF_BQ || info.pts != PTS_AXIS_C || !tf.wm || !tf.Wc || !tf.Wcc) return false;
    const int D = tf.dim_in, N = tf.n_pts;
    if (N != 2 * D + 1) return false;
    auto W = [&](int i, int j) { return tf.Wc[i * N + j]; };
    auto mirror = [&](int i) { return i == 0 ? 0 : (i <= D ? i + D : i - D); };
    auto axis = [&](int i) { return i == 0 ? -1 : (i - 1) % D; };
    for (int i = 0; i < N; ++i) {
        if (!(tf.wm[i] == tf.wm[mirror(i)])) return false;
        for (int j = 0; j < N; ++j) {
            if (!(W(i, j) == W(j, i))) return false;
            // reflecting the axis of point i alone, the axis of point j alone, or (same axis) both
            if (axis(i) != axis(j)) { if (!(W(i, j) == W(mirror(i), j)) || !(W(i, j) == W(i, mirror(j)))) return false; }
            else if (!(W(i, j) == W(mirror(i), mirror(j)))) return false;
        }
    }
    for (int d = 0; d < D; ++d)
        for (int i = 0; i < N; ++i) {
            const double w = tf.Wcc[d * N + i];
            if (axis(i) != d) { if (!(w == 0.0)) return false; }
            else if (!(w == -tf.Wcc[d * N + mirror(i)])) return false;
        }
    return true;
}

template <int D>
inline void pack_lower(const double *full, double *packed) {
    for (int r = 0; r < D; ++r)
        for (int c = 0; c <= r; ++c) packed[r * (r + 1) / 2 + c] = full[r * D + c];
}

template <class Tf>
inline void fill_tf_common(Tf &o, const ssm_transform &tf, const HostTfInfo &info) {
    o.n = tf.n_pts;
    o.tp_full = tf.tp_full_matrix;
    o.c = info.c;
    o.tp_a = tf.nu - 2.0;
    o.tp_b = 1.0 / (tf.nu - 2.0 + (double)tf.n_pts);
    const int E = tf.dim_out;
    for (int a = 0; a < E; ++a)
        for (int b = 0; b < E; ++b) o.mv_[a][b] = 0.0;
    if (tf.kind == SSM_TF_BQ && tf.model_var)
        for (int a = 0; a < E; ++a)
            for (int b = 0; b < E; ++b) o.mv_[a][b] = tf.model_var[a * E + b];
    if (tf.kind == SSM_TF_TP && tf.model_var) o.mv_[0][0] = tf.model_var[0];
}

template <int D, int E, int NCAP, int KIND, int PTS>
inline void fill_tf(TfConst<D, E, NCAP, KIND, PTS> &o, const ssm_transform &tf, const HostTfInfo &info) {
    memset(&o, 0, sizeof(o));
    fill_tf_common(o, tf, info);
    const int N = tf.n_pts;
    for (int i = 0; i < N; ++i) {
        o.wm_[i] = tf.wm[i];
        o.wc_[i] = tf.Wc[i * N + i];
    }
    if constexpr (KIND == SSM_TF_BQR) {
        // compact tables of the reflection-symmetric form (moment_transform): S in Wc_ / wch_, alpha in wc_, the one
        // cross-covariance weight per axis in Wcc_[d][0]; wm_[0 .. D] already are the weights of g
        auto W = [&](int i, int j) { return tf.Wc[i * N + j]; };
        for (int i = 0; i <= D; ++i)
            for (int j = 0; j <= D; ++j) o.Wc_[i][j] = W(i, j);
        for (int j = 0; j < D; ++j) {
            o.Wc_[1 + j][1 + j] = 0.5 * (W(1 + j, 1 + j) + W(1 + j, 1 + D + j));
            o.wc_[j] = 0.5 * (W(1 + j, 1 + j) - W(1 + j, 1 + D + j));
            o.Wcc_[j][0] = tf.Wcc[j * N + 1 + j];
        }
        for (int i = 0; i <= D; ++i) o.wch_[i] = 0.5 * o.Wc_[i][i];
        return;
    }
    if (KIND != SSM_TF_SP) {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) o.Wc_[i % o.NW][j] = tf.Wc[i * N + j];
        for (int i = 0; i < N; ++i) o.wch_[i % o.NW] = 0.5 * tf.Wc[i * N + i];
        for (int d = 0; d < D; ++d)
            for (int i = 0; i < N; ++i) o.Wcc_[d % o.DC][i] = tf.Wcc[d * N + i];
    }
    if (KIND == SSM_TF_TP)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) o.iK_[i % o.NK][j] = tf.iK[i * N + j];
    if (PTS == PTS_GENERIC)
        for (int d = 0; d < D; ++d)
            for (int i = 0; i < N; ++i) o.U_[d % o.DU][i] = tf.points[d * N + i];
}

void set_error(const char *fmt, ...);

struct FilterLaunch {
    const ssm_desc *desc;
    FilterBuffers buf;
    cudaStream_t stream;
    double *stats = nullptr;  // scoring launches (buf.x_truth != NULL): rows [k_lo, k_hi) of the (n_steps, ScoreRow::W) statistics
};

// fast path launcher (all weights in the parameter block)
template <class Dyn, class Obs, int PTS, int NPTS, int KIND, int FAMILY, int THREADS, int MINB, bool SCORE = false>
int launch_filter_const(const FilterLaunch &L, const HostTfInfo &id, const HostTfInfo &io) {
    constexpr int DX = Dyn::DX, DY = Obs::DY;
    using TfD = TfConst<DX, DX, NPTS, KIND, PTS>;
    using TfO = TfConst<DX, DY, NPTS, KIND, PTS>;
    using Par = FilterPar<DX, DY, TfD, TfO>;
    static_assert(sizeof(Par) <= 32000, "kernel parameter block too large");
    const ssm_desc &d = *L.desc;
    Par *pp = new Par;
    Par &p = *pp;
    memset(pp, 0, sizeof(Par));
    fill_tf(p.tf_dyn, d.tf_dyn, id);
    fill_tf(p.tf_obs, d.tf_obs, io);
    for (int i = 0; i < 4; ++i) p.dyn_par[i] = d.dyn_par[i];
    for (int i = 0; i < 8; ++i) p.obs_par[i] = d.obs_par[i];
    for (int i = 0; i < DX; ++i) p.m0[i] = d.m0[i];
    pack_lower<DX>(d.P0, p.P0);
    pack_lower<DX>(d.GQG, p.GQG);
    pack_lower<DY>(d.R, p.R);
    p.dof = d.dof; p.x0_dof = d.x0_dof; p.q_dof = d.q_dof; p.r_dof = d.r_dof;
    p.s0 = (d.family == SSM_FAMILY_STUDENT) ? (d.dof - 2.0) / d.dof : 1.0;
    p.fixed_dof = d.fixed_dof;
    p.b = L.buf;
    if (!stride_fits<DX>(L.buf.n_steps, L.buf.ld)) {
        delete pp;
        set_error("n_steps * ld = %lld elements per component: this model addresses components with a 32-bit stride (< 2^32); run the trajectories in chunks", (long long)L.buf.n_steps * L.buf.ld);
        return SSM_E_UNSUPPORTED;
    }
    const long long blocks = (L.buf.n_traj + THREADS - 1) / THREADS;
    constexpr bool SMEM_FX = (DX >= SSM_SMEM_FX_MIN_DX);
    auto kern = filter_kernel<Dyn, Obs, PTS, NPTS, KIND, FAMILY, Par, THREADS, MINB, SMEM_FX, SCORE>;
    double *partial = nullptr;
    const int WLEN = L.buf.k_hi - L.buf.k_lo;
    if (SCORE) {   // one partial statistics row per (trajectory block, step of the window)
        if (scratch_alloc((void **)&partial, (size_t)blocks * WLEN * ScoreRow<DX>::WP * sizeof(double), L.stream) != cudaSuccess) {
            delete pp; set_error("cudaMallocAsync failed"); return SSM_E_CUDA;
        }
        p.b.partial = partial;
    }
    const size_t smem = SMEM_FX ? sizeof(double) * DX * NPTS * THREADS : 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // multi-wave launches: persistent grid + ticket scheduler over (trajectory block, time chunk) items
    long long grid = blocks;
    void *work = nullptr;
    int occ = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
    const long long cap = (long long)occ * sms;
    // developer overrides: SSM_TICKET=0 disables the scheduler, SSM_TICKET_CHUNK=<steps> sets the item length
    const char *env_on = getenv("SSM_TICKET"), *env_chunk = getenv("SSM_TICKET_CHUNK");
    const int win = L.buf.k_hi - L.buf.k_lo;
    int CHUNK = env_chunk ? atoi(env_chunk) : SSM_TICKET_CHUNK;
    if (!env_chunk && cap > 0 && win < 8 * CHUNK) {
        // short windows (host-streaming driver): pick the item length that minimises the rounds x length product
        long long best = -1;
        for (int c : {25, 20, 16, 12, 10, 8}) {
            const long long items = blocks * ((win + c - 1) / c);
            const long long cost = ((items + cap - 1) / cap) * c;
            if (best < 0 || cost < best) { best = cost; CHUNK = c; }
        }
    }
    const bool want_ticket = SSM_TICKET_SCHED && !(env_on && atoi(env_on) == 0);
    if (want_ticket && CHUNK > 0 && cap > 0 && blocks > cap && win >= 2 * CHUNK) {
        const size_t n_int = ((size_t)blocks + 2 + 1) / 2 * 2;  // ticket, done[blocks], wait counter; doubles stay 8-byte aligned
        const size_t bytes = n_int * sizeof(int) + (size_t)blocks * THREADS * (DX + TriSize<DX>::value + 2) * sizeof(double);
        if (scratch_alloc((void **)&work, bytes, L.stream) != cudaSuccess) { delete pp; set_error("cudaMallocAsync failed"); return SSM_E_CUDA; }
        cudaMemsetAsync(work, 0, n_int * sizeof(int), L.stream);
        p.b.sched = (int *)work;
        p.b.state = (double *)((int *)work + n_int);
        p.b.chunk = CHUNK;
        p.b.n_blocks = (int)blocks;
        grid = cap;
    }
    double *ttab = make_time_tab<Dyn>(L.buf, L.stream);
    p.b.time_tab = ttab;
    kern<<<(unsigned)grid, THREADS, smem, L.stream>>>(p);
    cudaError_t err = cudaGetLastError();
    if (ttab) cudaFreeAsync(ttab, L.stream);
    if (err == cudaSuccess && filter_nan_fill(L.buf, DX, L.stream) != SSM_OK) err = cudaErrorUnknown;
    if (SCORE) {
        const long long row = (long long)WLEN * ScoreRow<DX>::WP;
        scores_finalize_packed_kernel<<<(unsigned)((row + 31) / 32), dim3(32, FIN_GROUPS), 0, L.stream>>>(
            partial, L.stats + (long long)L.buf.k_lo * ScoreRow<DX>::W, (int)blocks, WLEN, DX);
        if (err == cudaSuccess) err = cudaGetLastError();
        cudaFreeAsync(partial, L.stream);
    }
    if (work && getenv("SSM_TICKET_DEBUG")) {
        int waits = 0;
        cudaMemcpyAsync(&waits, (int *)work + 1 + blocks, sizeof(int), cudaMemcpyDeviceToHost, L.stream);
        cudaStreamSynchronize(L.stream);
        fprintf(stderr, "[ssm ticket] blocks=%lld grid=%lld chunk=%d dependency polls that waited: %d\n", blocks, grid, CHUNK, waits);
    }
    if (work) cudaFreeAsync(work, L.stream);
    delete pp;
    return err == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

}  // namespace ssm
