// K2p: fused forward pass with a WARP PAIR per 32 trajectories (round 2).
//
// The one-thread-per-trajectory kernel (ssm_filter.cuh) holds the whole 5-D step in one thread: 55 function values,
// the Cholesky factor, the state -- 168 registers with spills, 12 warps per SM, a 84 KB straight-line body.  Here two
// warps of a CTA share 32 trajectories and split the sigma points of every moment transform by SIGN:
//
//   main   warp: state (m, P), both Cholesky factorisations, the centre point and the D "+" points, its half of the
//                weighted sums, the measurement update, the stores of the filtered moments;
//   helper warp: the D "-" points, its half of the weighted sums, the complete cross-covariance rows
//                (fx Wcc^T L^T), the stores of the predictive moments.
//
// The halves meet through shared memory ([slot][trajectory] doubles, conflict-free) and named barriers
// (bar.sync id, 64 -- exactly the two warps of a pair): three exchanges per transform,
//   (1) main -> helper: mean and Cholesky factor;  (2) both: function values, helper's partial mean;
//   (3) helper -> main: partial covariance (and Cov(h, x) rows for the measurement transform).
// Roles are warp-uniform and compiled as separate code paths, so every weight is still an immediate constant-bank
// operand (a lane pair INSIDE a warp would need lane-dependent weights: one LDC per DFMA).
// The serial parts (Cholesky, gain, update) run once per trajectory on the main warp -- no duplicated fp64 work;
// the helper idles at a barrier meanwhile, which costs occupancy, not issue slots.
//
// Arithmetic: same formulas as moment_transform<..., SSM_TF_BQ / SSM_TF_TP> (bq/bqmtran.py:175, 198-199, 223,
// 394-415), with the double sum  sum_i sum_j f(a,i) W(i,j) f(b,j)  associated as  sum_i f(a,i) (sum_j W(i,j) f(b,j))
// and split over i between the two warps: the results differ from the single-thread kernel by rounding only (the
// un-centred form's own noise floor, tests/test_gpu_parity.py::test_bq_noise_floor, applies to both).
// Supported: additive models, [0 | cI | -cI] point sets (UT, fully-symmetric degree 3), BQ / TP transforms, Gaussian
// family.  Everything else stays on filter_kernel.
#pragma once
#include "ssm_filter.cuh"

namespace ssm {

template <int D, int E, int KIND>
struct TfPair {
    static constexpr int N = 2 * D + 1;
    static constexpr int NK = (KIND == SSM_TF_TP) ? N : 1;
    int tp_full;
    double c;           // axis point scale (read by sigma_point<D, PTS_AXIS_C>)
    double tp_a, tp_b;  // nu - 2, 1 / (nu - 2 + N)
    double wm[N];
    double W[N][N];
    double Wcc[D][N];
    double mv[E][E];
    double iK[NK][N];
};

template <int DX, int DY, int KIND>
struct PairPar {
    TfPair<DX, DX, KIND> tf_dyn;
    TfPair<DX, DY, KIND> tf_obs;
    double dyn_par[4], obs_par[8];
    double m0[DX];
    double P0[TriSize<DX>::value];
    double GQG[TriSize<DX>::value];
    double R[TriSize<DY>::value];
    int zero;  // always 0, unknown to the compiler (see weights_in_loop)
    FilterBuffers b;
};

// The weights of trip b of a rolled loop: the same table, addressed through an offset the compiler cannot fold
// ((b & 0) * 16).  With a loop-invariant address the compiler hoists all ~70 weight loads out of the loop, copies them
// from uniform to vector registers and spills them to LOCAL memory (ptxas: 1 KB of spills at 128 registers); with the
// opaque offset every weight is one LDCU.64 c[0x0][UR + imm] next to its DFMA.
template <class Tf>
SSM_DEV const Tf &weights_in_loop(const Tf &tf, int b, int zero) {
    return *(const Tf *)((const char *)&tf + (size_t)((b & zero) * 16));
}

// shared-memory slots of one trajectory (doubles; slot s of lane l lives at xs[s * TPB])
template <int DX, int DY, int KIND>
struct PairSlots {
    static constexpr int TX = TriSize<DX>::value, TY = TriSize<DY>::value;
    static constexpr int NV = (KIND == SSM_TF_TP) ? 2 : 1;  // covariance (+ fx iK fx^T) partial sums
    static constexpr int XA = 0;                         // main -> helper: mean (DX), Cholesky factor (TX)
    static constexpr int XFM = XA + DX + TX;             // main's function values  [a][0..DX]
    static constexpr int XFH = XFM + DX * (DX + 1);      // helper's function values [a][0..DX-1]
    static constexpr int XMH = XFH + DX * DX;            // helper's partial mean
    static constexpr int XC = XMH + DX;                  // helper -> main: partial sums, Cov(h, x)
    static constexpr int CDYN = NV * TX, COBS = NV * TY;
    static constexpr int XCSZ = CDYN > COBS ? CDYN : COBS;
    static constexpr int XCM = XC + XCSZ;                // main's own partial sums (kept out of its registers); aliased by XP
    static constexpr int XCMSZ = (CDYN > NV * TY + DY * DX) ? CDYN : NV * TY + DY * DX;
    static constexpr int XP = XCM;                       // main -> helper: predictive covariance (for its store)
    static constexpr int TOTAL = XCM + XCMSZ;
};

// Ordered shared-memory accesses of the weighted-sum loops: ptxas otherwise hoists the loads of all E output rows in
// front of the first one (no aliasing with the stores in between), which keeps 25 more doubles live and spills.
SSM_DEV double lds_o(const double *p) { return *(const volatile double *)p; }
SSM_DEV void sts_o(double *p, double v) { *(volatile double *)p = v; }

SSM_DEV void pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---- main warp: one moment transform ---------------------------------------------------------------
// The loop over the output row b is a REAL loop (not unrolled): unrolled, the compiler merges the loads of every weight
// W(i, j) across the E rows and keeps all of them live in registers and uniform registers (spilling both); rolled,
// each weight is an immediate operand used once per trip, the body is E times smaller and only the function values of
// this warp's own points stay in registers (for the second factor of the covariance sums).
// CROSS: the measurement transform's Cov(h, x) rows are computed here (the main warp still holds the Cholesky factor
// and needs the result for the gain); the dynamics transform's rows belong to the helper.
template <int D, int E, int KIND, bool CROSS, int TPB, class SL, class Tf, class F>
SSM_DEV bool pair_main_transform(const Tf &tf, const double (&m)[D], const double (&P)[TriSize<D>::value], F f,
                                 double (&mf)[E], double (&Cf)[TriSize<E>::value], double (&Cfx)[E][D], double *xs,
                                 const int bar_id, const int zero) {
    constexpr int TD = TriSize<D>::value, TE = TriSize<E>::value;
    double L[TD];
    const bool ok = chol_lower<D>(P, L);
#pragma unroll
    for (int r = 0; r < D; ++r) xs[(SL::XA + r) * TPB] = m[r];
#pragma unroll
    for (int a = 0; a < TD; ++a) xs[(SL::XA + D + a) * TPB] = L[a];
    pair_bar(bar_id);  // (1) the helper may generate its points
    // centre and "+" points: each value goes to shared memory as soon as it exists (the partner needs it there anyway)
    // and is re-read after the barrier, so no function value is live while the model is being evaluated
    double sp[E];
#pragma unroll
    for (int a = 0; a < E; ++a) sp[a] = 0.0;
#pragma unroll
    for (int q = 0; q <= D; ++q) {
        double x[D], o[E];
        sigma_point<D, PTS_AXIS_C>(tf, q, m, L, x);
        f(x, o);
#pragma unroll
        for (int a = 0; a < E; ++a) {
            xs[(SL::XFM + a * (D + 1) + q) * TPB] = o[a];
            sp[a] = fma(o[a], tf.wm[q], sp[a]);
        }
    }
    pair_bar(bar_id);  // (2) function values of both halves are visible
    double Fm[E][D + 1];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int q = 0; q <= D; ++q) Fm[a][q] = xs[(SL::XFM + a * (D + 1) + q) * TPB];
#pragma unroll
    for (int a = 0; a < E; ++a) mf[a] = sp[a] + xs[(SL::XMH + a) * TPB];  // mean_f = fx . wm   bqmtran.py:175
#pragma unroll 1
    for (int b = 0; b < E; ++b) {
        const Tf &tw = weights_in_loop(tf, b, zero);
        double fo[D + 1], fh[D];
#pragma unroll
        for (int j = 0; j <= D; ++j) fo[j] = xs[(SL::XFM + b * (D + 1) + j) * TPB];
#pragma unroll
        for (int q = 0; q < D; ++q) fh[q] = xs[(SL::XFH + b * D + q) * TPB];
        {
            double g[D + 1];  // g_i = sum_j W(i, j) f(b, j) for the rows i this warp owns
#pragma unroll
            for (int i = 0; i <= D; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= D; ++j) s = fma(fo[j], tw.W[i][j], s);
#pragma unroll
                for (int q = 0; q < D; ++q) s = fma(fh[q], tw.W[i][D + 1 + q], s);
                g[i] = s;
            }
#pragma unroll
            for (int a = 0; a < E; ++a) {
                if (a < b) continue;
                double s = 0.0;
#pragma unroll
                for (int i = 0; i <= D; ++i) s = fma(Fm[a][i], g[i], s);
                xs[(SL::XCM + a * (a + 1) / 2 + b) * TPB] = s;
            }
        }
        if (KIND == SSM_TF_TP) {  // fx iK fx^T, bqmod.py:1155-1158
            double g[D + 1];
#pragma unroll
            for (int i = 0; i <= D; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= D; ++j) s = fma(fo[j], tw.iK[i % Tf::NK][j], s);
#pragma unroll
                for (int q = 0; q < D; ++q) s = fma(fh[q], tw.iK[i % Tf::NK][D + 1 + q], s);
                g[i] = s;
            }
#pragma unroll
            for (int a = 0; a < E; ++a) {
                if (a < b || (!tf.tp_full && a != b)) continue;
                double s = 0.0;
#pragma unroll
                for (int i = 0; i <= D; ++i) s = fma(Fm[a][i], g[i], s);
                xs[(SL::XCM + TE + a * (a + 1) / 2 + b) * TPB] = s;
            }
        }
        if (CROSS) {  // row b of fx Wcc^T L^T over ALL points   bqmtran.py:223
            double T[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= D; ++j) s = fma(fo[j], tw.Wcc[d][j], s);
#pragma unroll
                for (int q = 0; q < D; ++q) s = fma(fh[q], tw.Wcc[d][D + 1 + q], s);
                T[d] = s;
            }
#pragma unroll
            for (int r = 0; r < D; ++r) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d <= r; ++d) s = fma(T[d], L[tri(r, d)], s);
                xs[(SL::XCM + SL::NV * TE + b * D + r) * TPB] = s;
            }
        }
    }
    pair_bar(bar_id);  // (3) the helper's partial sums are visible
    const double mv0 = tf.mv[0][0];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            const double s = xs[(SL::XCM + tri(a, b)) * TPB] + xs[(SL::XC + tri(a, b)) * TPB];
            double c = s - mf[a] * mf[b];  // fx Wc fx^T - m m^T   bqmtran.py:198-199
            if (KIND == SSM_TF_TP) {
                if (tf.tp_full || a == b) {
                    const double v = xs[(SL::XCM + TE + tri(a, b)) * TPB] + xs[(SL::XC + TE + tri(a, b)) * TPB];
                    c += ((tf.tp_a + v) * tf.tp_b) * mv0;  // bqmod.py:1155-1160
                }
            } else {
                c += tf.mv[a][b];
            }
            Cf[tri(a, b)] = c;
        }
    if (CROSS) {
#pragma unroll
        for (int a = 0; a < E; ++a)
#pragma unroll
            for (int d = 0; d < D; ++d) Cfx[a][d] = xs[(SL::XCM + SL::NV * TE + a * D + d) * TPB];
    }
    return ok;
}

// ---- helper warp: one moment transform ---------------------------------------------------------------
// post(m): called with the transform's input mean once it is loaded (the measurement transform stores the predictive
// moments there); CROSS: the helper computes the complete cross-covariance rows Cov(f_b, x) (bqmtran.py:223) and hands
// them to sink(b, row)
template <int D, int E, int KIND, bool CROSS, int TPB, class SL, class Tf, class F, class Post, class Sink>
SSM_DEV void pair_helper_transform(const Tf &tf, F f, const bool want_cross, Post post, Sink sink, double *xs,
                                   const int bar_id, const int zero) {
    constexpr int TD = TriSize<D>::value, TE = TriSize<E>::value;
    pair_bar(bar_id);  // (1)
    {
        double m[D], L[TD];
#pragma unroll
        for (int r = 0; r < D; ++r) m[r] = xs[(SL::XA + r) * TPB];
#pragma unroll
        for (int a = 0; a < TD; ++a) L[a] = xs[(SL::XA + D + a) * TPB];
        post(m);
        double sp[E];
#pragma unroll
        for (int a = 0; a < E; ++a) sp[a] = 0.0;
#pragma unroll
        for (int q = 0; q < D; ++q) {  // "-" points
            double x[D], o[E];
            sigma_point<D, PTS_AXIS_C>(tf, D + 1 + q, m, L, x);
            f(x, o);
#pragma unroll
            for (int a = 0; a < E; ++a) {
                xs[(SL::XFH + a * D + q) * TPB] = o[a];
                sp[a] = fma(o[a], tf.wm[D + 1 + q], sp[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < E; ++a) xs[(SL::XMH + a) * TPB] = sp[a];
    }
    pair_bar(bar_id);  // (2)
    double Fh[E][D];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int q = 0; q < D; ++q) Fh[a][q] = xs[(SL::XFH + a * D + q) * TPB];
#pragma unroll 1
    for (int b = 0; b < E; ++b) {
        const Tf &tw = weights_in_loop(tf, b, zero);
        double fm[D + 1], fo[D];
#pragma unroll
        for (int j = 0; j <= D; ++j) fm[j] = xs[(SL::XFM + b * (D + 1) + j) * TPB];
#pragma unroll
        for (int q = 0; q < D; ++q) fo[q] = xs[(SL::XFH + b * D + q) * TPB];
        {
            double g[D];
#pragma unroll
            for (int q = 0; q < D; ++q) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= D; ++j) s = fma(fm[j], tw.W[D + 1 + q][j], s);
#pragma unroll
                for (int r = 0; r < D; ++r) s = fma(fo[r], tw.W[D + 1 + q][D + 1 + r], s);
                g[q] = s;
            }
#pragma unroll
            for (int a = 0; a < E; ++a) {
                if (a < b) continue;
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < D; ++q) s = fma(Fh[a][q], g[q], s);
                xs[(SL::XC + a * (a + 1) / 2 + b) * TPB] = s;
            }
        }
        if (KIND == SSM_TF_TP) {
            double g[D];
#pragma unroll
            for (int q = 0; q < D; ++q) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= D; ++j) s = fma(fm[j], tw.iK[(D + 1 + q) % Tf::NK][j], s);
#pragma unroll
                for (int r = 0; r < D; ++r) s = fma(fo[r], tw.iK[(D + 1 + q) % Tf::NK][D + 1 + r], s);
                g[q] = s;
            }
#pragma unroll
            for (int a = 0; a < E; ++a) {
                if (a < b || (!tf.tp_full && a != b)) continue;
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < D; ++q) s = fma(Fh[a][q], g[q], s);
                xs[(SL::XC + TE + a * (a + 1) / 2 + b) * TPB] = s;
            }
        }
        if (CROSS && want_cross) {
            double T[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= D; ++j) s = fma(fm[j], tw.Wcc[d][j], s);
#pragma unroll
                for (int r = 0; r < D; ++r) s = fma(fo[r], tw.Wcc[d][D + 1 + r], s);
                T[d] = s;
            }
            double crow[D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d <= r; ++d) s = fma(T[d], xs[(SL::XA + D + tri(r, d)) * TPB], s);  // Cholesky factor re-read
                crow[r] = s;
            }
            sink(b, crow);
        }
    }
    pair_bar(bar_id);  // (3)
}

#ifndef SSM_DBG_ROLE
#define SSM_DBG_ROLE 2
#endif
#ifndef SSM_PAIR_SYNC_STEPS
#define SSM_PAIR_SYNC_STEPS 1
#endif

template <class Dyn, class Obs, int KIND, class Par, int TPB, int MINB>
__global__ void __launch_bounds__(2 * TPB, MINB) filter_pair_kernel(const __grid_constant__ Par p) {
    constexpr int DX = Dyn::DX, DY = Obs::DY;
    constexpr int TX = TriSize<DX>::value, TY = TriSize<DY>::value;
    using SL = PairSlots<DX, DY, KIND>;
    extern __shared__ double ssm_pair_smem[];
    const FilterBuffers &b = p.b;
    const int N = b.n_steps;
    const long long ld = b.ld;
    const CompStride<false> cs((long long)N * ld);
    constexpr int NSTATE = DX + TX + 2;
    const int role = threadIdx.x / TPB;  // 0 = main, 1 = helper; uniform within a warp (TPB is a multiple of 32)
    const int lane = threadIdx.x - role * TPB;
    const int bar_id = 1 + (lane >> 5);
    double *xs = ssm_pair_smem + lane;
    __shared__ int s_ticket;
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
        const long long t_raw = blk * TPB + lane;
        const bool active = t_raw < b.n_traj;
        const long long t = active ? t_raw : b.n_traj - 1;
        if (kc > 0) {
            if (threadIdx.x == 0) {
                int spins = 0;
                while (atomicAdd(b.sched + 1 + blk, 0) < kc) { __nanosleep(200); ++spins; }
                if (spins) atomicAdd(b.sched + 1 + b.n_blocks, spins);
                __threadfence();
            }
            __syncthreads();
        }
        const double tbase = (double)(b.k0 + (b.t_offset ? b.t_offset[t] : 0));

        if (role == 0 && SSM_DBG_ROLE != 1) {
            // ======================================= main warp =======================================
            double m[DX], P[TX];
            int fail = active ? 0 : -1, kfail = 0;
            if (kc > 0) {
                const double *st = b.state + (blk * NSTATE) * TPB + lane;
#pragma unroll
                for (int a = 0; a < DX; ++a) m[a] = __ldcg(st + (long long)a * TPB);
#pragma unroll
                for (int a = 0; a < TX; ++a) P[a] = __ldcg(st + (long long)(DX + a) * TPB);
                fail = (int)__ldcg(st + (long long)(DX + TX) * TPB);
                kfail = (int)__ldcg(st + (long long)(DX + TX + 1) * TPB);
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
                for (int a = 0; a < TX; ++a) P[a] = p.P0[a];
            }
            if (kc == 0 && b.resume && active && b.status[t] != 0) { fail = b.status[t] & 0xff; kfail = b.k_lo; }
            double ynext[DY];
#pragma unroll
            for (int a = 0; a < DY; ++a) ynext[a] = ld_stream(b.y + cs(a) + ((long long)k_begin * ld + t));

            for (int k = k_begin; k < k_end; ++k) {
                if (SSM_PAIR_SYNC_STEPS) __syncthreads();
                const long long rk = (long long)k * ld + t;
                double yk[DY];
#pragma unroll
                for (int a = 0; a < DY; ++a) yk[a] = ynext[a];
                if (k + 1 < k_end) {
#pragma unroll
                    for (int a = 0; a < DY; ++a) ynext[a] = ld_stream(b.y + cs(a) + (rk + ld));
                }
                const double time = tbase + (double)k;
                // ---- time update (ssinf.py:276-279) ----
                double mp[DX], Pp[TX], dummy[DX][DX];
                bool ok = pair_main_transform<DX, DX, KIND, false, TPB, SL>(
                    p.tf_dyn, m, P,
                    [&](const double (&x)[DX], double (&o)[DX]) {
                        const double q0[Dyn::DQ] = {};
                        Dyn::template f<false>(p.dyn_par, x, q0, time, o);
                    },
                    mp, Pp, dummy, xs, bar_id, p.zero);
                if (!ok && !fail) { fail = SSM_FAIL_CHOL_DYN; kfail = k; }
#pragma unroll
                for (int a = 0; a < TX; ++a) Pp[a] += p.GQG[a];
#pragma unroll
                for (int a = 0; a < TX; ++a) xs[(SL::XP + a) * TPB] = Pp[a];  // the helper stores pr_cov
                // ---- predictive measurement moments (ssinf.py:287-291) ----
                double my[DY], Sy[TY], Syx[DY][DX];
                ok = pair_main_transform<DX, DY, KIND, true, TPB, SL>(
                    p.tf_obs, mp, Pp,
                    [&](const double (&x)[DX], double (&o)[DY]) {
                        const double r0[DY] = {};
                        Obs::template h<false>(p.obs_par, x, r0, time, o);
                    },
                    my, Sy, Syx, xs, bar_id, p.zero);
                if (!ok && !fail) { fail = SSM_FAIL_CHOL_OBS; kfail = k; }
#pragma unroll
                for (int a = 0; a < TY; ++a) Sy[a] += p.R[a];
                // ---- measurement update (ssinf.py:321-323) ----
                bool fin = true;
#pragma unroll
                for (int a = 0; a < TY; ++a) fin = fin && finite_d(Sy[a]);
#pragma unroll
                for (int a = 0; a < DY; ++a)
#pragma unroll
                    for (int d = 0; d < DX; ++d) fin = fin && finite_d(Syx[a][d]);
                if (!fin && !fail) { fail = SSM_FAIL_NONFINITE_GAIN; kfail = k; }
                double K[DX][DY], Ls[TY];
                ok = spd_gain<DY, DX>(Sy, Syx, K, Ls);
                if (!ok && !fail) { fail = SSM_FAIL_CHOL_GAIN; kfail = k; }
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
                    double KS[DX][DY];
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
                if (active) {
                    store_vec<DX>(b.fi_mean, cs, rk, m);
                    store_sym<DX>(b.fi_cov, cs, rk, P);
                }
            }
            if (k_end < b.k_hi) {
                double *st = b.state + (blk * NSTATE) * TPB + lane;
#pragma unroll
                for (int a = 0; a < DX; ++a) __stcg(st + (long long)a * TPB, m[a]);
#pragma unroll
                for (int a = 0; a < TX; ++a) __stcg(st + (long long)(DX + a) * TPB, P[a]);
                __stcg(st + (long long)(DX + TX) * TPB, (double)fail);
                __stcg(st + (long long)(DX + TX + 1) * TPB, (double)kfail);
            } else if (active) {
                if (fail) {
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
        } else if (SSM_DBG_ROLE != 0) {
            // ====================================== helper warp ======================================
            const bool want_xx = b.pr_xx != nullptr;
            for (int k = k_begin; k < k_end; ++k) {
                if (SSM_PAIR_SYNC_STEPS) __syncthreads();
                const long long rk = (long long)k * ld + t;
                const double time = tbase + (double)k;
                double *q_xx = want_xx ? row_ptr(b.pr_xx, rk) : nullptr;
                pair_helper_transform<DX, DX, KIND, true, TPB, SL>(
                    p.tf_dyn,
                    [&](const double (&x)[DX], double (&o)[DX]) {
                        const double q0[Dyn::DQ] = {};
                        Dyn::template f<false>(p.dyn_par, x, q0, time, o);
                    },
                    want_xx, [](const double (&)[DX]) {},
                    [&](int a, const double (&row)[DX]) {  // Cov(x_k, x_{k-1}) row a -> pr_xx_cov[a][:][k][t]
                        if (active) {
#pragma unroll
                            for (int c = 0; c < DX; ++c) st_stream(q_xx + cs(a * DX + c), row[c]);
                        }
                    },
                    xs, bar_id, p.zero);
                pair_helper_transform<DX, DY, KIND, false, TPB, SL>(
                    p.tf_obs,
                    [&](const double (&x)[DX], double (&o)[DY]) {
                        const double r0[DY] = {};
                        Obs::template h<false>(p.obs_par, x, r0, time, o);
                    },
                    true,
                    [&](const double (&mp)[DX]) {  // predictive moments of this step (ssinf.py:276-279)
                        if (active) {
                            store_vec<DX>(b.pr_mean, cs, rk, mp);
                            if (b.pr_cov) {
                                double Pp[TX];
#pragma unroll
                                for (int a = 0; a < TX; ++a) Pp[a] = xs[(SL::XP + a) * TPB];
                                store_sym<DX>(b.pr_cov, cs, rk, Pp);
                            }
                        }
                    },
                    [](int, const double (&)[DX]) {},  // Cov(h, x) is the main warp's
                    xs, bar_id, p.zero);
            }
        }
        if (k_end < b.k_hi) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(b.sched + 1 + blk, kc + 1);
        }
        if (!ticketed) break;
    }
}

template <int D, int E, int KIND>
inline void fill_tf_pair(TfPair<D, E, KIND> &o, const ssm_transform &tf, const HostTfInfo &info) {
    memset(&o, 0, sizeof(o));
    const int N = tf.n_pts;
    o.tp_full = tf.tp_full_matrix;
    o.c = info.c;
    o.tp_a = tf.nu - 2.0;
    o.tp_b = 1.0 / (tf.nu - 2.0 + (double)N);
    for (int i = 0; i < N; ++i) o.wm[i] = tf.wm[i];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) o.W[i][j] = tf.Wc[i * N + j];
    for (int d = 0; d < D; ++d)
        for (int i = 0; i < N; ++i) o.Wcc[d][i] = tf.Wcc[d * N + i];
    if (tf.kind == SSM_TF_BQ && tf.model_var)
        for (int a = 0; a < E; ++a)
            for (int b = 0; b < E; ++b) o.mv[a][b] = tf.model_var[a * E + b];
    if (tf.kind == SSM_TF_TP && tf.model_var) o.mv[0][0] = tf.model_var[0];
    if (KIND == SSM_TF_TP)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) o.iK[i % o.NK][j] = tf.iK[i * N + j];
}

// Launch; returns SSM_E_UNSUPPORTED when the device cannot hold one CTA (caller falls back to filter_kernel).
template <class Dyn, class Obs, int KIND, int TPB, int MINB>
int launch_filter_pair(const FilterLaunch &L, const HostTfInfo &id, const HostTfInfo &io) {
    constexpr int DX = Dyn::DX, DY = Obs::DY;
    using Par = PairPar<DX, DY, KIND>;
    using SL = PairSlots<DX, DY, KIND>;
    static_assert(sizeof(Par) <= 32000, "kernel parameter block too large");
    static_assert(TPB % 32 == 0 && TPB / 32 <= 15, "one named barrier per warp pair");
    const ssm_desc &d = *L.desc;
    Par *pp = new Par;
    Par &p = *pp;
    memset(pp, 0, sizeof(Par));
    fill_tf_pair(p.tf_dyn, d.tf_dyn, id);
    fill_tf_pair(p.tf_obs, d.tf_obs, io);
    for (int i = 0; i < 4; ++i) p.dyn_par[i] = d.dyn_par[i];
    for (int i = 0; i < 8; ++i) p.obs_par[i] = d.obs_par[i];
    for (int i = 0; i < DX; ++i) p.m0[i] = d.m0[i];
    pack_lower<DX>(d.P0, p.P0);
    pack_lower<DX>(d.GQG, p.GQG);
    pack_lower<DY>(d.R, p.R);
    p.b = L.buf;
    constexpr int THREADS = 2 * TPB;
    const long long blocks = (L.buf.n_traj + TPB - 1) / TPB;
    auto kern = filter_pair_kernel<Dyn, Obs, KIND, Par, TPB, MINB>;
    const size_t smem = sizeof(double) * SL::TOTAL * TPB;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        delete pp;
        return SSM_E_UNSUPPORTED;
    }
    long long grid = blocks;
    void *work = nullptr;
    int occ = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
    if (occ < 1) { delete pp; return SSM_E_UNSUPPORTED; }
    const long long cap = (long long)occ * sms;
    const char *env_on = getenv("SSM_TICKET"), *env_chunk = getenv("SSM_TICKET_CHUNK");
    const int win = L.buf.k_hi - L.buf.k_lo;
    int CHUNK = env_chunk ? atoi(env_chunk) : SSM_TICKET_CHUNK;
    if (!env_chunk && win < 8 * CHUNK) {
        long long best = -1;
        for (int c : {25, 20, 16, 12, 10, 8}) {
            const long long items = blocks * ((win + c - 1) / c);
            const long long cost = ((items + cap - 1) / cap) * c;
            if (best < 0 || cost < best) { best = cost; CHUNK = c; }
        }
    }
    const bool want_ticket = SSM_TICKET_SCHED && !(env_on && atoi(env_on) == 0);
    if (want_ticket && CHUNK > 0 && blocks > cap && win >= 2 * CHUNK) {
        const size_t n_int = ((size_t)blocks + 2 + 1) / 2 * 2;
        const size_t bytes = n_int * sizeof(int) + (size_t)blocks * TPB * (DX + TriSize<DX>::value + 2) * sizeof(double);
        if (scratch_alloc((void **)&work, bytes, L.stream) != cudaSuccess) { delete pp; set_error("cudaMallocAsync failed"); return SSM_E_CUDA; }
        cudaMemsetAsync(work, 0, n_int * sizeof(int), L.stream);
        p.b.sched = (int *)work;
        p.b.state = (double *)((int *)work + n_int);
        p.b.chunk = CHUNK;
        p.b.n_blocks = (int)blocks;
        grid = cap;
    }
    kern<<<(unsigned)grid, THREADS, smem, L.stream>>>(p);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && filter_nan_fill(L.buf, DX, L.stream) != SSM_OK) err = cudaErrorUnknown;
    if (work && getenv("SSM_TICKET_DEBUG")) {
        int waits = 0;
        cudaMemcpyAsync(&waits, (int *)work + 1 + blocks, sizeof(int), cudaMemcpyDeviceToHost, L.stream);
        cudaStreamSynchronize(L.stream);
        fprintf(stderr, "[ssm pair ticket] blocks=%lld grid=%lld chunk=%d occ=%d smem=%zu polls that waited: %d\n", blocks, grid, CHUNK, occ, smem, waits);
    }
    if (work) cudaFreeAsync(work, L.stream);
    delete pp;
    return err == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

}  // namespace ssm
