// K3 (sm_100a path): RTS smoother with its HBM traffic moved by the TMA unit.
//
// The smoother is pure streaming: per trajectory-step it reads 65 (+5 truth) doubles and writes 30, with ~900
// FLOP in between.  The first version (smoother_kernel, ssm_smoother.cu) let every thread issue its own ld/st:
// a third of its instructions were 64-bit address arithmetic, it needed 255 registers (8 warps/SM) and ran at
// 52 % of the HBM peak, latency-bound.  Here one warp owns 32 trajectories and
//   * every input row of a step ([component][k][t0 .. t0+32) = 256 contiguous bytes) is fetched by ONE
//     cp.async.bulk global->shared instruction (lanes issue different rows), completion counted on an mbarrier;
//   * the warp copies the landed rows into registers, immediately re-arms the barrier and issues the loads of the
//     NEXT step into the same buffer, and only then does the arithmetic: a whole step of math hides the latency;
//   * results are staged in shared memory and leave through cp.async.bulk shared->global (the symmetric mirror of
//     the covariance is written from the same staged row), so no thread computes a global address in the loop;
//   * the per-step error statistics are reduced through shared memory (24 adds + 6 shuffles per thread instead of
//     115 shuffles).
// ~29 KB of shared memory per warp-CTA -> 7 resident CTAs per SM, each with a full step of loads in flight.
// Arithmetic and its order are those of smoother_kernel: smoothed moments are bitwise identical.
#pragma once
#include "ssm_scores.cuh"

namespace ssm {
namespace tma {

SSM_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SSM_DEV void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SSM_DEV void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SSM_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
SSM_DEV void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
SSM_DEV void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
SSM_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SSM_DEV void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SSM_DEV void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
SSM_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tma

struct SmootherArgs {
    const double *fi_mean, *fi_cov, *pr_mean, *pr_cov, *pr_xx;
    double *sm_mean, *sm_cov;
    int32_t *status;
    const double *x_truth;
    double *partial, *rmse_acc;
    long long ld;
    int N, k_lo, k_hi;
    double *quad;   // (N, ld) d' P_s^-1 d per (step, trajectory), nullable (per-thread kernel only)
};

template <int DX, bool SCORE>
struct SmootherTmaLayout {
    static constexpr int TX = TriSize<DX>::value;
    static constexpr int NA = DX + TX + (SCORE ? DX : 0);  // rows of step k:   fi_mean, tril(fi_cov), x_truth
    static constexpr int NB = DX + TX + DX * DX;           // rows of step k+1: pr_mean, tril(pr_cov), pr_xx_cov
    static constexpr int NIN = NA + NB;
    static constexpr int NOUT_S = DX + TX;                 // staged rows: sm_mean, tril(sm_cov)
    static constexpr int NOUT_G = DX + DX * DX;            // rows written to global memory (<= 32: one per lane)
    static constexpr int WP = ScoreRow<DX>::WP;
    static constexpr int RPL = (NIN + 31) / 32;            // input rows per lane
    static constexpr int ROW_BYTES = 32 * (int)sizeof(double);
    static constexpr size_t SMEM = (size_t)ROW_BYTES * (NIN + NOUT_S + (SCORE ? WP : 0)) + 16;
    static_assert(NOUT_G <= 32, "one output row per lane");
};

SSM_DEV void tri_decode(int q, int &r, int &c) {
    r = 0;
    while ((r + 1) * (r + 2) / 2 <= q) ++r;
    c = q - r * (r + 1) / 2;
}

template <int DX, bool SCORE>
__global__ void __launch_bounds__(32) smoother_tma_kernel(const SmootherArgs a) {
    using Lay = SmootherTmaLayout<DX, SCORE>;
    constexpr int TX = Lay::TX, NA = Lay::NA, NIN = Lay::NIN, WP = Lay::WP, RPL = Lay::RPL, RB = Lay::ROW_BYTES;
    extern __shared__ __align__(128) unsigned char ssm_smoother_smem[];
    double *sin = reinterpret_cast<double *>(ssm_smoother_smem);  // [NIN][32]
    double *sout = sin + NIN * 32;                                // [NOUT_S][32]
    double *ssc = sout + Lay::NOUT_S * 32;                        // [WP][32] (SCORE)
    uint64_t *bar = reinterpret_cast<uint64_t *>(ssc + (SCORE ? WP * 32 : 0));
    const int lane = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * 32, t = t0 + lane;  // full blocks only (host splits off the tail)
    const int N = a.N, k_lo = a.k_lo, k_hi = a.k_hi, WLEN = k_hi - k_lo;
    const long long ld = a.ld, cs = (long long)N * ld;               // step stride, component stride

    // ---- per-lane row tables: which global row this lane fetches / writes back -------------------------------
    const double *in_src[RPL];
    int in_kofs[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        const int row = lane + 32 * j;
        const double *base = nullptr;
        int comp = 0, kofs = 0, r, c;
        if (row < DX) { base = a.fi_mean; comp = row; }
        else if (row < DX + TX) { tri_decode(row - DX, r, c); base = a.fi_cov; comp = r * DX + c; }
        else if (row < NA) { base = a.x_truth; comp = row - DX - TX; }
        else if (row < NA + DX) { base = a.pr_mean; comp = row - NA; kofs = 1; }
        else if (row < NA + DX + TX) { tri_decode(row - NA - DX, r, c); base = a.pr_cov; comp = r * DX + c; kofs = 1; }
        else if (row < NIN) { base = a.pr_xx; comp = row - NA - DX - TX; kofs = 1; }
        in_src[j] = base ? base + comp * cs + t0 : nullptr;
        in_kofs[j] = kofs;
    }
    double *out_dst = nullptr;
    int out_row = 0;
    if (lane < DX) { out_dst = a.sm_mean + lane * cs + t0; out_row = lane; }
    else if (lane < Lay::NOUT_G) { const int comp = lane - DX; out_dst = a.sm_cov + comp * cs + t0; out_row = DX + sym(comp / DX, comp % DX); }

    auto issue_loads = [&](int k) {
        const int nrows = (k < N - 2) ? NIN : NA;  // slots N-1, N-2 keep their filtered values: no k+1 rows needed
        if (lane == 0) tma::mbar_arrive_expect_tx(bar, (uint32_t)(nrows * RB));
        __syncwarp();
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            const int row = lane + 32 * j;
            if (row < nrows) tma::bulk_g2s(sin + row * 32, in_src[j] + (long long)(k + in_kofs[j]) * ld, RB, bar);
        }
    };

    if (lane == 0) tma::mbar_init(bar, 1);
    tma::fence_proxy_async();
    __syncwarp();
    issue_loads(k_hi - 1);

    bool alive = a.status[t] == 0;
    double se_acc[DX];
#pragma unroll
    for (int i = 0; i < DX; ++i) se_acc[i] = (SCORE && a.rmse_acc && k_hi < N) ? a.rmse_acc[(long long)i * ld + t] : 0.0;
    double ms[DX], Ps[TX];
    if (k_hi < N && alive) {
        const int ki = (k_hi >= N - 2) ? N - 1 : k_hi;  // the recursion starts from slot N-1 (ssinf.py:117, 137)
#pragma unroll
        for (int i = 0; i < DX; ++i) ms[i] = ld_stream(a.sm_mean + (i * cs + (long long)ki * ld + t));
#pragma unroll
        for (int r = 0; r < DX; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) Ps[tri(r, c)] = ld_stream(a.sm_cov + ((r * DX + c) * cs + (long long)ki * ld + t));
    }
    int fail = 0, kfail = 0;
    uint32_t phase = 0;
    for (int k = k_hi - 1; k >= k_lo; --k) {
        tma::mbar_wait(bar, phase);
        phase ^= 1;
        const bool rec = k < N - 2;  // a recursion step (uniform)
        double mf[DX], Pf[TX], xt[DX], mp[DX], Pp[TX], Pxx[DX][DX];
#pragma unroll
        for (int i = 0; i < DX; ++i) mf[i] = sin[i * 32 + lane];
#pragma unroll
        for (int i = 0; i < TX; ++i) Pf[i] = sin[(DX + i) * 32 + lane];
        if (SCORE) {
#pragma unroll
            for (int i = 0; i < DX; ++i) xt[i] = sin[(DX + TX + i) * 32 + lane];
        }
        if (rec) {
#pragma unroll
            for (int i = 0; i < DX; ++i) mp[i] = sin[(NA + i) * 32 + lane];
#pragma unroll
            for (int i = 0; i < TX; ++i) Pp[i] = sin[(NA + DX + i) * 32 + lane];
#pragma unroll
            for (int r = 0; r < DX; ++r)
#pragma unroll
                for (int c = 0; c < DX; ++c) Pxx[r][c] = sin[(NA + DX + TX + r * DX + c) * 32 + lane];
        }
        __syncwarp();
        if (k - 1 >= k_lo) {  // the buffer is free again: fetch the next (earlier) step while this one is computed
            tma::fence_proxy_async();
            issue_loads(k - 1);
        }

        double om[DX], oP[TX];  // what goes to sm_mean / sm_cov at slot k
        bool live = alive;
        if (!rec) {
            if (k == N - 1 && alive) {
#pragma unroll
                for (int i = 0; i < DX; ++i) ms[i] = mf[i];
#pragma unroll
                for (int i = 0; i < TX; ++i) Ps[i] = Pf[i];
            }
#pragma unroll
            for (int i = 0; i < DX; ++i) om[i] = mf[i];
#pragma unroll
            for (int i = 0; i < TX; ++i) oP[i] = Pf[i];
        } else {
          do {
            if (!alive) break;
            // scipy's cho_factor / cho_solve reject non-finite input (ValueError)        ssinf.py:342
            bool fin = true;
#pragma unroll
            for (int i = 0; i < TX; ++i) fin = fin && finite_d(Pp[i]);
#pragma unroll
            for (int r = 0; r < DX; ++r)
#pragma unroll
                for (int c = 0; c < DX; ++c) fin = fin && finite_d(Pxx[r][c]);
            if (!fin) { fail = SSM_FAIL_NONFINITE_GAIN; kfail = k; alive = false; break; }
            // D = (Pp^-1 Pxx)^T                                                       ssinf.py:342
            double Dg[DX][DX], Ls[TX];
            if (!spd_gain<DX, DX>(Pp, Pxx, Dg, Ls)) { fail = SSM_FAIL_CHOL_SMOOTH; kfail = k; alive = false; break; }
            // m_s = m_f + D (m_s+ - m_p)                                              ssinf.py:343
            double dm[DX];
#pragma unroll
            for (int i = 0; i < DX; ++i) dm[i] = ms[i] - mp[i];
#pragma unroll
            for (int i = 0; i < DX; ++i) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < DX; ++c) s = fma(Dg[i][c], dm[c], s);
                ms[i] = mf[i] + s;
            }
            // P_s = P_f + D (P_s+ - P_p) D^T                                          ssinf.py:344
            double dP[TX], T[DX][DX];
#pragma unroll
            for (int i = 0; i < TX; ++i) dP[i] = Ps[i] - Pp[i];
#pragma unroll
            for (int i = 0; i < DX; ++i)
#pragma unroll
                for (int c = 0; c < DX; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int e = 0; e < DX; ++e) s = fma(Dg[i][e], dP[sym(e, c)], s);
                    T[i][c] = s;
                }
#pragma unroll
            for (int r = 0; r < DX; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int e = 0; e < DX; ++e) s = fma(T[r][e], Dg[c][e], s);
                    Ps[tri(r, c)] = Pf[tri(r, c)] + s;
                }
          } while (0);
            live = alive;
#pragma unroll
            for (int i = 0; i < DX; ++i) om[i] = ms[i];
#pragma unroll
            for (int i = 0; i < TX; ++i) oP[i] = Ps[i];
        }
        // ---- stage the outputs; failed trajectories are NaN from the failing slot down ------------------------
        tma::bulk_wait_read0();  // this lane's previous store has finished reading the staging rows
        __syncwarp();
#pragma unroll
        for (int i = 0; i < DX; ++i) sout[i * 32 + lane] = live ? om[i] : qnan();
#pragma unroll
        for (int i = 0; i < TX; ++i) sout[(DX + i) * 32 + lane] = live ? oP[i] : qnan();
        tma::fence_proxy_async();
        __syncwarp();
        if (out_dst) tma::bulk_s2g(out_dst + (long long)k * ld, sout + out_row * 32, RB);
        tma::bulk_commit();

        if (SCORE) {
            double v[WP];
#pragma unroll
            for (int i = 0; i < WP; ++i) v[i] = 0.0;
            if (live) {
                double d[DX], se[DX];
#pragma unroll
                for (int i = 0; i < DX; ++i) d[i] = xt[i] - om[i];
                score_step<DX>(d, oP, v, se);
#pragma unroll
                for (int i = 0; i < DX; ++i) se_acc[i] += se[i];
            }
            // warp sum of the WP statistics through shared memory: lane = (row group, column segment); fixed order
#pragma unroll
            for (int i = 0; i < WP; ++i) ssc[i * 32 + lane] = v[i];
            __syncwarp();
            const int seg = lane & 3, rb = lane >> 2;
            double *prow = a.partial + ((long long)blockIdx.x * WLEN + (k - k_lo)) * WP;
#pragma unroll
            for (int m = 0; m < (WP + 7) / 8; ++m) {
                const int row = rb + 8 * m;
                double s = 0.0;
                if (row < WP) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) s += ssc[row * 32 + seg * 8 + ((i + rb) & 7)];
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (seg == 0 && row < WP) prow[row] = s;
            }
            __syncwarp();
        }
    }
    tma::bulk_wait0();
    if (SCORE && a.rmse_acc) {
#pragma unroll
        for (int i = 0; i < DX; ++i) a.rmse_acc[(long long)i * ld + t] = (alive && !fail) ? se_acc[i] : qnan();
    }
    if (fail) a.status[t] = ((kfail + 1) << 8) | fail;
}

// Can the TMA path take this problem?  cp.async.bulk needs 16-byte aligned addresses and sizes: every row
// ((c * N + k) * ld + t0) * 8 with t0 a multiple of 32 is aligned iff the bases are and ld is even.
inline bool smoother_tma_eligible(const SmootherArgs &a, long long n_traj) {
    auto al = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    // Opt-in (SSM_SMOOTH_TMA=1).  Measured on B200 at 125 000 x 500: 11.6 ms (20.9 ms with scores) against 9.6 ms (17.4 ms)
    // of the per-thread ld/st kernel: 65 bulk copies of 256 B per warp-step and 7 single-warp CTAs per SM leave the
    // FP64 dependency chains of a step exposed; kept as the measured alternative, not the default.
    const char *env = getenv("SSM_SMOOTH_TMA");
    if (!env || atoi(env) == 0) return false;
    return n_traj >= 32 && (a.ld % 2 == 0) && al(a.fi_mean) && al(a.fi_cov) && al(a.pr_mean) && al(a.pr_cov) && al(a.pr_xx) &&
           al(a.sm_mean) && al(a.sm_cov) && al(a.x_truth);
}

}  // namespace ssm
