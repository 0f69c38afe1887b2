// K6: error statistics over the trajectory axis, per time step.
// Replaces utils.squared_error / mse_matrix / neg_log_likelihood / log_cred_ratio (utils.py:18-148)
// and the Python reduction loops of research/gpq/icinco_demo.py:17-52 and
// research/bsq/bsq_tracking.py:311-337.
//
// One thread per trajectory walks the time axis; per step the CTA reduces its 128 trajectories with
// warp shuffles + one shared-memory pass (fixed order), writes one partial row per (CTA, step), and a
// second kernel sums the partial rows in CTA order: the result is bitwise reproducible for a given
// trajectory count, and the packed rows are what the multi-GPU driver all-reduces (NCCL).
// The log credibility ratio needs the GLOBAL per-step MSE matrix first (two-phase reduction).
#include "ssm_scores.cuh"

#ifndef SSM_P2_GROUP
#define SSM_P2_GROUP 4   // measured on 125 000 x 500: 1 / 2 / 4 / 8 steps per reduction 1.37 / 1.26 / 1.12 / 2.01 ms (8: 33 KB of shared memory per CTA)
#endif
namespace ssm {

void set_error(const char *fmt, ...);

template <int DX>
__global__ void __launch_bounds__(SC_THREADS) scores_phase1_kernel(const double *__restrict__ x, const double *__restrict__ mean,
                                                                   const double *__restrict__ cov, const int32_t *__restrict__ status,
                                                                   double *__restrict__ partial, double *__restrict__ rmse_acc,
                                                                   double *__restrict__ nll_acc, double *__restrict__ quad,
                                                                   long long n_traj, int N, int k_lo, int k_hi, long long ld) {
    constexpr int TX = TriSize<DX>::value, W = ScoreRow<DX>::WP;
    const int WLEN = k_hi - k_lo;
    __shared__ double smem[BlockReduce<W>::SIZE];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n_traj && (status == nullptr || status[t] == 0);
    const long long cs = (long long)N * ld;  // component stride (row_ptr, ssm_common.cuh)
    double se_acc[DX];
#pragma unroll
    for (int a = 0; a < DX; ++a) se_acc[a] = (rmse_acc && k_lo > 0 && t < n_traj) ? rmse_acc[(long long)a * ld + t] : 0.0;
    // per-trajectory time-sum of the NLL (the nllData of research/gpq/icinco_demo.py:28-48 before its time mean)
    double nll_sum = (nll_acc && k_lo > 0 && t < n_traj) ? nll_acc[t] : 0.0;
    for (int k = k_lo; k < k_hi; ++k) {
        double v[W];
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = 0.0;
        if (live) {
            double d[DX], P[TX], se[DX];
            const long long rk = (long long)k * ld + t;
            const double *qx = row_ptr(x, rk), *qm = row_ptr(mean, rk), *qc = row_ptr(cov, rk);
#pragma unroll
            for (int a = 0; a < DX; ++a) d[a] = ld_stream(qx + a * cs) - ld_stream(qm + a * cs);
#pragma unroll
            for (int r = 0; r < DX; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) P[tri(r, c)] = ld_stream(qc + (r * DX + c) * cs);
            double qf;
            score_step<DX>(d, P, v, se, &qf);
            if (quad) st_stream(quad + rk, qf);   // d' P^-1 d for ssm_scores_phase2_quad
#pragma unroll
            for (int a = 0; a < DX; ++a) se_acc[a] += se[a];
            nll_sum += v[DX + TX];
        }
        block_reduce_store<W>(v, smem, k, partial + ((long long)blockIdx.x * WLEN + (k - k_lo)) * W);
    }
    if (rmse_acc && t < n_traj) {
#pragma unroll
        for (int a = 0; a < DX; ++a) rmse_acc[(long long)a * ld + t] = live ? se_acc[a] : qnan();
    }
    if (nll_acc && t < n_traj) nll_acc[t] = live ? nll_sum : qnan();
}

// Both finalise kernels: blockDim = (32, FIN_GROUPS); thread (x, y) adds the partial rows y, y + FIN_GROUPS, ... of
// element x in ascending order, then the groups are combined in ascending y: a fixed order for a given row count, and
// FIN_GROUPS times the memory-level parallelism of one thread per element (the sums run over ~1000 rows of 90 MB).
SSM_DEV double finalize_sum(const double *__restrict__ p, long long stride, int n_rows, double (*sm)[33]) {
    double s = 0.0;
    for (int c = threadIdx.y; c < n_rows; c += FIN_GROUPS) s += p[(long long)c * stride];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.y == 0)
        for (int g = 0; g < FIN_GROUPS; ++g) tot += sm[g][threadIdx.x];
    return tot;
}

__global__ void scores_finalize_kernel(const double *__restrict__ partial, double *__restrict__ stats, int n_cta, long long row) {
    __shared__ double sm[FIN_GROUPS][33];
    const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
    const long long ic = i < row ? i : row - 1;
    const double s = finalize_sum(partial + ic, row, n_cta, sm);
    if (threadIdx.y == 0 && i < row) stats[i] = s;
}

__global__ void scores_finalize_packed_kernel(const double *__restrict__ partial, double *__restrict__ stats, int n_cta, int n_steps, int dx) {
    __shared__ double sm[FIN_GROUPS][33];
    const int tx = dx * (dx + 1) / 2, wp = dx + tx + 3, w = dx + dx * dx + 3;
    const long long n = (long long)n_steps * wp;
    const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
    const long long ic = i < n ? i : n - 1;
    const double s = finalize_sum(partial + ic, n, n_cta, sm);   // partial[(c * n_steps + k) * wp + j] = partial[c * n + i]
    if (threadIdx.y != 0 || i >= n) return;
    const int k = (int)(i / wp), j = (int)(i % wp);
    double *row = stats + (long long)k * w;
    if (j < dx) row[j] = s;
    else if (j < dx + tx) {
        int r = 0, q = j - dx;
        while ((r + 1) * (r + 2) / 2 <= q) ++r;
        const int cc = q - r * (r + 1) / 2;
        row[dx + r * dx + cc] = s;
        row[dx + cc * dx + r] = s;
    } else row[dx + dx * dx + (j - dx - tx)] = s;
}

// Cholesky factor of the per-step MSE matrix, once per step instead of once per trajectory-step: row k of `tab` =
// [ packed lower factor (TX) | reciprocal pivots (DX) | ok ]
template <int DX>
__global__ void phase2_prepare_kernel(const double *__restrict__ mse, double *__restrict__ tab, int N, int k_lo, int k_hi) {
    constexpr int TX = TriSize<DX>::value, TW = TX + DX + 1;
    const int k = k_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    double S[TX], Ls[TX], inv[DX];
#pragma unroll
    for (int r = 0; r < DX; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) S[tri(r, c)] = mse[(long long)(r * DX + c) * N + k];
    const bool ok = chol_lower<DX>(S, Ls, inv);
    double *row = tab + (long long)(k - k_lo) * TW;
#pragma unroll
    for (int i = 0; i < TX; ++i) row[i] = Ls[i];
#pragma unroll
    for (int i = 0; i < DX; ++i) row[TX + i] = inv[i];
    row[TX + DX] = ok ? 1.0 : 0.0;
}

// QUAD: `cov` is not the covariance array but the (N, ld) array of d' P^-1 d that the smoother's in-kernel scoring
// stored (ssm_smooth_quad): no covariance read (120 of 200 bytes per unit for dx = 5), no factorisation.
// RES (with QUAD): `x` is the array of errors d = x - m the score-only smoother stored (ssm_smooth_scores), `mean` is not
// read: 48 instead of 88 bytes per unit for dx = 5.
template <int DX, bool QUAD, bool RES = false>
__global__ void __launch_bounds__(SC_THREADS) scores_phase2_kernel(const double *__restrict__ x, const double *__restrict__ mean,
                                                                   const double *__restrict__ cov, const int32_t *__restrict__ status,
                                                                   const double *__restrict__ tab, double *__restrict__ partial,
                                                                   double *__restrict__ lcr_acc,
                                                                   long long n_traj, int N, int k_lo, int k_hi, long long ld) {
    constexpr int TX = TriSize<DX>::value, TW = TX + DX + 1;
    // G time steps per CTA reduction.  The stored-error form reads 6 doubles per step and has next to no arithmetic: one
    // step at a time it is a chain of (6 loads -> wait -> reduce -> barrier); four steps at a time 24 loads are in flight
    // per thread and there is one barrier per four steps.  Same per-step sums in the same order: bitwise equal results.
    constexpr int G = (QUAD && RES) ? SSM_P2_GROUP : 1;
    const int WLEN = k_hi - k_lo;
    __shared__ double smem[BlockReduce<2 * G>::SIZE];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n_traj && (status == nullptr || status[t] == 0);
    const long long cs = (long long)N * ld;
    // per-trajectory time-sum of the log credibility ratio (nciData of research/gpq/icinco_demo.py:36-47)
    double lcr_sum = (lcr_acc && k_lo > 0 && t < n_traj) ? lcr_acc[t] : 0.0;
    for (int kg = k_lo; kg < k_hi; kg += G) {
        double v[2 * G];
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) v[i] = 0.0;
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const int k = kg + gi;
        if (live && k < k_hi) {
            double d[DX], P[TX], L[TX], inv[DX];
            const long long rk = (long long)k * ld + t;
            const double *qx = row_ptr(x, rk), *qm = RES ? nullptr : row_ptr(mean, rk), *qc = row_ptr(cov, rk);
#pragma unroll
            for (int a = 0; a < DX; ++a) d[a] = RES ? ld_stream(qx + a * cs) : ld_stream(qx + a * cs) - ld_stream(qm + a * cs);
            double q_given = 0.0;
            if (QUAD) q_given = ld_stream(qc);
            else {
#pragma unroll
                for (int r = 0; r < DX; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) P[tri(r, c)] = ld_stream(qc + (r * DX + c) * cs);
            }
            // log_cred_ratio, utils.py:113-120: both quadratic forms through Cholesky factors; the factor of the MSE
            // matrix comes from the per-step table (one address per warp: broadcast loads), substitutions multiply by
            // the reciprocal pivots the factorisations already have
            const double *ts = tab + (long long)(k - k_lo) * TW;
            bool ok = __ldg(ts + TX + DX) != 0.0;
            if (!QUAD) ok = chol_lower<DX>(P, L, inv) & ok;
            double qa = 0.0, qb = 0.0, za[DX], zb[DX];
#pragma unroll
            for (int i = 0; i < DX; ++i) {
                double sa = d[i], sb = d[i];
#pragma unroll
                for (int c = 0; c < i; ++c) {
                    if (!QUAD) sa = fma(-L[tri(i, c)], za[c], sa);
                    sb = fma(-__ldg(ts + tri(i, c)), zb[c], sb);
                }
                if (!QUAD) {
                    za[i] = sa * inv[i];
                    qa = fma(za[i], za[i], qa);
                }
                zb[i] = sb * __ldg(ts + TX + i);
                qb = fma(zb[i], zb[i], qb);
            }
            if (QUAD) qa = q_given;   // NaN where the covariance was not positive definite
            const double g = ok ? 10.0 * (log10(qa) - log10(qb)) : qnan();
            v[2 * gi] = g;
            v[2 * gi + 1] = fabs(g);
            lcr_sum += g;
        }
      }
        block_reduce_store<2 * G>(v, smem, (kg - k_lo) / G, partial + ((long long)blockIdx.x * WLEN + (kg - k_lo)) * 2, 2 * min(G, k_hi - kg));
    }
    if (lcr_acc && t < n_traj) lcr_acc[t] = live ? lcr_sum : qnan();
}

template <int DX>
static int run_phase1(const double *x, const double *mean, const double *cov, const int32_t *status, double *stats,
                      double *rmse_acc, double *nll_acc, double *quad, long long n_traj, int N, int k_lo, int k_hi, long long ld, cudaStream_t s) {
    constexpr int W = ScoreRow<DX>::WP;
    const int n_cta = (int)((n_traj + SC_THREADS - 1) / SC_THREADS);
    const int WLEN = k_hi - k_lo;
    double *partial = nullptr;
    if (scratch_alloc((void **)&partial, (size_t)n_cta * WLEN * W * sizeof(double), s) != cudaSuccess) return SSM_E_CUDA;
    scores_phase1_kernel<DX><<<n_cta, SC_THREADS, 0, s>>>(x, mean, cov, status, partial, rmse_acc, nll_acc, quad, n_traj, N, k_lo, k_hi, ld);
    const long long row = (long long)WLEN * W;
    scores_finalize_packed_kernel<<<(unsigned)((row + 31) / 32), dim3(32, FIN_GROUPS), 0, s>>>(partial, stats + (long long)k_lo * ScoreRow<DX>::W, n_cta, WLEN, DX);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(partial, s);
    return e == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

template <int DX, bool QUAD, bool RES = false>
static int run_phase2(const double *x, const double *mean, const double *cov, const int32_t *status, const double *mse,
                      double *lcr, double *lcr_acc, long long n_traj, int N, int k_lo, int k_hi, long long ld, cudaStream_t s) {
    constexpr int TW = TriSize<DX>::value + DX + 1;
    const int n_cta = (int)((n_traj + SC_THREADS - 1) / SC_THREADS);
    const int WLEN = k_hi - k_lo;
    double *partial = nullptr;  // [n_cta][WLEN][2] partial rows, then the [WLEN][TW] table of MSE factors
    if (scratch_alloc((void **)&partial, ((size_t)n_cta * WLEN * 2 + (size_t)WLEN * TW) * sizeof(double), s) != cudaSuccess) return SSM_E_CUDA;
    double *tab = partial + (size_t)n_cta * WLEN * 2;
    phase2_prepare_kernel<DX><<<(WLEN + 63) / 64, 64, 0, s>>>(mse, tab, N, k_lo, k_hi);
    scores_phase2_kernel<DX, QUAD, RES><<<n_cta, SC_THREADS, 0, s>>>(x, mean, cov, status, tab, partial, lcr_acc, n_traj, N, k_lo, k_hi, ld);
    const long long row = (long long)WLEN * 2;
    scores_finalize_kernel<<<(unsigned)((row + 31) / 32), dim3(32, FIN_GROUPS), 0, s>>>(partial, lcr + (long long)k_lo * 2, n_cta, row);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(partial, s);
    return e == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

}  // namespace ssm

using namespace ssm;

extern "C" int32_t ssm_scores_width(int32_t dx) { return dx + dx * dx + 3; }

extern "C" int ssm_scores_phase1_quad(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                      double *stats, double *rmse_acc, double *nll_acc, double *quad, int64_t n_traj, int32_t n_steps,
                                      int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    if (!x || !mean || !cov || !stats) { set_error("ssm_scores_phase1: NULL buffer"); return SSM_E_INVALID; }
    if (n_traj <= 0 || n_steps <= 0 || ld < n_traj) { set_error("ssm_scores_phase1: bad sizes"); return SSM_E_INVALID; }
    if (k_lo < 0 || k_hi <= k_lo || k_hi > n_steps) { set_error("ssm_scores_phase1: bad time window [%d, %d) of %d steps", k_lo, k_hi, n_steps); return SSM_E_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    switch (dx) {
        case 1: rc = run_phase1<1>(x, mean, cov, status, stats, rmse_acc, nll_acc, quad, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 2: rc = run_phase1<2>(x, mean, cov, status, stats, rmse_acc, nll_acc, quad, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 3: rc = run_phase1<3>(x, mean, cov, status, stats, rmse_acc, nll_acc, quad, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 4: rc = run_phase1<4>(x, mean, cov, status, stats, rmse_acc, nll_acc, quad, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 5: rc = run_phase1<5>(x, mean, cov, status, stats, rmse_acc, nll_acc, quad, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        default: set_error("ssm_scores: state dimension %d has no device implementation (1 .. 5)", dx); return SSM_E_UNSUPPORTED;
    }
    if (rc == SSM_E_CUDA) set_error("ssm_scores_phase1: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

extern "C" int ssm_scores_phase1_traj(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                      double *stats, double *rmse_acc, double *nll_acc, int64_t n_traj, int32_t n_steps,
                                      int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    return ssm_scores_phase1_quad(dx, x, mean, cov, status, stats, rmse_acc, nll_acc, nullptr, n_traj, n_steps, k_lo, k_hi, ld, stream);
}

extern "C" int ssm_scores_phase1_window(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                        double *stats, double *rmse_acc, int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi,
                                        int64_t ld, void *stream) {
    return ssm_scores_phase1_traj(dx, x, mean, cov, status, stats, rmse_acc, nullptr, n_traj, n_steps, k_lo, k_hi, ld, stream);
}

extern "C" int ssm_scores_phase1(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                 double *stats, double *rmse_acc, int64_t n_traj, int32_t n_steps, int64_t ld, void *stream) {
    return ssm_scores_phase1_window(dx, x, mean, cov, status, stats, rmse_acc, n_traj, n_steps, 0, n_steps, ld, stream);
}

// lcr: (n_steps, 2) = per step [ sum of log credibility ratios | sum of their absolute values ]
static int scores_phase2_impl(bool quad, int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                              const double *mse, double *lcr, double *lcr_acc, int64_t n_traj, int32_t n_steps,
                              int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    if (!x || !mean || !cov || !mse || !lcr) { set_error("ssm_scores_phase2: NULL buffer"); return SSM_E_INVALID; }
    if (n_traj <= 0 || n_steps <= 0 || ld < n_traj) { set_error("ssm_scores_phase2: bad sizes"); return SSM_E_INVALID; }
    if (k_lo < 0 || k_hi <= k_lo || k_hi > n_steps) { set_error("ssm_scores_phase2: bad time window [%d, %d) of %d steps", k_lo, k_hi, n_steps); return SSM_E_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    switch (dx) {
        case 1: rc = quad ? run_phase2<1, true>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s)
                        : run_phase2<1, false>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 2: rc = quad ? run_phase2<2, true>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s)
                        : run_phase2<2, false>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 3: rc = quad ? run_phase2<3, true>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s)
                        : run_phase2<3, false>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 4: rc = quad ? run_phase2<4, true>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s)
                        : run_phase2<4, false>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        case 5: rc = quad ? run_phase2<5, true>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s)
                        : run_phase2<5, false>(x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s); break;
        default: set_error("ssm_scores: state dimension %d has no device implementation (1 .. 5)", dx); return SSM_E_UNSUPPORTED;
    }
    if (rc == SSM_E_CUDA) set_error("ssm_scores_phase2: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

extern "C" int ssm_scores_phase2_traj(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                      const double *mse, double *lcr, double *lcr_acc, int64_t n_traj, int32_t n_steps,
                                      int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    return scores_phase2_impl(false, dx, x, mean, cov, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, stream);
}

extern "C" int ssm_scores_phase2_quad(int32_t dx, const double *x, const double *mean, const double *quad, const int32_t *status,
                                      const double *mse, double *lcr, double *lcr_acc, int64_t n_traj, int32_t n_steps,
                                      int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    return scores_phase2_impl(true, dx, x, mean, quad, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, stream);
}

extern "C" int ssm_scores_phase2_res(int32_t dx, const double *dres, const double *quad, const int32_t *status,
                                     const double *mse, double *lcr, double *lcr_acc, int64_t n_traj, int32_t n_steps,
                                     int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    if (!dres || !quad || !mse || !lcr) { set_error("ssm_scores_phase2_res: NULL buffer"); return SSM_E_INVALID; }
    if (n_traj <= 0 || n_steps <= 0 || ld < n_traj) { set_error("ssm_scores_phase2_res: bad sizes"); return SSM_E_INVALID; }
    if (k_lo < 0 || k_hi <= k_lo || k_hi > n_steps) { set_error("ssm_scores_phase2_res: bad time window [%d, %d) of %d steps", k_lo, k_hi, n_steps); return SSM_E_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
#define SSM_P2R_CASE(D) case D: rc = run_phase2<D, true, true>(dres, nullptr, quad, status, mse, lcr, lcr_acc, n_traj, n_steps, k_lo, k_hi, ld, s); break;
    switch (dx) {
        SSM_P2R_CASE(1) SSM_P2R_CASE(2) SSM_P2R_CASE(3) SSM_P2R_CASE(4) SSM_P2R_CASE(5)
        default: set_error("ssm_scores: state dimension %d has no device implementation (1 .. 5)", dx); return SSM_E_UNSUPPORTED;
    }
#undef SSM_P2R_CASE
    if (rc == SSM_E_CUDA) set_error("ssm_scores_phase2_res: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

extern "C" int ssm_scores_phase2_window(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                        const double *mse, double *lcr, int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi,
                                        int64_t ld, void *stream) {
    return ssm_scores_phase2_traj(dx, x, mean, cov, status, mse, lcr, nullptr, n_traj, n_steps, k_lo, k_hi, ld, stream);
}

extern "C" int ssm_scores_phase2(int32_t dx, const double *x, const double *mean, const double *cov, const int32_t *status,
                                 const double *mse, double *lcr, int64_t n_traj, int32_t n_steps, int64_t ld, void *stream) {
    return ssm_scores_phase2_window(dx, x, mean, cov, status, mse, lcr, n_traj, n_steps, 0, n_steps, ld, stream);
}
