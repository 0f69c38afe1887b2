// Per-step error statistics of one trajectory and the CTA-level reduction used by K6 and by the kernels that
// accumulate scores in place (RTS smoother).  Row layout (width W = DX + DX*DX + 3):
//   [ squared error (DX) | d d^T (DX*DX) | negative log-likelihood | |d| | 1 ]
// squared_error utils.py:18-38, mse_matrix summand utils.py:62-64, neg_log_likelihood utils.py:123-148.
#pragma once
#include "ssm_common.cuh"

namespace ssm {

constexpr int SC_THREADS = 128;

template <int DX>
struct ScoreRow {
    static constexpr int W = DX + DX * DX + 3;                    // public row width (full d d^T)
    static constexpr int WP = DX + TriSize<DX>::value + 3;        // packed row reduced in-kernel: lower triangle of d d^T
};

// d = x - m, P = packed lower covariance; fills v[W]; returns the squared errors through se[]
// packed row: [ squared error (DX) | lower triangle of d d^T (DX(DX+1)/2) | NLL | |d| | 1 ]; the finalise kernel
// mirrors the triangle into the public full-matrix layout (a third fewer shuffles in the per-step reduction)
template <int DX>
SSM_DEV void score_step(const double (&d)[DX], const double (&P)[TriSize<DX>::value], double (&v)[ScoreRow<DX>::WP], double (&se)[DX]) {
    constexpr int TX = TriSize<DX>::value;
    double sse = 0.0;
#pragma unroll
    for (int a = 0; a < DX; ++a) {
        const double s = d[a] * d[a];
        v[a] = s;
        se[a] = s;
        sse += s;
    }
#pragma unroll
    for (int r = 0; r < DX; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) v[DX + tri(r, c)] = d[r] * d[c];
    // 0.5 (log|P| + d' P^-1 d + dx log 2 pi) through chol(P)
    double L[TX];
    const bool ok = chol_lower<DX>(P, L);
    double logdet = 0.0, quad = 0.0, z[DX];
#pragma unroll
    for (int i = 0; i < DX; ++i) {
        double s = d[i];
#pragma unroll
        for (int c = 0; c < i; ++c) s = fma(-L[tri(i, c)], z[c], s);
        z[i] = s / L[tri(i, i)];
        quad = fma(z[i], z[i], quad);
        logdet += log(L[tri(i, i)]);
    }
    v[DX + TX] = ok ? 0.5 * (2.0 * logdet + quad + DX * 1.8378770664093453) : qnan();
    v[DX + TX + 1] = sqrt(sse);  // per-trajectory error norm, bsq_tracking.py:331
    v[DX + TX + 2] = 1.0;
}

// sum v[] over the CTA (warp shuffles, then one shared-memory pass in fixed order) and store the row
template <int W>
SSM_DEV void block_reduce_store(double (&v)[W], double *smem /* [blockDim/32][W] */, double *dst) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < W; ++i) smem[wid * W + i] = v[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += smem[w * W + i];
        dst[i] = s;
    }
    __syncthreads();
}

// stats[i] = sum over CTAs (fixed order) of partial[cta][i]
__global__ void scores_finalize_kernel(const double *__restrict__ partial, double *__restrict__ stats, int n_cta, long long row);
// same, expanding packed rows (width WP) into public rows (width W) with the symmetric matrix mirrored
__global__ void scores_finalize_packed_kernel(const double *__restrict__ partial, double *__restrict__ stats, int n_cta, int n_steps, int dx);

}  // namespace ssm
