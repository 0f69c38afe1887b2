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
SSM_DEV void score_step(const double (&d)[DX], const double (&P)[TriSize<DX>::value], double (&v)[ScoreRow<DX>::WP], double (&se)[DX],
                        double *quad_out = nullptr) {
    constexpr int TX = TriSize<DX>::value;
    double sse = 0.0;
#pragma unroll
    for (int a = 0; a < DX; ++a) {
        // rounded product: the callers add it to running sums, and a multiply-add contracted in one kernel but not in
        // another would break the bitwise equality of the scoring passes (stand-alone, in-smoother, in-filter)
        const double s = __dmul_rn(d[a], d[a]);
        v[a] = s;
        se[a] = s;
        sse = __dadd_rn(sse, s);
    }
#pragma unroll
    for (int r = 0; r < DX; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) v[DX + tri(r, c)] = d[r] * d[c];
    // 0.5 (log|P| + d' P^-1 d + dx log 2 pi) through chol(P)
    // One log per step: log|P| = 2 log(prod L_ii) (the product of DX <= 5 pivots cannot leave the double range for
    // any covariance a filter produces), and the forward substitution multiplies by the reciprocal pivots that the
    // factorisation already has.  A few ulp from the reference's slogdet / solve, far below the 1e-8 score tolerance.
    double L[TX], inv[DX];
    const bool ok = chol_lower<DX>(P, L, inv);
    double pdiag = 1.0, quad = 0.0, z[DX];
#pragma unroll
    for (int i = 0; i < DX; ++i) {
        double s = d[i];
#pragma unroll
        for (int c = 0; c < i; ++c) s = fma(-L[tri(i, c)], z[c], s);
        z[i] = s * inv[i];
        quad = fma(z[i], z[i], quad);
        pdiag *= L[tri(i, i)];
    }
    const double logdet = log(pdiag);
    v[DX + TX] = ok ? 0.5 * (2.0 * logdet + quad + DX * 1.8378770664093453) : qnan();
    // d' P^-1 d, the first quadratic form of the log credibility ratio (utils.py:113-120): identical arithmetic to
    // scores_phase2_kernel, so a pass that keeps it saves phase 2 the covariance read and its factorisation
    if (quad_out) *quad_out = ok ? quad : qnan();
    v[DX + TX + 1] = sqrt(sse);  // per-trajectory error norm, bsq_tracking.py:331
    v[DX + TX + 2] = 1.0;
}

// Sum v[] over the CTA and store the row.  The values are transposed through shared memory ([W][SC_THREADS + 1],
// conflict-free stores), then four threads per row add 32 values each in a fixed rotated order and combine with two
// shuffles: W stores + 32 loads/adds per thread instead of 5 W shuffle pairs (the shuffle tree was a fifth of the
// smoother's instructions).  Two buffers alternate by step parity, so one barrier per step is enough: a buffer is
// rewritten two steps later, after the barrier of the step in between.  Fixed order => bitwise reproducible.
template <int W>
struct BlockReduce {
    static constexpr int LD = SC_THREADS + 1;
    static constexpr int SIZE = 2 * W * LD;  // doubles of shared memory
    static_assert(4 * W <= SC_THREADS, "four threads per row");
};
template <int W>
SSM_DEV void block_reduce_store(const double (&v)[W], double *smem /* [BlockReduce<W>::SIZE] */, int parity, double *dst, int n_rows = W) {
    constexpr int LD = BlockReduce<W>::LD;
    double *buf = smem + (parity & 1) * (W * LD);
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < W; ++i) buf[i * LD + tid] = v[i];
    __syncthreads();
    if ((tid & ~31) < 4 * W) {  // warp-uniform
        const int row = tid >> 2, seg = tid & 3;
        const double *src = buf + (row < W ? row : W - 1) * LD + seg * 32;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {  // element order rotated by 4 seg: the 16 lanes of a half-warp hit 16 banks
            s0 += src[(i + 4 * seg) & 31];
            s1 += src[(i + 1 + 4 * seg) & 31];
            s2 += src[(i + 2 + 4 * seg) & 31];
            s3 += src[(i + 3 + 4 * seg) & 31];
        }
        double s = (s0 + s1) + (s2 + s3);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (seg == 0 && row < n_rows) dst[row] = s;   // n_rows < W: the last, partial group of a multi-step reduction
    }
}

constexpr int FIN_GROUPS = 8;  // blockDim.y of the finalise kernels (ssm_scores.cu)
// stats[i] = sum over CTAs (fixed order) of partial[cta][i]
__global__ void scores_finalize_kernel(const double *__restrict__ partial, double *__restrict__ stats, int n_cta, long long row);
// same, expanding packed rows (width WP) into public rows (width W) with the symmetric matrix mirrored
__global__ void scores_finalize_packed_kernel(const double *__restrict__ partial, double *__restrict__ stats, int n_cta, int n_steps, int dx);

}  // namespace ssm
