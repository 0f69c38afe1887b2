// Bootstrap estimate of the variance of a sample mean, on the device.
// Replaces utils.bootstrap_var (utils.py:223-244): `np.var(np.mean(np.random.choice(data, (samples, n)), 1))`, which
// the research drivers call six times per algorithm with 10 000 resamples (research/gpq/icinco_demo.py:54-66,
// research/bsq/bsq_ungm.py:64-76) -- 10^4 x n gathers, n = number of Monte-Carlo trajectories.
//
// One CTA per resample: its threads draw the n indices with Philox4x32-10 keyed by (seed, resample index) and counted
// by the draw index (two 64-bit words per call -> two indices, multiply-high mapping onto [0, n)), gather from the
// data vector (n <= 10^7 doubles: L2-resident) and reduce in a fixed order, so the result depends on (data, seed,
// samples) only.  A second single-CTA kernel takes the population variance of the resample means (np.var, ddof = 0).
// numpy's MT19937 stream cannot be reproduced: validated statistically against the reference.
#include "ssm_rng.cuh"

namespace ssm {

void set_error(const char *fmt, ...);

constexpr int BS_THREADS = 256;

SSM_DEV double block_sum(double s, double *smem /* [BS_THREADS / 32] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[wid] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += smem[w];
    return tot;  // every thread holds the total
}

__global__ void __launch_bounds__(BS_THREADS) bootstrap_means_kernel(const double *__restrict__ data, long long n, unsigned long long seed,
                                                                     double *__restrict__ means) {
    __shared__ double smem[BS_THREADS / 32];
    Philox ph;
    ph.k0 = (uint32_t)seed;
    ph.k1 = (uint32_t)(seed >> 32);
    const uint32_t b = blockIdx.x;
    double s = 0.0;
    const long long n_pairs = (n + 1) / 2;
    for (long long j = threadIdx.x; j < n_pairs; j += BS_THREADS) {
        uint32_t r[4];
        ph.gen((uint32_t)j, (uint32_t)(j >> 32), b, 0x626f6f74u /* 'boot' */, r);
        const unsigned long long w0 = ((unsigned long long)r[1] << 32) | r[0], w1 = ((unsigned long long)r[3] << 32) | r[2];
        s += __ldg(data + __umul64hi(w0, (unsigned long long)n));
        if (2 * j + 1 < n) s += __ldg(data + __umul64hi(w1, (unsigned long long)n));
    }
    const double tot = block_sum(s, smem);
    if (threadIdx.x == 0) means[b] = tot / (double)n;
}

__global__ void __launch_bounds__(BS_THREADS) bootstrap_var_kernel(const double *__restrict__ means, int n_boot, double *__restrict__ var) {
    __shared__ double smem[BS_THREADS / 32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n_boot; i += BS_THREADS) s += means[i];
    const double mu = block_sum(s, smem) / n_boot;
    double q = 0.0;
    for (int i = threadIdx.x; i < n_boot; i += BS_THREADS) {
        const double d = means[i] - mu;
        q = fma(d, d, q);
    }
    const double tot = block_sum(q, smem);
    if (threadIdx.x == 0) var[0] = tot / n_boot;
}

}  // namespace ssm

using namespace ssm;

extern "C" int ssm_bootstrap_var(const double *data, int64_t n, int32_t n_boot, uint64_t seed, double *means, double *var,
                                 void *stream) {
    if (!data || !means || !var) { set_error("ssm_bootstrap_var: NULL buffer"); return SSM_E_INVALID; }
    if (n <= 0 || n_boot <= 0) { set_error("ssm_bootstrap_var: bad sizes"); return SSM_E_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    bootstrap_means_kernel<<<(unsigned)n_boot, BS_THREADS, 0, s>>>(data, n, seed, means);
    bootstrap_var_kernel<<<1, BS_THREADS, 0, s>>>(means, n_boot, var);
    if (cudaGetLastError() != cudaSuccess) {
        set_error("ssm_bootstrap_var: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
        return SSM_E_CUDA;
    }
    return SSM_OK;
}

// ---- device math probe (test hook, include/ssm_b200.h) --------------------------------------------------------------
namespace ssm {
__global__ void math_probe_kernel(int which, const double *a, const double *b, double *out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = a[i], y = (which >= 3) ? b[i] : 0.0;
    double r;
    switch (which) {
        case 0: r = m_exp(x); break;
        case 1: r = m_sqrt(x); break;
        case 2: r = m_rsqrt(x); break;
        case 3: r = m_div(x, y); break;
        default: r = m_atan2(x, y); break;
    }
    out[i] = r;
}
}  // namespace ssm

extern "C" int ssm_math_probe(int32_t which, const double *a, const double *b, double *out, int64_t n, void *stream) {
    if (!a || !out || (which >= 3 && !b) || which < 0 || which > 4 || n < 0) { set_error("ssm_math_probe: bad arguments"); return SSM_E_INVALID; }
    if (n == 0) return SSM_OK;
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(which, a, b, out, (long long)n);
    return cudaGetLastError() == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}
