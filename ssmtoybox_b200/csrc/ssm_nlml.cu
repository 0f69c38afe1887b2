// K5c: negative log marginal likelihood of the GP / TP regression model of the integrand and its gradient, batched
// over kernel log-parameter vectors (optimiser restarts, line-search candidates, parameter grids).
// Replaces GaussianProcessModel.neg_log_marginal_likelihood (bq/bqmod.py:537-596), StudentTProcessModel's
// (bq/bqmod.py:1191-1245) and RBFGauss.der_par (bq/bqkern.py:426-436):
//     par = exp(log_par);  K = k(X, X; par) + jitter;  L = chol(K);  A = K^-1 Y  (N, E)
//     GP:  nlml = E sum_i log L_ii + 1/2 (sum_e y_e' a_e + E N log 2 pi)
//     TP:  nlml = (nu + N)/2 sum_e log(1 + y_e' a_e / (nu - 2)) + E (sum_i log L_ii + N/2 log((nu - 2) pi)
//                 - lgamma((nu + N)/2) + lgamma(nu/2))
//     grad_p = 1/2 tr(E K^-1 dK_p - (sum_e s_e a_e a_e') dK_p),  s_e = 1 (GP) or (nu + N)/(nu + y_e' a_e - 2) (TP)
// with der_par exactly as the reference writes it: dK_0 = 2 K / alpha (derivative w.r.t. alpha), dK_d = K_ij
// (x_di - x_dj)^2 / l_d^2 (derivative w.r.t. log l_d).  One CTA per parameter vector; N <= 32, D <= 8, E <= 8: a
// latency-bound set-up kernel (the matrices are tiny), its value is the batch axis.
#include "ssm_common.cuh"

namespace ssm {

void set_error(const char *fmt, ...);

constexpr int NL_MAXN = 32, NL_MAXD = 8, NL_MAXE = 8, NL_THREADS = 128;

struct NlmlPar {
    int D, N, E, n_par;
    double nu;                 // 0: GP, > 2: TP
    const double *log_par;     // (n_par, D + 1)   device
    const double *x;           // (D, N)
    const double *y;           // (N, E)
    const double *jitter;      // (N, N) or NULL
    double *nlml, *grad;       // (n_par), (n_par, D + 1)
    int32_t *info;
};

__global__ void __launch_bounds__(NL_THREADS) gp_nlml_kernel(const NlmlPar p) {
    __shared__ double K[NL_MAXN][NL_MAXN + 1];    // kernel matrix without jitter (needed by der_par)
    __shared__ double L[NL_MAXN][NL_MAXN + 1];    // Cholesky factor of K + jitter
    __shared__ double S[NL_MAXN][NL_MAXN + NL_MAXE + 1];   // solutions of (K + jitter) S = [I | Y]
    __shared__ double X[NL_MAXD][NL_MAXN];
    __shared__ double par[NL_MAXD + 1], yda[NL_MAXE], sc[NL_MAXE];
    __shared__ double red[NL_THREADS / 32][NL_MAXD + 1];
    __shared__ int bad;
    const int D = p.D, N = p.N, E = p.E, tid = threadIdx.x, ip = blockIdx.x;
    if (tid <= D) par[tid] = exp(p.log_par[ip * (D + 1) + tid]);
    for (int e = tid; e < D * N; e += NL_THREADS) X[e / N][e % N] = p.x[e];
    if (tid == 0) bad = 0;
    __syncthreads();
    const double two_log_alpha = 2.0 * log(par[0]);
    for (int e = tid; e < N * N; e += NL_THREADS) {
        const int i = e / N, j = e % N;
        double m = 0.0;
        for (int d = 0; d < D; ++d) {
            const double u = (X[d][i] - X[d][j]) / par[1 + d];
            m = fma(u, u, m);
        }
        const double k = exp(two_log_alpha - 0.5 * m);     // bqkern.py:341-343
        K[i][j] = k;
        L[i][j] = k + (p.jitter ? p.jitter[e] : 0.0);
    }
    __syncthreads();
    // Cholesky, lower, column by column
    for (int c = 0; c < N; ++c) {
        if (tid == 0) {
            const double dcc = L[c][c];
            if (!(dcc > 0.0)) bad = 1;
            L[c][c] = sqrt(dcc);
        }
        __syncthreads();
        if (bad) break;
        const double inv = 1.0 / L[c][c];
        for (int r = c + 1 + tid; r < N; r += NL_THREADS) L[r][c] *= inv;
        __syncthreads();
        for (int e = tid; e < (N - c - 1) * (N - c - 1); e += NL_THREADS) {
            const int r = c + 1 + e / (N - c - 1), q = c + 1 + e % (N - c - 1);
            if (q <= r) L[r][q] = fma(-L[r][c], L[q][c], L[r][q]);
        }
        __syncthreads();
    }
    if (bad) {   // numpy.linalg.LinAlgError in the reference (scipy cho_factor, bqmod.py:578)
        if (tid == 0) { p.nlml[ip] = qnan(); p.info[ip] = 1; }
        if (tid <= D) p.grad[ip * (D + 1) + tid] = qnan();
        return;
    }
    // one right-hand side per thread: columns 0..N-1 of the identity, then the E columns of Y
    if (tid < N + E) {
        double v[NL_MAXN];
        for (int i = 0; i < N; ++i) v[i] = tid < N ? (i == tid ? 1.0 : 0.0) : p.y[i * E + (tid - N)];
        for (int i = 0; i < N; ++i) {
            double s = v[i];
            for (int j = 0; j < i; ++j) s = fma(-L[i][j], v[j], s);
            v[i] = s / L[i][i];
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = v[i];
            for (int j = i + 1; j < N; ++j) s = fma(-L[j][i], v[j], s);
            v[i] = s / L[i][i];
        }
        for (int i = 0; i < N; ++i) S[i][tid] = v[i];
        if (tid >= N) {
            double s = 0.0;
            for (int i = 0; i < N; ++i) s = fma(p.y[i * E + (tid - N)], v[i], s);
            yda[tid - N] = s;
            sc[tid - N] = p.nu > 0.0 ? (p.nu + N) / (p.nu + s - 2.0) : 1.0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double hl = 0.0;
        for (int i = 0; i < N; ++i) hl += log(L[i][i]);
        double v;
        if (p.nu > 0.0) {
            const double cst = 0.5 * N * log((p.nu - 2.0) * 3.141592653589793) - lgamma(0.5 * (p.nu + N)) + lgamma(0.5 * p.nu);
            double ls = 0.0;
            for (int e = 0; e < E; ++e) ls += log(1.0 + yda[e] / (p.nu - 2.0));
            v = 0.5 * (p.nu + N) * ls + E * (hl + cst);
        } else {
            double s = 0.0;
            for (int e = 0; e < E; ++e) s += yda[e];
            v = E * hl + 0.5 * (s + (double)E * N * log(2.0 * 3.141592653589793));
        }
        p.nlml[ip] = v;
        p.info[ip] = 0;
    }
    // gradient: 1/2 sum_ij (E iK_ij - sum_e s_e a_ie a_je) dK_p,ij
    double g[NL_MAXD + 1];
    for (int q = 0; q <= D; ++q) g[q] = 0.0;
    for (int e = tid; e < N * N; e += NL_THREADS) {
        const int i = e / N, j = e % N;
        double w = (double)E * S[i][j];
        for (int o = 0; o < E; ++o) w = fma(-sc[o] * S[i][N + o], S[j][N + o], w);
        const double wk = w * K[i][j];
        g[0] = fma(wk, 2.0 / par[0], g[0]);
        for (int d = 0; d < D; ++d) {
            const double u = (X[d][i] - X[d][j]) / par[1 + d];
            g[1 + d] = fma(wk, u * u, g[1 + d]);
        }
    }
    for (int q = 0; q <= D; ++q) {
        double v = g[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][q] = v;
    }
    __syncthreads();
    if (tid <= D) {
        double v = 0.0;
        for (int w = 0; w < NL_THREADS / 32; ++w) v += red[w][tid];
        p.grad[ip * (D + 1) + tid] = 0.5 * v;
    }
}

}  // namespace ssm

using namespace ssm;

extern "C" int ssm_gp_nlml(int32_t dim, int32_t n_pts, int32_t n_out, int32_t n_par, const double *log_par,
                           const double *x_obs, const double *fcn_obs, const double *jitter, double nu, double *nlml,
                           double *grad, int32_t *info, void *stream) {
    if (!log_par || !x_obs || !fcn_obs || !nlml || !grad || !info) { set_error("ssm_gp_nlml: NULL argument"); return SSM_E_INVALID; }
    if (dim < 1 || dim > NL_MAXD || n_pts < 1 || n_pts > NL_MAXN || n_out < 1 || n_out > NL_MAXE) {
        set_error("ssm_gp_nlml: supports dim <= %d, n_pts <= %d, n_out <= %d (got %d, %d, %d)", NL_MAXD, NL_MAXN, NL_MAXE, dim, n_pts, n_out);
        return SSM_E_UNSUPPORTED;
    }
    if (nu != 0.0 && !(nu > 2.0)) { set_error("ssm_gp_nlml: nu must be 0 (GP) or > 2 (TP)"); return SSM_E_INVALID; }
    if (n_par <= 0) return SSM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n_lp = (size_t)n_par * (dim + 1), n_x = (size_t)dim * n_pts, n_y = (size_t)n_pts * n_out;
    const size_t n_j = jitter ? (size_t)n_pts * n_pts : 0, n_in = n_lp + n_x + n_y + n_j;
    double *dev = nullptr;
    if (scratch_alloc((void **)&dev, n_in * sizeof(double), s) != cudaSuccess) { set_error("ssm_gp_nlml: cudaMallocAsync failed"); return SSM_E_CUDA; }
    double *host = (double *)malloc(n_in * sizeof(double));
    memcpy(host, log_par, n_lp * sizeof(double));
    memcpy(host + n_lp, x_obs, n_x * sizeof(double));
    memcpy(host + n_lp + n_x, fcn_obs, n_y * sizeof(double));
    if (jitter) memcpy(host + n_lp + n_x + n_y, jitter, n_j * sizeof(double));
    cudaMemcpyAsync(dev, host, n_in * sizeof(double), cudaMemcpyHostToDevice, s);   // pageable source: staged before return
    NlmlPar p;
    p.D = dim; p.N = n_pts; p.E = n_out; p.n_par = n_par; p.nu = nu;
    p.log_par = dev; p.x = dev + n_lp; p.y = dev + n_lp + n_x; p.jitter = jitter ? dev + n_lp + n_x + n_y : nullptr;
    p.nlml = nlml; p.grad = grad; p.info = info;
    gp_nlml_kernel<<<n_par, NL_THREADS, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dev, s);
    free(host);
    if (e != cudaSuccess) { set_error("ssm_gp_nlml: CUDA error: %s", cudaGetErrorString(e)); return SSM_E_CUDA; }
    return SSM_OK;
}
