// K5: batched Bayesian-quadrature weights, one CTA per kernel-parameter vector (hyper-parameter
// sweeps launch tens to hundreds of CTAs).  N <= 64 points, D <= 8: latency-bound set-up work where
// correctness matters, not speed; every small matrix lives in an L2-resident per-CTA workspace.
//
// Replaces (file:line in /root/reference/ssmtoybox/bq):
//   RBFGauss.eval (scaling=False)            bqkern.py:329-343 (utils.maha, utils.py:385-409)
//   Kernel.eval_inv_dot / _cho_inv           bqkern.py:38-64, 96-120  (jitter 1e-8, symmetrised)
//   RBFGauss.exp_x_kx / exp_x_xkx / exp_x_kxkx / exp_xy_kxy     bqkern.py:345-424
//   GaussianProcessModel.bq_weights          bqmod.py:495-523
//   BayesSardModel.bq_weights                bqmod.py:893-992 (+ _exp_x_kxpx :733-797,
//                                            utils.vandermonde utils.py:478-502)
// The hyper-parameter-free polynomial expectations _exp_x_px / _xpx / _pxpx (bqmod.py:635-731) are
// integer combinatorics of the multi-index and are evaluated by the host part of this file.
#include <vector>

#include "ssm_dd.cuh"

namespace ssm {

void set_error(const char *fmt, ...);

constexpr int W_MAXN = 64, W_MAXD = 8;

struct WeightsPar {
    int D, N, Q;       // Q = 0: plain GP weights
    int n_par;
    const double *par;     // (n_par, D+1) device
    const double *x;       // (D, N) device
    const int *mulind;     // (D, Q) device
    const double *px, *xpx, *pxpx;  // (Q), (D, Q), (Q, Q) device
    double *wm, *Wc, *Wcc, *iK, *scal;
    int *info;
    double *work;          // per-CTA workspace
    long long work_stride;
};

// ---- CTA-cooperative dense helpers (row-major, global/L2 memory), scalar type T = double or dd ------
// C (m x n) = op(A) (m x k) . op(B) (k x n); lda/ldb are the leading dimensions of the stored arrays
template <class T, class TA, class TB>
__device__ void mm(T *C, const TA *A, bool tA, int lda, const TB *B, bool tB, int ldb, int m, int k, int n) {
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        T s(0.0);
        for (int l = 0; l < k; ++l) {
            const T a(tA ? A[l * lda + i] : A[i * lda + l]);
            const T b(tB ? B[j * ldb + l] : B[l * ldb + j]);
            s = s + a * b;
        }
        C[e] = s;
    }
    __syncthreads();
}

// in-place lower Cholesky of the n x n matrix A; returns false when not positive definite
template <class T>
__device__ bool chol_cta(T *A, int n, int *flag) {
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        if (threadIdx.x == 0) {
            T s = A[j * n + j];
            for (int k = 0; k < j; ++k) s = s - A[j * n + k] * A[j * n + k];
            if (!(s > 0.0)) *flag = 1;
            A[j * n + j] = tsqrt(s);
        }
        __syncthreads();
        const T d = A[j * n + j];
        for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) {
            T t = A[i * n + j];
            for (int k = 0; k < j; ++k) t = t - A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / d;
        }
        __syncthreads();
    }
    return *flag == 0;
}

// X = (L L^T)^-1 by column-wise forward/back substitution with the identity (cho_solve(., I)); symmetrised
// 0.5 (X + X^T) when sym (Kernel._cho_inv, bqkern.py:59-63).  T_ is scratch (n x n).
template <class T>
__device__ void chol_inverse(T *X, const T *L, T *T_, int n, bool sym) {
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        for (int i = 0; i < n; ++i) {  // L y = e_c
            T t((i == c) ? 1.0 : 0.0);
            for (int k = 0; k < i; ++k) t = t - L[i * n + k] * T_[k * n + c];
            T_[i * n + c] = t / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {  // L^T x = y
            T t = T_[i * n + c];
            for (int k = i + 1; k < n; ++k) t = t - L[k * n + i] * T_[k * n + c];
            T_[i * n + c] = t / L[i * n + i];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        X[e] = sym ? T(0.5) * (T_[i * n + j] + T_[j * n + i]) : T_[e];
    }
    __syncthreads();
}

// general inverse by Gaussian elimination with partial pivoting (scipy.linalg.solve(V, I), bqmod.py:949)
// A is destroyed, X receives A^-1.  Serial pivot search, parallel row updates.
template <class T>
__device__ bool lu_inverse(T *X, T *A, int n, int *flag) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) X[e] = T((e / n == e % n) ? 1.0 : 0.0);
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        if (threadIdx.x == 0) {
            int piv = j;
            double best = fabs(to_double(A[j * n + j]));
            for (int i = j + 1; i < n; ++i)
                if (fabs(to_double(A[i * n + j])) > best) { best = fabs(to_double(A[i * n + j])); piv = i; }
            if (!(best > 0.0)) *flag = 1;
            flag[1] = piv;
        }
        __syncthreads();
        const int piv = flag[1];
        if (piv != j) {
            for (int c = threadIdx.x; c < n; c += blockDim.x) {
                T t = A[j * n + c]; A[j * n + c] = A[piv * n + c]; A[piv * n + c] = t;
                t = X[j * n + c]; X[j * n + c] = X[piv * n + c]; X[piv * n + c] = t;
            }
        }
        __syncthreads();
        const T d = A[j * n + j];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (i == j) continue;
            const T f = A[i * n + j] / d;
            for (int c = 0; c < n; ++c) {
                if (c > j) A[i * n + c] = A[i * n + c] - f * A[j * n + c];
                X[i * n + c] = X[i * n + c] - f * X[j * n + c];
            }
            A[i * n + j] = T(0.0);
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) X[e] = X[e] / A[(e / n) * n + (e / n)];
    __syncthreads();
    return *flag == 0;
}

// tr(A B) with A (n x m), B (m x n); serial in thread 0 (tiny, deterministic)
template <class T, class TA, class TB>
__device__ T trace_prod(const TA *A, const TB *B, int n, int m, T *red) {
    if (threadIdx.x == 0) {
        T s(0.0);
        for (int e = 0; e < n * m; ++e) s = s + T(A[e]) * T(B[(e % m) * n + e / m]);
        red[0] = s;
    }
    __syncthreads();
    const T r = red[0];
    __syncthreads();
    return r;
}

template <class T, class TA, class TB>
__device__ T dot_cta(const TA *a, const TB *b, int n, T *red) {
    if (threadIdx.x == 0) {
        T t(0.0);
        for (int i = 0; i < n; ++i) t = t + T(a[i]) * T(b[i]);
        red[0] = t;
    }
    __syncthreads();
    const T r = red[0];
    __syncthreads();
    return r;
}

template <class T>
__global__ void __launch_bounds__(128) bq_weights_kernel(const WeightsPar p) {
    __shared__ T red[2];
    __shared__ int flag[2];
    __shared__ double ell[W_MAXD];
    const int D = p.D, N = p.N, Q = p.Q, ip = blockIdx.x;
    const double *par = p.par + (long long)ip * (D + 1);
    const double *x = p.x;
    T *w = reinterpret_cast<T *>(p.work + (long long)ip * p.work_stride);
    const int NN = N * N;
    // workspace carving
    T *K = w;            w += NN;   // kernel matrix, then its Cholesky factor
    T *iK = w;           w += NN;
    T *Qm = w;           w += NN;
    T *T1 = w;           w += NN;
    T *T2 = w;           w += NN;
    T *q = w;            w += N;
    T *R = w;            w += D * N;
    T *xs = w;           w += D * N;   // scaled points
    T *x2 = w;           w += N;
    T *wmT = w;          w += N;
    T *WccT = w;         w += D * N;
    int info = 0;
    const T alpha(par[0]);
    if (threadIdx.x < D) ell[threadIdx.x] = par[1 + threadIdx.x];
    __syncthreads();
    auto il2 = [&](int d) { return T(1.0) / (T(ell[d]) * T(ell[d])); };   // 1 / l^2

    // ---- K = exp(-0.5 maha(x / l)), scaling=False                      bqkern.py:329-343
    for (int e = threadIdx.x; e < D * N; e += blockDim.x) xs[e] = (T(1.0) / T(ell[e / N])) * T(x[e]);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        T s(0.0);
        for (int d = 0; d < D; ++d) s = s + xs[d * N + i] * xs[d * N + i];
        x2[i] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
        const int i = e / N, j = e % N;
        T cr(0.0);
        for (int d = 0; d < D; ++d) cr = cr + xs[d * N + i] * xs[d * N + j];
        const T mh = (x2[i] + x2[j]) - T(2.0) * cr;
        K[e] = texp(T(-0.5) * mh) + T((i == j) ? 1e-8 : 0.0);  // + jitter I, bqkern.py:120
    }
    __syncthreads();
    // ---- iK = sym(cho_solve(cho_factor(K + jitter I), I))               bqkern.py:38-64
    if (!chol_cta<T>(K, N, flag)) info |= 1;
    chol_inverse<T>(iK, K, T1, N, true);

    // ---- q = E[k(x, x_i)], R = E[x k(x, x_i)]                           bqkern.py:345-364
    T cdet(1.0), rdet(1.0);
    for (int d = 0; d < D; ++d) {
        cdet = cdet * (il2(d) + T(1.0));
        rdet = rdet * (T(2.0) * il2(d) + T(1.0));   // |2 Lambda^-1 + I|, also det(r) of exp_x_kxkx
    }
    const T cq = T(1.0) / tsqrt(cdet);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        T s(0.0);
        for (int d = 0; d < D; ++d) {
            const T lam1 = T(1.0) / (T(ell[d]) * T(ell[d]) + T(1.0));  // (Lambda + I)^-1
            s = s + T(x[d * N + i]) * (lam1 * T(x[d * N + i]));
        }
        q[i] = cq * texp(T(-0.5) * s);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * N; e += blockDim.x) {
        const int d = e / N;
        R[e] = q[e % N] * ((T(1.0) / (T(ell[d]) * T(ell[d]) + T(1.0))) * T(x[e]));
    }
    // ---- Q_ij = E[k(x, x_i) k(x, x_j)]                                  bqkern.py:366-415
    const T cQ = T(1.0) / tsqrt(rdet);
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
        const int i = e / N, j = e % N;
        const T n = T(-0.5) * x2[i] + T(-0.5) * x2[j];
        T mh(0.0);
        for (int d = 0; d < D; ++d) {
            const T a = il2(d) * T(x[d * N + i]) + il2(d) * T(x[d * N + j]);
            mh = mh + (a * a) * (T(1.0) / (T(2.0) * il2(d) + T(1.0)));
        }
        Qm[e] = cQ * texp(n + T(0.5) * mh);
    }
    __syncthreads();
    const T kbar = alpha * alpha / tsqrt(rdet);   // bqkern.py:421-424

    double *wm = p.wm + (long long)ip * N;
    double *Wc = p.Wc + (long long)ip * NN;
    double *Wcc = p.Wcc + (long long)ip * D * N;
    T model_var, integral_var;

    if (Q == 0) {
        // ---- GP weights                                                   bqmod.py:508-517
        mm<T>(wmT, q, false, N, iK, false, N, 1, N, N);
        mm<T>(T1, iK, false, N, Qm, false, N, N, N, N);
        mm<T>(T2, T1, false, N, iK, false, N, N, N, N);
        mm<T>(WccT, R, false, N, iK, false, N, D, N, N);
        model_var = alpha * alpha * (T(1.0) - trace_prod<T>(Qm, iK, N, N, red));
        integral_var = kbar - dot_cta<T>(wmT, q, N, red);
    } else {
        const int M = N > Q ? N : Q;
        T *V = w;        w += N * Q;     // Vandermonde (N x Q)
        T *kxpx = w;     w += N * Q;
        T *iViKV = w;    w += Q * Q;
        T *S1 = w;       w += M * M;
        T *S2 = w;       w += M * M;
        T *S3 = w;       w += M * M;
        T *v1 = w;       w += M;
        T *v2 = w;       w += M;
        // V[n, b] = prod_d x[d, n] ** mulind[d, b]                          utils.py:478-502
        for (int e = threadIdx.x; e < N * Q; e += blockDim.x) {
            const int n = e / Q, b = e % Q;
            T pr(1.0);
            for (int d = 0; d < D; ++d) pr = pr * tpowi<T>(T(x[d * N + n]), p.mulind[d * Q + b]);
            V[e] = pr;
        }
        // kxpx[n, q] = prod_d a_d b_d                                       bqmod.py:733-797
        // (quirk reproduced: the reference's "ell" is the SQUARED length-scale and is squared again)
        for (int e = threadIdx.x; e < N * Q; e += blockDim.x) {
            const int n = e / Q, b = e % Q;
            T pr(1.0);
            for (int d = 0; d < D; ++d) {
                const int a = p.mulind[d * Q + b];
                const T el = T(ell[d]) * T(ell[d]);  // l^2
                const T el2 = el * el;                // "ell ** 2"
                const T xv(x[d * N + n]);
                const T one_el2 = T(1.0) + el2;
                // (1 + el2) ** (-(1 + a) / 2)
                T pw = T(1.0) / tpowi<T>(tsqrt(one_el2), 1 + a);
                const T ea = el * pw * texp(-(xv * xv) / (T(2.0) * one_el2));
                T bs(0.0);
                const T xn = xv / tsqrt(one_el2);
                for (int m = 0; m <= a / 2; ++m) {
                    double fa = 1.0, fm = 1.0, fam = 1.0;
                    for (int u = 2; u <= a; ++u) fa *= u;
                    for (int u = 2; u <= m; ++u) fm *= u;
                    for (int u = 2; u <= a - 2 * m; ++u) fam *= u;
                    const T p1(fa / (ldexp(1.0, m) * fm * fam));
                    const T p2 = tpowi<T>(el, 2 * m) * tpowi<T>(xn, a - 2 * m);
                    bs = bs + p1 * p2;
                }
                pr = pr * (ea * bs);
            }
            kxpx[e] = pr;
        }
        __syncthreads();
        // iViKV = inv(V^T iK V + 1e-8 I) by Cholesky (NOT symmetrised in the reference)   bqmod.py:936
        mm<T>(S1, V, true, Q, iK, false, N, Q, N, N);          // Z = V^T iK  (Q x N)
        mm<T>(S2, S1, false, N, V, false, Q, Q, N, Q);          // V^T iK V
        for (int i = threadIdx.x; i < Q; i += blockDim.x) S2[i * Q + i] = S2[i * Q + i] + T(1e-8);
        __syncthreads();
        if (!chol_cta<T>(S2, Q, flag)) info |= 2;
        chol_inverse<T>(iViKV, S2, S3, Q, false);
        if (Q == N) {
            // ---- pi-unisolvent special case: classical rule via iV = V^-1   bqmod.py:948-961
            T *iV = S3;
            for (int e = threadIdx.x; e < NN; e += blockDim.x) S2[e] = V[e];
            __syncthreads();
            if (!lu_inverse<T>(iV, S2, N, flag)) info |= 4;
            mm<T>(wmT, p.px, false, Q, iV, false, N, 1, Q, N);                  // iV^T px
            mm<T>(T1, iV, true, N, p.pxpx, false, Q, N, Q, Q);                  // iV^T pxpx
            mm<T>(T2, T1, false, Q, iV, false, N, N, Q, N);                     // . iV
            mm<T>(WccT, p.xpx, false, Q, iV, false, N, D, Q, N);                // xpx iV
            // model_var = a^2 (1 - tr(kxpx^T iV^T + kxpx iV - pxpx iViKV))
            const T t1 = trace_prod<T>(kxpx, iV, N, Q, red);                    // tr(kxpx iV) = tr(kxpx^T iV^T)
            const T t3 = trace_prod<T>(p.pxpx, iViKV, Q, Q, red);
            model_var = alpha * alpha * (T(1.0) - (t1 + t1 - t3));
            // integral_var = kbar - q^T iV^T px - px^T iV q + px^T iViKV px
            const T a1 = dot_cta<T>(wmT, q, N, red);
            mm<T>(v1, p.px, false, Q, iViKV, false, Q, 1, Q, Q);
            const T a3 = dot_cta<T>(v1, p.px, Q, red);
            integral_var = kbar - a1 - a1 + a3;
        } else {
            // ---- general case                                              bqmod.py:963-982
            T *Z = S1;                                   // (Q x N), still holds V^T iK
            T *A = w;      w += N * Q;                   // V iViKV (N x Q)
            T *B = w;      w += Q * Q;
            T *Dm = w;     w += D * Q;
            T *b = w;      w += Q;
            mm<T>(A, V, false, Q, iViKV, false, Q, N, Q, Q);
            mm<T>(b, Z, false, N, q, false, 1, Q, N, 1);
            for (int i = threadIdx.x; i < Q; i += blockDim.x) b[i] = b[i] - T(p.px[i]);
            // B = Z Q Z^T + pxpx - Z kxpx - kxpx^T Z^T
            mm<T>(S2, Z, false, N, Qm, false, N, Q, N, N);
            mm<T>(B, S2, false, N, Z, true, N, Q, N, Q);
            mm<T>(S3, Z, false, N, kxpx, false, Q, Q, N, Q);      // Z kxpx (Q x Q)
            for (int e = threadIdx.x; e < Q * Q; e += blockDim.x)
                B[e] = B[e] + T(p.pxpx[e]) - S3[e] - S3[(e % Q) * Q + e / Q];
            __syncthreads();
            // D = R Z^T - xpx
            mm<T>(Dm, R, false, N, Z, true, N, D, N, Q);
            for (int e = threadIdx.x; e < D * Q; e += blockDim.x) Dm[e] = Dm[e] - T(p.xpx[e]);
            __syncthreads();
            // w_m = iK (q - A b)
            mm<T>(v1, A, false, Q, b, false, 1, N, Q, 1);
            for (int i = threadIdx.x; i < N; i += blockDim.x) v1[i] = q[i] - v1[i];
            __syncthreads();
            mm<T>(wmT, iK, false, N, v1, false, 1, N, N, 1);
            // w_c = iK (Q - A B A^T) iK
            mm<T>(S2, A, false, Q, B, false, Q, N, Q, Q);
            mm<T>(S3, S2, false, Q, A, true, Q, N, Q, N);
            for (int e = threadIdx.x; e < NN; e += blockDim.x) S3[e] = Qm[e] - S3[e];
            __syncthreads();
            mm<T>(T1, iK, false, N, S3, false, N, N, N, N);
            mm<T>(T2, T1, false, N, iK, false, N, N, N, N);
            // w_cc = (R - D A^T) iK
            mm<T>(S2, Dm, false, Q, A, true, Q, D, Q, N);
            for (int e = threadIdx.x; e < D * N; e += blockDim.x) S2[e] = R[e] - S2[e];
            __syncthreads();
            mm<T>(WccT, S2, false, N, iK, false, N, D, N, N);
            model_var = alpha * alpha * (T(1.0) - trace_prod<T>(Qm, iK, N, N, red) + trace_prod<T>(B, iViKV, Q, Q, red));
            mm<T>(v2, q, false, N, iK, false, N, 1, N, N);
            const T a1 = dot_cta<T>(v2, q, N, red);
            mm<T>(v2, b, false, Q, iViKV, false, Q, 1, Q, Q);
            integral_var = kbar - a1 + dot_cta<T>(v2, b, Q, red);
        }
    }
    // results rounded to float64 once; covariance weights symmetrised (bqmod.py:520-521, 985-986)
    for (int e = threadIdx.x; e < N; e += blockDim.x) wm[e] = to_double(wmT[e]);
    for (int e = threadIdx.x; e < D * N; e += blockDim.x) Wcc[e] = to_double(WccT[e]);
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
        const int i = e / N, j = e % N;
        Wc[e] = to_double(T(0.5) * (T2[i * N + j] + T2[j * N + i]));
    }
    if (p.iK)
        for (int e = threadIdx.x; e < NN; e += blockDim.x) p.iK[(long long)ip * NN + e] = to_double(iK[e]);
    if (threadIdx.x == 0) {
        p.scal[2 * ip] = to_double(model_var);
        p.scal[2 * ip + 1] = to_double(integral_var);
        p.info[ip] = info;
    }
}

// ---- host: polynomial expectations under N(0, I)                        bqmod.py:635-731 ------
static double fact2(int n) {  // (-1)!! = 0!! = 1
    double r = 1.0;
    for (int k = n; k > 1; k -= 2) r *= k;
    return r;
}

static void poly_expectations(int D, int Q, const int32_t *mi, std::vector<double> &px, std::vector<double> &xpx,
                              std::vector<double> &pxpx) {
    px.assign(Q, 0.0);
    xpx.assign((size_t)D * Q, 0.0);
    pxpx.assign((size_t)Q * Q, 0.0);
    auto a = [&](int d, int q) { return mi[d * Q + q]; };
    for (int q = 0; q < Q; ++q) {
        bool even = true;
        for (int d = 0; d < D; ++d) even = even && (a(d, q) % 2 == 0);
        if (even) {
            double pr = 1.0;
            for (int d = 0; d < D; ++d) pr *= fact2(a(d, q) - 1);
            px[q] = pr;
        }
    }
    for (int e = 0; e < D; ++e)
        for (int q = 0; q < Q; ++q) {
            bool ok = ((a(e, q) + 1) % 2 == 0);
            for (int d = 0; d < D; ++d)
                if (d != e) ok = ok && (a(d, q) % 2 == 0);
            if (ok) {
                double pr = a(e, q);
                for (int d = 0; d < D; ++d)
                    if (d != e) pr *= fact2(a(d, q) - 1);
                xpx[(size_t)e * Q + q] = pr;
            }
        }
    for (int r = 0; r < Q; ++r)
        for (int q = 0; q < Q; ++q) {
            bool even = true;
            for (int d = 0; d < D; ++d) even = even && ((a(d, r) + a(d, q)) % 2 == 0);
            if (even) {
                double pr = 1.0;
                for (int d = 0; d < D; ++d) pr *= fact2(a(d, r) + a(d, q) - 1);
                pxpx[(size_t)r * Q + q] = pr;
            }
        }
}

}  // namespace ssm

using namespace ssm;

extern "C" int ssm_bq_weights(int32_t dim, int32_t n_pts, int32_t n_par, const double *par, const double *points,
                              const int32_t *mulind, int32_t n_basis, double *wm, double *Wc, double *Wcc, double *iK,
                              double *scal, int32_t *info, int32_t precision, void *stream) {
    if (!par || !points || !wm || !Wc || !Wcc || !scal || !info) { set_error("ssm_bq_weights: NULL argument"); return SSM_E_INVALID; }
    if (dim < 1 || dim > W_MAXD || n_pts < 1 || n_pts > W_MAXN) {
        set_error("ssm_bq_weights: dim %d (<= %d) / n_pts %d (<= %d) out of range", dim, W_MAXD, n_pts, W_MAXN);
        return SSM_E_UNSUPPORTED;
    }
    const int Q = mulind ? n_basis : 0;
    if (mulind && (Q < 1 || Q > n_pts)) {
        set_error("ssm_bq_weights: number of basis functions (%d) must be in 1..n_pts (%d)", Q, n_pts);  // bqmod.py:983-984
        return SSM_E_INVALID;
    }
    if (n_par <= 0) return SSM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int D = dim, N = n_pts, M = N > Q ? N : Q;
    // host staging: par | points | px | xpx | pxpx  (doubles), then mulind (ints)
    std::vector<double> px, xpx, pxpx;
    if (Q) poly_expectations(D, Q, mulind, px, xpx, pxpx);
    const size_t n_in = (size_t)n_par * (D + 1) + (size_t)D * N + (size_t)Q + (size_t)D * Q + (size_t)Q * Q;
    const long long work_elems = 5LL * N * N + 3LL * N + 3LL * D * N + 3LL * N * Q + 2LL * Q * Q + 3LL * M * M + 2LL * M + (long long)D * Q + Q + 64;
    const long long work_stride = work_elems * (precision ? 2 : 1);   // in doubles; dd = 2 doubles per element
    const size_t bytes = (n_in + (size_t)work_stride * n_par) * sizeof(double) + (size_t)(D * Q + 2) * sizeof(int);
    double *dev = nullptr;
    if (scratch_alloc((void **)&dev, bytes, s) != cudaSuccess) { set_error("ssm_bq_weights: cudaMallocAsync failed"); return SSM_E_CUDA; }
    std::vector<double> host(n_in);
    size_t off = 0;
    WeightsPar p;
    memset(&p, 0, sizeof(p));
    auto put = [&](const double *src, size_t cnt, const double *&dst) {
        if (cnt) memcpy(host.data() + off, src, cnt * sizeof(double));
        dst = dev + off;
        off += cnt;
    };
    put(par, (size_t)n_par * (D + 1), p.par);
    put(points, (size_t)D * N, p.x);
    put(px.data(), Q, p.px);
    put(xpx.data(), (size_t)D * Q, p.xpx);
    put(pxpx.data(), (size_t)Q * Q, p.pxpx);
    cudaMemcpyAsync(dev, host.data(), n_in * sizeof(double), cudaMemcpyHostToDevice, s);
    p.work = dev + n_in;
    p.work_stride = work_stride;
    int *dmi = (int *)(dev + n_in + (size_t)work_stride * n_par);
    if (Q) cudaMemcpyAsync(dmi, mulind, (size_t)D * Q * sizeof(int), cudaMemcpyHostToDevice, s);
    p.mulind = dmi;
    p.D = D; p.N = N; p.Q = Q; p.n_par = n_par;
    p.wm = wm; p.Wc = Wc; p.Wcc = Wcc; p.iK = iK; p.scal = scal; p.info = info;
    if (precision) bq_weights_kernel<dd><<<n_par, 128, 0, s>>>(p);
    else bq_weights_kernel<double><<<n_par, 128, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dev, s);
    if (e != cudaSuccess) { set_error("ssm_bq_weights: CUDA error: %s", cudaGetErrorString(e)); return SSM_E_CUDA; }
    return SSM_OK;
}

// ---- stand-alone RBF kernel evaluation and expectations (RBFGauss public methods) ----------------
namespace ssm {

struct RbfPar {
    int D, n1, n2, scaling;
    const double *par, *x1, *x2;  // device
    double *K, *q, *R, *Q, *kbar;
};

__global__ void rbf_eval_kernel(const RbfPar p) {
    // K_ij = exp(2 log(alpha) - 0.5 maha(x1_i / l, x2_j / l))                     bqkern.py:329-343
    const int D = p.D;
    const double alpha = p.scaling ? p.par[0] : 1.0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n1 * p.n2; e += gridDim.x * blockDim.x) {
        const int i = e / p.n2, j = e % p.n2;
        double a2 = 0.0, b2 = 0.0, cr = 0.0;
        for (int d = 0; d < D; ++d) {
            const double il = 1.0 / p.par[1 + d];
            const double a = il * p.x1[d * p.n1 + i], b = il * p.x2[d * p.n2 + j];
            a2 = fma(a, a, a2);
            b2 = fma(b, b, b2);
            cr = fma(a, b, cr);
        }
        p.K[e] = exp(2.0 * log(alpha) - 0.5 * ((a2 + b2) - 2.0 * cr));
    }
}

__global__ void rbf_expect_kernel(const RbfPar p) {
    const int D = p.D, N = p.n1;
    const double *x = p.x1;
    const double alpha = p.scaling ? p.par[0] : 1.0;
    double cdet = 1.0, rdet = 1.0;
    for (int d = 0; d < D; ++d) {
        const double il2 = 1.0 / (p.par[1 + d] * p.par[1 + d]);
        cdet *= il2 + 1.0;
        rdet *= 2.0 * il2 + 1.0;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double s = 0.0, s1 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double l = p.par[1 + d];
            s = fma(x[d * N + i], (1.0 / (l * l + 1.0)) * x[d * N + i], s);
            s1 = fma(x[d * N + i], (1.0 / (l * l + 1.0)) * x[d * N + i], s1);
        }
        p.q[i] = alpha * alpha / sqrt(cdet) * exp(-0.5 * s);
        const double q1 = 1.0 / sqrt(cdet) * exp(-0.5 * s1);  // exp_x_xkx always uses the unscaled q (bqkern.py:362)
        for (int d = 0; d < D; ++d) {
            const double l = p.par[1 + d];
            p.R[d * N + i] = q1 * ((1.0 / (l * l + 1.0)) * x[d * N + i]);
        }
    }
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int i = e / N, j = e % N;
        double xi = 0.0, xj = 0.0, mh = 0.0;
        for (int d = 0; d < D; ++d) {
            const double il = 1.0 / p.par[1 + d], il2 = il * il;
            xi = fma(il * x[d * N + i], il * x[d * N + i], xi);
            xj = fma(il * x[d * N + j], il * x[d * N + j], xj);
            const double a = il2 * x[d * N + i] + il2 * x[d * N + j];
            mh = fma(a * a, 1.0 / (2.0 * il2 + 1.0), mh);
        }
        const double n = (2.0 * log(alpha) - 0.5 * xi) + (2.0 * log(alpha) - 0.5 * xj) + 0.5 * mh;
        p.Q[e] = 1.0 / sqrt(rdet) * exp(n);
    }
    if (threadIdx.x == 0) p.kbar[0] = p.par[0] * p.par[0] / sqrt(rdet);
}

}  // namespace ssm

extern "C" int ssm_rbf_eval(int32_t dim, int32_t n1, int32_t n2, const double *par, const double *x1, const double *x2,
                            int32_t scaling, double *K, void *stream) {
    if (!par || !x1 || !K || dim < 1 || dim > W_MAXD || n1 < 1 || n2 < 1) { set_error("ssm_rbf_eval: bad arguments"); return SSM_E_INVALID; }
    if (!x2) { x2 = x1; n2 = n1; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t cnt = (size_t)(dim + 1) + (size_t)dim * n1 + (size_t)dim * n2;
    std::vector<double> host(cnt);
    memcpy(host.data(), par, (dim + 1) * sizeof(double));
    memcpy(host.data() + dim + 1, x1, (size_t)dim * n1 * sizeof(double));
    memcpy(host.data() + dim + 1 + (size_t)dim * n1, x2, (size_t)dim * n2 * sizeof(double));
    double *dev = nullptr;
    if (scratch_alloc((void **)&dev, cnt * sizeof(double), s) != cudaSuccess) { set_error("ssm_rbf_eval: cudaMallocAsync failed"); return SSM_E_CUDA; }
    cudaMemcpyAsync(dev, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, s);
    RbfPar p{dim, n1, n2, scaling, dev, dev + dim + 1, dev + dim + 1 + (size_t)dim * n1, K, nullptr, nullptr, nullptr, nullptr};
    const int total = n1 * n2;
    rbf_eval_kernel<<<(total + 127) / 128 > 1024 ? 1024 : (total + 127) / 128, 128, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dev, s);
    if (e != cudaSuccess) { set_error("ssm_rbf_eval: CUDA error: %s", cudaGetErrorString(e)); return SSM_E_CUDA; }
    return SSM_OK;
}

extern "C" int ssm_rbf_expectations(int32_t dim, int32_t n_pts, const double *par, const double *points, int32_t scaling,
                                    double *q, double *R, double *Q, double *kbar, void *stream) {
    if (!par || !points || !q || !R || !Q || !kbar || dim < 1 || dim > W_MAXD || n_pts < 1) { set_error("ssm_rbf_expectations: bad arguments"); return SSM_E_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t cnt = (size_t)(dim + 1) + (size_t)dim * n_pts;
    std::vector<double> host(cnt);
    memcpy(host.data(), par, (dim + 1) * sizeof(double));
    memcpy(host.data() + dim + 1, points, (size_t)dim * n_pts * sizeof(double));
    double *dev = nullptr;
    if (scratch_alloc((void **)&dev, cnt * sizeof(double), s) != cudaSuccess) { set_error("ssm_rbf_expectations: cudaMallocAsync failed"); return SSM_E_CUDA; }
    cudaMemcpyAsync(dev, host.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, s);
    RbfPar p{dim, n_pts, n_pts, scaling, dev, dev + dim + 1, dev + dim + 1, nullptr, q, R, Q, kbar};
    rbf_expect_kernel<<<1, 128, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dev, s);
    if (e != cudaSuccess) { set_error("ssm_rbf_expectations: CUDA error: %s", cudaGetErrorString(e)); return SSM_E_CUDA; }
    return SSM_OK;
}
