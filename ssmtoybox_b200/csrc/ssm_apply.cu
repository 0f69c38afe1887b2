// Stand-alone moment transform and model-function evaluation (set-up / interactive path).
// Replaces MomentTransform.apply as a public call -- SigmaPointTransform.apply (mtran.py:105-149),
// BQTransform.apply (bq/bqmtran.py:60-109) -- for a batch of (mean, cov) pairs, and single-point
// evaluations TransitionModel.dyn_fcn / dyn_eval and MeasurementModel.meas_fcn / meas_eval
// (ssmod.py:129-166, 960-1009).  Uses the runtime-N code path of the filter kernel (weights in global
// memory): these calls are latency-bound, the fused forward pass is the fast path.
#include "ssm_filter_dispatch.cuh"

namespace ssm {

void set_error(const char *fmt, ...);

struct FnDynUngm { static constexpr int D = 1, E = 1, NQ = 1; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[1], const double (&n)[1], double t, double (&o)[1]) { DynUngm::f<NZ>(p, x, n, t, o); } };
struct FnDynPend { static constexpr int D = 2, E = 2, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[2], const double (&n)[2], double t, double (&o)[2]) { DynPendulum::f<NZ>(p, x, n, t, o); } };
struct FnDynReentry { static constexpr int D = 5, E = 5, NQ = 3; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[5], const double (&n)[3], double t, double (&o)[5]) { DynReentry::f<NZ>(p, x, n, t, o); } };
struct FnDynCt { static constexpr int D = 5, E = 5, NQ = 5; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[5], const double (&n)[5], double t, double (&o)[5]) { DynCoordTurn::f<NZ>(p, x, n, t, o); } };
struct FnDynReentry1D { static constexpr int D = 3, E = 3, NQ = 3; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[3], const double (&n)[3], double t, double (&o)[3]) { DynReentry1D::f<NZ>(p, x, n, t, o); } };
struct FnObsRange { static constexpr int D = 3, E = 1, NQ = 1; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[3], const double (&n)[1], double t, double (&o)[1]) { ObsRange<3, 0>::h<NZ>(p, x, n, t, o); } };
struct FnDynCv { static constexpr int D = 4, E = 4, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[4], const double (&n)[2], double t, double (&o)[4]) { DynConstVel::f<NZ>(p, x, n, t, o); } };
struct FnObsRadar4_01 { static constexpr int D = 4, E = 2, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[4], const double (&n)[2], double t, double (&o)[2]) { ObsRadar<4, 0, 1>::h<NZ>(p, x, n, t, o); } };
struct FnObsRadar4_02 { static constexpr int D = 4, E = 2, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[4], const double (&n)[2], double t, double (&o)[2]) { ObsRadar<4, 0, 2>::h<NZ>(p, x, n, t, o); } };
struct FnObsBearing02 { static constexpr int D = 5, E = 4, NQ = 4; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[5], const double (&n)[4], double t, double (&o)[4]) { ObsBearing4<5, 0, 2>::h<NZ>(p, x, n, t, o); } };
struct FnDynCtrs { static constexpr int D = 5, E = 5, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[5], const double (&n)[2], double t, double (&o)[5]) { DynCtrs::f<NZ>(p, x, n, t, o); } };
struct FnDynCtrsAug { static constexpr int D = 7, E = 5, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&xq)[7], const double (&)[2], double t, double (&o)[5]) { const double x[5] = {xq[0], xq[1], xq[2], xq[3], xq[4]}, q[2] = {xq[5], xq[6]}; DynCtrs::f<true>(p, x, q, t, o); } };
// non-additive models: plain (state, noise) form for ssm_model_eval, augmented form [x; noise] for ssm_transform_apply
struct FnDynUngmNA { static constexpr int D = 1, E = 1, NQ = 1; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[1], const double (&n)[1], double t, double (&o)[1]) { DynUngmNA::f<NZ>(p, x, n, t, o); } };
struct FnObsUngmNA { static constexpr int D = 1, E = 1, NQ = 1; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[1], const double (&n)[1], double t, double (&o)[1]) { ObsUngmNA<1, 0>::h<NZ>(p, x, n, t, o); } };
struct FnDynUngmNAaug { static constexpr int D = 2, E = 1, NQ = 1; static constexpr bool EXACT = true; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&xq)[2], const double (&)[1], double t, double (&o)[1]) { const double x[1] = {xq[0]}, q[1] = {xq[1]}; DynUngmNA::f<true>(p, x, q, t, o); } };
struct FnObsUngmNAaug { static constexpr int D = 2, E = 1, NQ = 1; static constexpr bool EXACT = true; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&xr)[2], const double (&)[1], double t, double (&o)[1]) { const double x[1] = {xr[0]}, r[1] = {xr[1]}; ObsUngmNA<1, 0>::h<true>(p, x, r, t, o); } };
struct FnObsUngm { static constexpr int D = 1, E = 1, NQ = 1; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[1], const double (&n)[1], double t, double (&o)[1]) { ObsUngm<1, 0>::h<NZ>(p, x, n, t, o); } };
struct FnObsPend { static constexpr int D = 2, E = 1, NQ = 1; static constexpr bool EXACT = true; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[2], const double (&n)[1], double t, double (&o)[1]) { ObsPendulum<2, 0>::h<NZ>(p, x, n, t, o); } };
struct FnObsRadar01 { static constexpr int D = 5, E = 2, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[5], const double (&n)[2], double t, double (&o)[2]) { ObsRadar<5, 0, 1>::h<NZ>(p, x, n, t, o); } };
struct FnObsRadar02 { static constexpr int D = 5, E = 2, NQ = 2; static constexpr bool EXACT = false; template <bool NZ> SSM_DEV static void ev(const double *p, const double (&x)[5], const double (&n)[2], double t, double (&o)[2]) { ObsRadar<5, 0, 2>::h<NZ>(p, x, n, t, o); } };

template <int D, int E>
struct ApplyPar {
    TfGlobal<D, E> tf;
    double par[8];
    double time;
    const double *mean, *cov;
    double *mean_f, *cov_f, *cov_fx;
    int32_t *status;
    long long n, ld;
    // per-column weight sets (ssm_transform_apply_batched): column t reads wm + t N, Wc + t N^2, Wcc + t D N and adds
    // model_var[t] I to the covariance; batched = 0: one weight set for all columns
    int batched;
    const double *mv_col;
};

template <class Fn, int KIND>
__global__ void __launch_bounds__(64) apply_kernel(const __grid_constant__ ApplyPar<Fn::D, Fn::E> p) {
    constexpr int D = Fn::D, E = Fn::E;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    TfGlobal<D, E> tf = p.tf;
    if (p.batched) {
        const long long N = tf.n;
        tf.wm_ += t * N;
        tf.Wc_ += t * N * N;
        tf.Wcc_ += t * D * N;
        if (p.mv_col) {
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b < E; ++b) tf.mv_[a][b] = (a == b) ? p.mv_col[t] : 0.0;
        }
    }
    double m[D], P[TriSize<D>::value], mf[E], Cf[TriSize<E>::value], Cfx[E][D];
#pragma unroll
    for (int a = 0; a < D; ++a) m[a] = p.mean[(long long)a * p.ld + t];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) P[tri(r, c)] = p.cov[(long long)(r * D + c) * p.ld + t];
    const bool ok = moment_transform<D, E, PTS_GENERIC, 0, KIND, 0, Fn::EXACT>(
        tf, m, P,
        [&](const double (&x)[D], double (&o)[E]) {
            const double z[Fn::NQ] = {};
            Fn::template ev<false>(p.par, x, z, p.time, o);
        },
        mf, Cf, true,
        [&](int a, int c, double v) { Cfx[a][c] = v; },
        nullptr);
#pragma unroll
    for (int a = 0; a < E; ++a) p.mean_f[(long long)a * p.ld + t] = ok ? mf[a] : qnan();
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
        for (int c = 0; c < E; ++c) p.cov_f[(long long)(r * E + c) * p.ld + t] = ok ? Cf[sym(r, c)] : qnan();
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) p.cov_fx[(long long)(r * D + c) * p.ld + t] = ok ? Cfx[r][c] : qnan();
    p.status[t] = ok ? 0 : ((1 << 8) | SSM_FAIL_CHOL_DYN);
}

template <class Fn, int KIND>
static int launch_apply(const ssm_transform &tf, const double *par, double time, const double *mean, const double *cov,
                        double *mean_f, double *cov_f, double *cov_fx, int32_t *status, long long n, long long ld, cudaStream_t s) {
    constexpr int D = Fn::D, E = Fn::E;
    if (tf.dim_in != D || tf.dim_out != E) { set_error("ssm_transform_apply: transform is %dx%d, model function is %dx%d", tf.dim_in, tf.dim_out, D, E); return SSM_E_INVALID; }
    if (!tf_global_fits(tf)) {
        set_error("ssm_transform_apply: at most %d points (sigma-point rules: %d), got %d", GEN_CAP, GEN_CAP_STREAM, tf.n_pts);
        return SSM_E_UNSUPPORTED;
    }
    const size_t cnt = tf_global_count(tf, D);
    double *host = (double *)malloc(cnt * sizeof(double)), *dev = nullptr;
    if (scratch_alloc((void **)&dev, cnt * sizeof(double), s) != cudaSuccess) { free(host); return SSM_E_CUDA; }
    ApplyPar<D, E> p;
    memset(&p, 0, sizeof(p));
    size_t off = 0;
    HostTfInfo gi{PTS_GENERIC, 0.0};
    fill_tf_global(p.tf, tf, gi, dev, host, off);   // same staging as the generic filter path
    cudaMemcpyAsync(dev, host, off * sizeof(double), cudaMemcpyHostToDevice, s);
    for (int i = 0; i < 8; ++i) p.par[i] = par ? par[i] : 0.0;
    p.time = time; p.mean = mean; p.cov = cov; p.mean_f = mean_f; p.cov_f = cov_f; p.cov_fx = cov_fx; p.status = status;
    p.n = n; p.ld = ld;
    apply_kernel<Fn, KIND><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dev, s);
    free(host);
    return e == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

// BQ transform with one weight set PER COLUMN, all of them already on the device (the (n, N) / (n, N, N) / (n, D, N) outputs
// of ssm_bq_weights): the moment transforms of MarginalInference, where every (trajectory, parameter vector) pair has
// its own kernel parameters (ssinf.py:1107-1185)
template <class Fn>
static int launch_apply_batched(int n_pts, const double *points, const double *wm, const double *Wc, const double *Wcc,
                                const double *mv_col, const double *par, double time, const double *mean, const double *cov,
                                double *mean_f, double *cov_f, double *cov_fx, int32_t *status, long long n, long long ld, cudaStream_t s) {
    constexpr int D = Fn::D, E = Fn::E;
    if (n_pts < 1 || n_pts > GEN_CAP) { set_error("ssm_transform_apply_batched: 1 .. %d points, got %d", GEN_CAP, n_pts); return SSM_E_UNSUPPORTED; }
    double *dev = nullptr;
    if (scratch_alloc((void **)&dev, (size_t)D * n_pts * sizeof(double), s) != cudaSuccess) return SSM_E_CUDA;
    cudaMemcpyAsync(dev, points, (size_t)D * n_pts * sizeof(double), cudaMemcpyHostToDevice, s);   // pageable source: staged before return
    ApplyPar<D, E> p;
    memset(&p, 0, sizeof(p));
    p.tf.n = n_pts;
    p.tf.wm_ = wm; p.tf.wc_ = wm; p.tf.Wc_ = Wc; p.tf.Wcc_ = Wcc; p.tf.iK_ = Wc; p.tf.U_ = dev;
    p.batched = 1;
    p.mv_col = mv_col;
    for (int i = 0; i < 8; ++i) p.par[i] = par ? par[i] : 0.0;
    p.time = time; p.mean = mean; p.cov = cov; p.mean_f = mean_f; p.cov_f = cov_f; p.cov_fx = cov_fx; p.status = status;
    p.n = n; p.ld = ld;
    apply_kernel<Fn, SSM_TF_BQ><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dev, s);
    return e == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

template <class Fn>
static int apply_kind(const ssm_transform &tf, const double *par, double time, const double *mean, const double *cov,
                      double *mean_f, double *cov_f, double *cov_fx, int32_t *status, long long n, long long ld, cudaStream_t s) {
    switch (tf.kind) {
        case SSM_TF_SP: return launch_apply<Fn, SSM_TF_SP>(tf, par, time, mean, cov, mean_f, cov_f, cov_fx, status, n, ld, s);
        case SSM_TF_BQ: return launch_apply<Fn, SSM_TF_BQ>(tf, par, time, mean, cov, mean_f, cov_f, cov_fx, status, n, ld, s);
        case SSM_TF_TP: return launch_apply<Fn, SSM_TF_TP>(tf, par, time, mean, cov, mean_f, cov_f, cov_fx, status, n, ld, s);
    }
    set_error("ssm_transform_apply: unknown transform kind %d", tf.kind);
    return SSM_E_INVALID;
}

struct Par4 { double v[8]; };
template <class Fn>
__global__ void eval_kernel_v(const Par4 par, double time, const double *x, const double *noise, double *out, long long n, long long ld) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double xv[Fn::D], nz[Fn::NQ], o[Fn::E];
#pragma unroll
    for (int a = 0; a < Fn::D; ++a) xv[a] = x[(long long)a * ld + t];
#pragma unroll
    for (int a = 0; a < Fn::NQ; ++a) nz[a] = noise ? noise[(long long)a * ld + t] : 0.0;
    Fn::template ev<true>(par.v, xv, nz, time, o);
#pragma unroll
    for (int a = 0; a < Fn::E; ++a) out[(long long)a * ld + t] = o[a];
}

}  // namespace ssm

using namespace ssm;

// which: 0 = dynamics (model = SSM_DYN_*), 1 = measurement (model = SSM_OBS_*, with state_index)
#define SSM_FN_DISPATCH(CALL)                                                                          \
    if (which == 0 && model == SSM_DYN_UNGM) { CALL(FnDynUngm) }                                       \
    else if (which == 0 && model == SSM_DYN_PENDULUM) { CALL(FnDynPend) }                              \
    else if (which == 0 && model == SSM_DYN_REENTRY) { CALL(FnDynReentry) }                            \
    else if (which == 0 && model == SSM_DYN_COORDTURN) { CALL(FnDynCt) }                               \
    else if (which == 0 && model == SSM_DYN_REENTRY1D) { CALL(FnDynReentry1D) }                        \
    else if (which == 1 && model == SSM_OBS_RANGE && dim_state == 3 && si0 == 0) { CALL(FnObsRange) }  \
    else if (which == 1 && model == SSM_OBS_UNGM && dim_state == 1) { CALL(FnObsUngm) }                \
    else if (which == 0 && model == SSM_DYN_CONSTVEL) { CALL(FnDynCv) }                                \
    else if (which == 1 && model == SSM_OBS_RADAR && dim_state == 4 && si0 == 0 && si1 == 1) { CALL(FnObsRadar4_01) } \
    else if (which == 1 && model == SSM_OBS_RADAR && dim_state == 4 && si0 == 0 && si1 == 2) { CALL(FnObsRadar4_02) } \
    else if (which == 1 && model == SSM_OBS_BEARING && dim_state == 5 && si0 == 0 && si1 == 2) { CALL(FnObsBearing02) } \
    else if (which == 1 && model == SSM_OBS_PENDULUM && dim_state == 2) { CALL(FnObsPend) }            \
    else if (which == 1 && model == SSM_OBS_RADAR && dim_state == 5 && si0 == 0 && si1 == 1) { CALL(FnObsRadar01) } \
    else if (which == 1 && model == SSM_OBS_RADAR && dim_state == 5 && si0 == 0 && si1 == 2) { CALL(FnObsRadar02) } \
    else { set_error("no device implementation for which=%d model=%d dim_state=%d state_index=(%d,%d)", which, model, dim_state, si0, si1); return SSM_E_UNSUPPORTED; }

extern "C" int ssm_transform_apply(int32_t which, int32_t model, int32_t dim_state, int32_t si0, int32_t si1,
                                   const double *par, const ssm_transform *tf, double time, const double *mean,
                                   const double *cov, double *mean_f, double *cov_f, double *cov_fx, int32_t *status,
                                   int64_t n, int64_t ld, void *stream) {
    if (!tf || !mean || !cov || !mean_f || !cov_f || !cov_fx || !status) { set_error("ssm_transform_apply: NULL argument"); return SSM_E_INVALID; }
    if (!tf->points || !tf->wm || !tf->Wc || (tf->kind != SSM_TF_SP && !tf->Wcc) || (tf->kind == SSM_TF_TP && (!tf->iK || !tf->model_var))) {
        set_error("ssm_transform_apply: incomplete transform description");
        return SSM_E_INVALID;
    }
    if (n <= 0) return SSM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = SSM_OK;
#define CALL(FN) rc = apply_kind<FN>(*tf, par, time, mean, cov, mean_f, cov_f, cov_fx, status, n, ld, s);
    if (which == 0 && model == SSM_DYN_UNGMNA) { CALL(FnDynUngmNAaug) }          // mean / cov of [x; q] (ssmod.py:158-160)
    else if (which == 0 && model == SSM_DYN_CTRS) { CALL(FnDynCtrsAug) }
    else if (which == 1 && model == SSM_OBS_UNGMNA && dim_state == 1) { CALL(FnObsUngmNAaug) }
    else SSM_FN_DISPATCH(CALL)
#undef CALL
    if (rc == SSM_E_CUDA) set_error("ssm_transform_apply: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

extern "C" int ssm_transform_apply_batched(int32_t which, int32_t model, int32_t dim_state, int32_t si0, int32_t si1,
                                           const double *par, int32_t n_pts, const double *points, const double *wm,
                                           const double *Wc, const double *Wcc, const double *model_var, double time,
                                           const double *mean, const double *cov, double *mean_f, double *cov_f,
                                           double *cov_fx, int32_t *status, int64_t n, int64_t ld, void *stream) {
    if (!points || !wm || !Wc || !Wcc || !mean || !cov || !mean_f || !cov_f || !cov_fx || !status) {
        set_error("ssm_transform_apply_batched: NULL argument");
        return SSM_E_INVALID;
    }
    if (n <= 0) return SSM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = SSM_OK;
#define CALL(FN) rc = launch_apply_batched<FN>(n_pts, points, wm, Wc, Wcc, model_var, par, time, mean, cov, mean_f, cov_f, cov_fx, status, n, ld, s);
    SSM_FN_DISPATCH(CALL)
#undef CALL
    if (rc == SSM_E_CUDA) set_error("ssm_transform_apply_batched: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

extern "C" int ssm_model_eval(int32_t which, int32_t model, int32_t dim_state, int32_t si0, int32_t si1, const double *par,
                              double time, const double *x, const double *noise, double *out, int64_t n, int64_t ld,
                              void *stream) {
    if (!x || !out) { set_error("ssm_model_eval: NULL argument"); return SSM_E_INVALID; }
    if (n <= 0) return SSM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    Par4 p4;
    for (int i = 0; i < 8; ++i) p4.v[i] = par ? par[i] : 0.0;
#define CALL(FN) eval_kernel_v<FN><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p4, time, x, noise, out, (long long)n, (long long)ld);
    if (which == 0 && model == SSM_DYN_UNGMNA) { CALL(FnDynUngmNA) }
    else if (which == 0 && model == SSM_DYN_CTRS) { CALL(FnDynCtrs) }
    else if (which == 1 && model == SSM_OBS_UNGMNA && dim_state == 1) { CALL(FnObsUngmNA) }
    else SSM_FN_DISPATCH(CALL)
#undef CALL
    if (cudaGetLastError() != cudaSuccess) { set_error("ssm_model_eval: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError())); return SSM_E_CUDA; }
    return SSM_OK;
}
