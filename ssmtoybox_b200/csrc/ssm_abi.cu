// extern "C" entry points (include/ssm_b200.h): argument validation and model dispatch.
#include <stdarg.h>
#include <stdio.h>

#include "ssm_filter_dispatch.cuh"

namespace ssm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int filter_ungm(const FilterLaunch &L);
int filter_pendulum(const FilterLaunch &L);
int filter_reentry(const FilterLaunch &L);
int filter_coordturn(const FilterLaunch &L);
int filter_reentry1d(const FilterLaunch &L);
int filter_ungmna(const FilterLaunch &L);
int filter_constvel01(const FilterLaunch &L);
int filter_constvel02(const FilterLaunch &L);
int filter_coordturn_bearing(const FilterLaunch &L);
int filter_ctrs(const FilterLaunch &L);

__global__ void __launch_bounds__(128) filter_nan_fill_kernel(const FilterBuffers b, int dx) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int from = 0x7fffffff;
    if (t < b.n_traj) {
        const int st = b.status[t];   // (failing step, 1-based) << 8 | code; failures of earlier windows start before k_lo
        if (st != 0) from = max((st >> 8) - 1, b.k_lo);
    }
    const int first = __reduce_min_sync(0xffffffffu, from);
    if (first >= b.k_hi) return;      // no failed trajectory in this warp
    const long long cs = (long long)b.n_steps * b.ld;
    const double q = __longlong_as_double(0x7ff8000000000000LL);
    for (int k = first; k < b.k_hi; ++k) {
        if (k < from) continue;
        const long long rk = (long long)k * b.ld + t;
        if (b.fi_mean) for (int c = 0; c < dx; ++c) b.fi_mean[c * cs + rk] = q;
        if (b.pr_mean) for (int c = 0; c < dx; ++c) b.pr_mean[c * cs + rk] = q;
        if (b.fi_cov) for (int c = 0; c < dx * dx; ++c) b.fi_cov[c * cs + rk] = q;
        if (b.pr_cov) for (int c = 0; c < dx * dx; ++c) b.pr_cov[c * cs + rk] = q;
        if (b.pr_xx) for (int c = 0; c < dx * dx; ++c) b.pr_xx[c * cs + rk] = q;
    }
}

int filter_nan_fill(const FilterBuffers &b, int dx, cudaStream_t stream) {
    if (b.n_traj <= 0) return SSM_OK;
    filter_nan_fill_kernel<<<(unsigned)((b.n_traj + 127) / 128), 128, 0, stream>>>(b, dx);
    return cudaGetLastError() == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

static bool tf_valid(const ssm_transform &t) {
    if (t.n_pts < 1 || !t.points || !t.wm || !t.Wc) return false;
    if (t.kind != SSM_TF_SP && t.kind != SSM_TF_BQ && t.kind != SSM_TF_TP) return false;
    if (t.kind != SSM_TF_SP && !t.Wcc) return false;
    if (t.kind == SSM_TF_TP && (!t.iK || !t.model_var)) return false;
    return true;
}

}  // namespace ssm

using namespace ssm;

extern "C" int ssm_abi_version(void) { return SSM_ABI_VERSION; }
extern "C" const char *ssm_last_error(void) { return g_err; }

extern "C" int ssm_weights_reflective(const ssm_transform *tf) {
    if (!tf || !tf->points || tf->dim_in < 1 || tf->n_pts != 2 * tf->dim_in + 1) return 0;
    if (tf->kind == SSM_TF_TP) {   // a TPQ transform runs as a BQ one with folded weights (tp_fold): the check sees those
        if (!tp_foldable(*tf) || !wc_symmetric(*tf)) return 0;
        TpFold f;
        tp_fold(*tf, f);
        return weights_reflective(f.tf, classify_points(f.tf)) ? 1 : 0;
    }
    return weights_reflective(*tf, classify_points(*tf)) ? 1 : 0;
}

struct ScoreOut {
    const double *x_truth;
    double *stats, *rmse_acc, *quad, *dres;
};

static int filter_window_impl(const ssm_desc *desc, const double *y, double *fi_mean, double *fi_cov, double *pr_mean,
                              double *pr_cov, double *pr_xx_cov, const double *init_mean, const double *init_cov,
                              double *last_mean, double *last_cov, const int32_t *t_offset, int32_t k0, int32_t *status,
                              int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream,
                              const ScoreOut *sc, bool lower_only = false) {
    if (!desc || !y || !status) { set_error("ssm_filter: desc, y and status must not be NULL"); return SSM_E_INVALID; }
    if (n_traj < 0 || n_steps < 0 || ld < n_traj) { set_error("ssm_filter: bad sizes (n_traj=%lld n_steps=%d ld=%lld)", (long long)n_traj, n_steps, (long long)ld); return SSM_E_INVALID; }
    if (k_lo < 0 || k_hi < k_lo || k_hi > n_steps) { set_error("ssm_filter: bad time window [%d, %d) of %d steps", k_lo, k_hi, n_steps); return SSM_E_INVALID; }
    if (!desc->m0 || !desc->P0 || !desc->R || (!desc->GQG && !desc->q_cov)) { set_error("ssm_filter: m0, P0, GQG (or q_cov), R must not be NULL"); return SSM_E_INVALID; }
    if ((init_mean == nullptr) != (init_cov == nullptr)) { set_error("ssm_filter: init_mean and init_cov go together"); return SSM_E_INVALID; }
    if (!tf_valid(desc->tf_dyn) || !tf_valid(desc->tf_obs)) { set_error("ssm_filter: incomplete transform description"); return SSM_E_INVALID; }
    if (desc->family != SSM_FAMILY_GAUSS && desc->family != SSM_FAMILY_STUDENT) { set_error("ssm_filter: unknown family %d", desc->family); return SSM_E_INVALID; }
    if (n_traj == 0 || k_hi == k_lo) return SSM_OK;
    FilterLaunch L;
    L.desc = desc;
    L.stream = (cudaStream_t)stream;
    L.buf = FilterBuffers{y, fi_mean, fi_cov, pr_mean, pr_cov, pr_xx_cov, init_mean, init_cov, last_mean, last_cov,
                          t_offset, status, (long long)n_traj, (long long)ld, n_steps, k0, nullptr, nullptr, 0, 0,
                          k_lo, k_hi, k_lo > 0 ? 1 : 0};
    L.buf.lower_only = lower_only ? 1 : 0;
    if (sc) {
        L.buf.x_truth = sc->x_truth; L.buf.rmse_acc = sc->rmse_acc; L.buf.quad = sc->quad; L.buf.dres = sc->dres;
        L.stats = sc->stats;
    }
    const int dm = desc->dyn_model, om = desc->obs_model;
    const int nsi = desc->n_state_index;
    const int32_t *si = desc->state_index;
    int rc = SSM_E_UNSUPPORTED;
    if (dm == SSM_DYN_UNGM && om == SSM_OBS_UNGM && desc->dx == 1 && desc->dy == 1 && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = filter_ungm(L);
    else if (dm == SSM_DYN_PENDULUM && om == SSM_OBS_PENDULUM && desc->dx == 2 && desc->dy == 1 && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = filter_pendulum(L);
    else if (dm == SSM_DYN_REENTRY && om == SSM_OBS_RADAR && desc->dx == 5 && desc->dy == 2 && (nsi == 0 || (nsi == 2 && si[0] == 0 && si[1] == 1)))
        rc = filter_reentry(L);
    else if (dm == SSM_DYN_COORDTURN && om == SSM_OBS_RADAR && desc->dx == 5 && desc->dy == 2 && nsi == 2 && si[0] == 0 && si[1] == 2)
        rc = filter_coordturn(L);
    else if (dm == SSM_DYN_REENTRY1D && om == SSM_OBS_RANGE && desc->dx == 3 && desc->dy == 1 && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = filter_reentry1d(L);
    else if (dm == SSM_DYN_UNGMNA && om == SSM_OBS_UNGMNA && desc->dx == 1 && desc->dy == 1 && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = filter_ungmna(L);
    else if (dm == SSM_DYN_CONSTVEL && om == SSM_OBS_RADAR && desc->dx == 4 && desc->dy == 2 && (nsi == 0 || (nsi == 2 && si[0] == 0 && si[1] == 1)))
        rc = filter_constvel01(L);
    else if (dm == SSM_DYN_CONSTVEL && om == SSM_OBS_RADAR && desc->dx == 4 && desc->dy == 2 && nsi == 2 && si[0] == 0 && si[1] == 2)
        rc = filter_constvel02(L);
    else if (dm == SSM_DYN_COORDTURN && om == SSM_OBS_BEARING && desc->dx == 5 && desc->dy == 4 && nsi == 2 && si[0] == 0 && si[1] == 2)
        rc = filter_coordturn_bearing(L);
    else if (dm == SSM_DYN_CTRS && om == SSM_OBS_RADAR && desc->dx == 5 && desc->dy == 2 && (nsi == 0 || (nsi == 2 && si[0] == 0 && si[1] == 1)))
        rc = filter_ctrs(L);
    else
        set_error("ssm_filter: no device implementation for dyn_model=%d obs_model=%d dx=%d dy=%d state_index(n=%d)", dm, om, desc->dx, desc->dy, nsi);
    if (rc == SSM_E_CUDA) set_error("ssm_filter: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

extern "C" int ssm_filter_window(const ssm_desc *desc, const double *y, double *fi_mean, double *fi_cov, double *pr_mean,
                                 double *pr_cov, double *pr_xx_cov, const double *init_mean, const double *init_cov,
                                 double *last_mean, double *last_cov, const int32_t *t_offset, int32_t k0, int32_t *status,
                                 int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    return filter_window_impl(desc, y, fi_mean, fi_cov, pr_mean, pr_cov, pr_xx_cov, init_mean, init_cov, last_mean, last_cov,
                              t_offset, k0, status, n_traj, n_steps, k_lo, k_hi, ld, stream, nullptr);
}

extern "C" int ssm_filter_window_lower(const ssm_desc *desc, const double *y, double *fi_mean, double *fi_cov, double *pr_mean,
                                       double *pr_cov, double *pr_xx_cov, const double *init_mean, const double *init_cov,
                                       double *last_mean, double *last_cov, const int32_t *t_offset, int32_t k0, int32_t *status,
                                       int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    return filter_window_impl(desc, y, fi_mean, fi_cov, pr_mean, pr_cov, pr_xx_cov, init_mean, init_cov, last_mean, last_cov,
                              t_offset, k0, status, n_traj, n_steps, k_lo, k_hi, ld, stream, nullptr, true);
}

extern "C" int ssm_filter_scores(const ssm_desc *desc, const double *y, const double *x_truth, double *fi_mean, double *fi_cov,
                                 double *stats, double *rmse_acc, double *quad, double *dres,
                                 const double *init_mean, const double *init_cov, double *last_mean, double *last_cov,
                                 const int32_t *t_offset, int32_t k0, int32_t *status,
                                 int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream) {
    if (!x_truth || !stats) { set_error("ssm_filter_scores: x_truth and stats must not be NULL"); return SSM_E_INVALID; }
    const ScoreOut sc{x_truth, stats, rmse_acc, quad, dres};
    return filter_window_impl(desc, y, fi_mean, fi_cov, nullptr, nullptr, nullptr, init_mean, init_cov, last_mean, last_cov,
                              t_offset, k0, status, n_traj, n_steps, k_lo, k_hi, ld, stream, &sc);
}

extern "C" int ssm_filter(const ssm_desc *desc, const double *y, double *fi_mean, double *fi_cov, double *pr_mean,
                          double *pr_cov, double *pr_xx_cov, const double *init_mean, const double *init_cov,
                          double *last_mean, double *last_cov, const int32_t *t_offset, int32_t k0, int32_t *status,
                          int64_t n_traj, int32_t n_steps, int64_t ld, void *stream) {
    return ssm_filter_window(desc, y, fi_mean, fi_cov, pr_mean, pr_cov, pr_xx_cov, init_mean, init_cov, last_mean, last_cov,
                             t_offset, k0, status, n_traj, n_steps, 0, n_steps, ld, stream);
}

// Strided host <-> device copy of a trajectory range of a [component][step][trajectory] array:
// `height` rows of `width` bytes, row pitches in bytes.  Pinned host memory -> asynchronous DMA on the stream.
extern "C" int ssm_memcpy2d(void *dst, uint64_t dpitch, const void *src, uint64_t spitch, uint64_t width, uint64_t height,
                            int32_t host_to_device, void *stream) {
    if (!dst || !src) { set_error("ssm_memcpy2d: NULL pointer"); return SSM_E_INVALID; }
    if (width == 0 || height == 0) return SSM_OK;
    const cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height,
                                            host_to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ssm_memcpy2d: CUDA error: %s", cudaGetErrorString(e)); return SSM_E_CUDA; }
    return SSM_OK;
}
