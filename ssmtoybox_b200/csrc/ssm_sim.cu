// K1: batched simulation of state trajectories and measurements, one thread per trajectory.
// Replaces TransitionModel.simulate_discrete / simulate_continuous (ssmod.py:168-244),
// MeasurementModel.simulate_measurements (ssmod.py:1011-1039) and the samplers GaussRV.sample /
// StudentRV.sample / multivariate_t (utils.py:349-382, 618-619, 670-671).
//
// Noise is either injected (device arrays, parity mode: results follow the reference bit for bit
// up to libm) or drawn in-kernel with Philox4x32-10 keyed by (seed, GLOBAL trajectory index) and
// counted by (time step, stream id, call), so the draws do not depend on the grid shape, on the
// number of GPUs or on how trajectories are sharded.  numpy's MT19937 stream cannot be reproduced
// (SURVEY.md Q10): Philox mode is validated statistically.
#include "ssm_models.cuh"
#include "ssm_rng.cuh"

namespace ssm {

void set_error(const char *fmt, ...);

template <int DX, int DQ, int DY>
struct SimPar {
    double dyn_par[4], obs_par[8];
    double x0_mean[DX], x0_F[DX * DX], q_F[DQ * DQ], r_F[DY * DY];
    double x0_dof, q_dof, r_dof;
    unsigned long long seed;
    long long traj_offset;
    const double *x0_inj, *q_inj, *r_inj;
    double *x, *y;
    long long n_traj, ld;
    int n_steps, mode, sub, use_rng;
    double dt;
};

enum { STREAM_X0 = 0, STREAM_Q = 1, STREAM_R = 2 };

template <class Dyn, class Obs>
__global__ void __launch_bounds__(128) sim_kernel(const __grid_constant__ SimPar<Dyn::DX, Dyn::DQ, Obs::DY> p) {
    constexpr int DX = Dyn::DX, DQ = Dyn::DQ, DY = Obs::DY;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n_traj) return;
    const int N = p.n_steps;
    const long long ld = p.ld;
    Rng rng;
    rng.ph.k0 = (uint32_t)p.seed;
    rng.ph.k1 = (uint32_t)(p.seed >> 32);
    const unsigned long long gt = (unsigned long long)(p.traj_offset + t);
    rng.t_lo = (uint32_t)gt;
    rng.t_hi = (uint32_t)(gt >> 32);

    double x[DX];
    if (p.mode == SSM_SIM_MEASURE) {
    } else if (p.x0_inj) {
#pragma unroll
        for (int a = 0; a < DX; ++a) x[a] = p.x0_inj[(long long)a * ld + t];
    } else {
        double z[DX];
        rng.template draw<DX>(0, STREAM_X0, p.x0_F, p.x0_dof, z);
#pragma unroll
        for (int a = 0; a < DX; ++a) x[a] = p.x0_mean[a] + z[a];
    }
    const bool cont = p.mode == SSM_SIM_CONTINUOUS;
    const int S = cont ? (N - 1) * p.sub + 1 : N;  // internal steps
    const double qs = cont ? sqrt(p.dt) / p.dt : 1.0;  // ssmod.py:236

    auto emit = [&](int k_out, int k_int) {  // store state k_out and its measurement
        if (p.mode != SSM_SIM_MEASURE) {
#pragma unroll
            for (int a = 0; a < DX; ++a) st_stream(p.x + ((long long)a * N + k_out) * ld + t, x[a]);
        }
        if (!p.y) return;
        double r[DY], o[DY];
        if (p.r_inj) {
#pragma unroll
            for (int a = 0; a < DY; ++a) r[a] = p.r_inj[((long long)a * N + k_out) * ld + t];
        } else {
            rng.template draw<DY>((uint32_t)k_int, STREAM_R, p.r_F, p.r_dof, r);
        }
        Obs::template h<true>(p.obs_par, x, r, (double)(k_int + 1), o);  // time = k + 1, ssmod.py:1038
#pragma unroll
        for (int a = 0; a < DY; ++a) st_stream(p.y + ((long long)a * N + k_out) * ld + t, o[a]);
    };

    if (p.mode == SSM_SIM_MEASURE) {  // simulate_measurements(x) on a given state array
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int a = 0; a < DX; ++a) x[a] = ld_stream(p.x + ((long long)a * N + k) * ld + t);
            emit(k, k);
        }
        return;
    }
    if (!cont) emit(0, 0);  // discrete: slot 0 is the initial state (ssmod.py:190)
    for (int j = 1; j <= S; ++j) {
        if (!cont && j == S) break;  // discrete: steps - 1 transitions
        double q[DQ], o[DX];
        if (p.q_inj) {
            const int nq = cont ? S : N;
#pragma unroll
            for (int a = 0; a < DQ; ++a) q[a] = p.q_inj[((long long)a * nq + (j - 1)) * ld + t];
        } else {
            rng.template draw<DQ>((uint32_t)(j - 1), STREAM_Q, p.q_F, p.q_dof, q);
        }
        if (cont) {  // Euler-Maruyama, ssmod.py:238-243
#pragma unroll
            for (int a = 0; a < DQ; ++a) q[a] *= qs;
            Dyn::fc(p.dyn_par, x, q, (double)(j - 1), o);
#pragma unroll
            for (int a = 0; a < DX; ++a) x[a] = x[a] + p.dt * o[a];
            // returned array drops x0 (ssmod.py:244) and is sub-sampled [::sub] by the caller's convention
            if ((j - 1) % p.sub == 0) emit((j - 1) / p.sub, j - 1);
        } else {  // ssmod.py:198
            Dyn::template f<true>(p.dyn_par, x, q, (double)(j - 1), o);
#pragma unroll
            for (int a = 0; a < DX; ++a) x[a] = o[a];
            emit(j, j);
        }
    }
}

template <class Dyn, class Obs>
static int launch_sim(const ssm_desc *d, const ssm_rng *rng, int mode, double dt, int sub, const double *x0_inj,
                      const double *q_inj, const double *r_inj, double *x, double *y, long long n_traj, int n_steps,
                      long long ld, cudaStream_t s) {
    constexpr int DX = Dyn::DX, DQ = Dyn::DQ, DY = Obs::DY;
    if (mode == SSM_SIM_CONTINUOUS && !Dyn::HAS_CONT) {
        set_error("ssm_simulate: model %d has no continuous-time dynamics (dyn_fcn_cont)", d->dyn_model);
        return SSM_E_UNSUPPORTED;
    }
    SimPar<DX, DQ, DY> p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < 4; ++i) p.dyn_par[i] = d->dyn_par[i];
    for (int i = 0; i < 8; ++i) p.obs_par[i] = d->obs_par[i];
    const bool meas_only = mode == SSM_SIM_MEASURE;
    const bool need_rng = meas_only ? !r_inj : (!x0_inj || !q_inj || (y && !r_inj));
    if (need_rng && meas_only) {
        if (!rng || !rng->r_factor) { set_error("ssm_simulate: RNG description required when noise is not injected"); return SSM_E_INVALID; }
        for (int i = 0; i < DY * DY; ++i) p.r_F[i] = rng->r_factor[i];
        p.r_dof = rng->r_dof;
        p.seed = rng->seed;
        p.traj_offset = rng->traj_offset;
    } else if (need_rng) {
        if (!rng || !rng->x0_mean || !rng->x0_factor || !rng->q_factor || (y && !rng->r_factor)) {
            set_error("ssm_simulate: RNG description required when noise is not injected");
            return SSM_E_INVALID;
        }
        if (rng->dq != DQ) { set_error("ssm_simulate: noise dimension %d != %d", rng->dq, DQ); return SSM_E_INVALID; }
        for (int i = 0; i < DX; ++i) p.x0_mean[i] = rng->x0_mean[i];
        for (int i = 0; i < DX * DX; ++i) p.x0_F[i] = rng->x0_factor[i];
        for (int i = 0; i < DQ * DQ; ++i) p.q_F[i] = rng->q_factor[i];
        if (y) for (int i = 0; i < DY * DY; ++i) p.r_F[i] = rng->r_factor[i];
        p.x0_dof = rng->x0_dof; p.q_dof = rng->q_dof; p.r_dof = rng->r_dof;
        p.seed = rng->seed;
        p.traj_offset = rng->traj_offset;
    }
    p.x0_inj = x0_inj; p.q_inj = q_inj; p.r_inj = r_inj;
    p.x = x; p.y = y;
    p.n_traj = n_traj; p.ld = ld; p.n_steps = n_steps; p.mode = mode; p.sub = sub < 1 ? 1 : sub; p.dt = dt;
    const long long blocks = (n_traj + 127) / 128;
    sim_kernel<Dyn, Obs><<<(unsigned)blocks, 128, 0, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

}  // namespace ssm

using namespace ssm;

extern "C" int ssm_simulate(const ssm_desc *desc, const ssm_rng *rng, int32_t mode, double dt_cont, int32_t sub,
                            const double *x0_inj, const double *q_inj, const double *r_inj, double *x, double *y,
                            int64_t n_traj, int32_t n_steps, int64_t ld, void *stream) {
    if (!desc || !x) { set_error("ssm_simulate: desc and x must not be NULL"); return SSM_E_INVALID; }
    if (mode == SSM_SIM_MEASURE && !y) { set_error("ssm_simulate: y must not be NULL in measurement mode"); return SSM_E_INVALID; }
    if (mode != SSM_SIM_DISCRETE && mode != SSM_SIM_CONTINUOUS && mode != SSM_SIM_MEASURE) { set_error("ssm_simulate: bad mode %d", mode); return SSM_E_INVALID; }
    if (mode == SSM_SIM_CONTINUOUS && !(dt_cont > 0.0)) { set_error("ssm_simulate: dt must be > 0"); return SSM_E_INVALID; }
    if (n_traj < 0 || n_steps < 0 || ld < n_traj) { set_error("ssm_simulate: bad sizes"); return SSM_E_INVALID; }
    if (n_traj == 0 || n_steps == 0) return SSM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int dm = desc->dyn_model, om = desc->obs_model, nsi = desc->n_state_index;
    const int32_t *si = desc->state_index;
    int rc = SSM_E_UNSUPPORTED;
#define SSM_SIM_ARGS desc, rng, mode, dt_cont, sub, x0_inj, q_inj, r_inj, x, y, n_traj, n_steps, ld, s
    if (dm == SSM_DYN_UNGM && om == SSM_OBS_UNGM && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = launch_sim<DynUngm, ObsUngm<1, 0>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_PENDULUM && om == SSM_OBS_PENDULUM && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = launch_sim<DynPendulum, ObsPendulum<2, 0>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_REENTRY && om == SSM_OBS_RADAR && (nsi == 0 || (nsi == 2 && si[0] == 0 && si[1] == 1)))
        rc = launch_sim<DynReentry, ObsRadar<5, 0, 1>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_COORDTURN && om == SSM_OBS_RADAR && nsi == 2 && si[0] == 0 && si[1] == 2)
        rc = launch_sim<DynCoordTurn, ObsRadar<5, 0, 2>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_REENTRY1D && om == SSM_OBS_RANGE && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = launch_sim<DynReentry1D, ObsRange<3, 0>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_UNGMNA && om == SSM_OBS_UNGMNA && (nsi == 0 || (nsi == 1 && si[0] == 0)))
        rc = launch_sim<DynUngmNA, ObsUngmNA<1, 0>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_CONSTVEL && om == SSM_OBS_RADAR && (nsi == 0 || (nsi == 2 && si[0] == 0 && si[1] == 1)))
        rc = launch_sim<DynConstVel, ObsRadar<4, 0, 1>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_CONSTVEL && om == SSM_OBS_RADAR && nsi == 2 && si[0] == 0 && si[1] == 2)
        rc = launch_sim<DynConstVel, ObsRadar<4, 0, 2>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_COORDTURN && om == SSM_OBS_BEARING && nsi == 2 && si[0] == 0 && si[1] == 2)
        rc = launch_sim<DynCoordTurn, ObsBearing4<5, 0, 2>>(SSM_SIM_ARGS);
    else if (dm == SSM_DYN_CTRS && om == SSM_OBS_RADAR && (nsi == 0 || (nsi == 2 && si[0] == 0 && si[1] == 1)))
        rc = launch_sim<DynCtrs, ObsRadar<5, 0, 1>>(SSM_SIM_ARGS);
    else
        set_error("ssm_simulate: no device implementation for dyn_model=%d obs_model=%d", dm, om);
#undef SSM_SIM_ARGS
    if (rc == SSM_E_CUDA) set_error("ssm_simulate: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError()));
    return rc;
}

// ---- stand-alone sampler: GaussRV.sample / StudentRV.sample (utils.py:618-619, 670-671) ----------
namespace ssm {
constexpr int SAMPLE_MAXD = 8;
struct SamplePar {
    int dim;
    double mean[SAMPLE_MAXD], F[SAMPLE_MAXD * SAMPLE_MAXD];
    double dof;
    unsigned long long seed;
    long long offset, n, ld;
    double *out;
};
__global__ void sample_kernel(const SamplePar p) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    Rng rng;
    rng.ph.k0 = (uint32_t)p.seed;
    rng.ph.k1 = (uint32_t)(p.seed >> 32);
    const unsigned long long gt = (unsigned long long)(p.offset + t);
    rng.t_lo = (uint32_t)gt;
    rng.t_hi = (uint32_t)(gt >> 32);
    double z[SAMPLE_MAXD];
    rng.normals<SAMPLE_MAXD>(0, 3, z);
    double sc = 1.0;
    if (p.dof > 0.0) sc = rsqrt(rng.gamma(0, 3, 0.5 * p.dof) * (2.0 / p.dof));
    for (int i = 0; i < p.dim; ++i) {
        double s = 0.0;
        for (int j = 0; j < p.dim; ++j) s = fma(p.F[i * p.dim + j], z[j], s);
        p.out[(long long)i * p.ld + t] = p.mean[i] + s * sc;
    }
}
}  // namespace ssm

extern "C" int ssm_sample(int32_t dim, const double *mean, const double *factor, double dof, uint64_t seed, int64_t offset,
                          double *out, int64_t n, int64_t ld, void *stream) {
    if (!mean || !factor || !out || dim < 1 || dim > SAMPLE_MAXD || n < 0 || ld < n) { set_error("ssm_sample: bad arguments"); return SSM_E_INVALID; }
    if (n == 0) return SSM_OK;
    SamplePar p;
    memset(&p, 0, sizeof(p));
    p.dim = dim;
    for (int i = 0; i < dim; ++i) p.mean[i] = mean[i];
    for (int i = 0; i < dim * dim; ++i) p.F[i] = factor[i];
    p.dof = dof; p.seed = seed; p.offset = offset; p.n = n; p.ld = ld; p.out = out;
    sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
    if (cudaGetLastError() != cudaSuccess) { set_error("ssm_sample: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError())); return SSM_E_CUDA; }
    return SSM_OK;
}

// ---- Gaussian-mixture sampler: utils.gauss_mixture (utils.py:261-301), GaussianMixtureRV (research/tpq/tpq_base.py:13-31) ----
// One thread per sample: component index from one uniform against the cumulative mixing proportions, then
// mean_k + F_k z.  (The reference draws the component counts first, fills the components block by block and shuffles;
// the joint distribution of (sample, index) is the same.)
namespace ssm {
constexpr int MIX_MAXK = 4;
struct MixturePar {
    int dim, n_comp;
    double mean[MIX_MAXK][SAMPLE_MAXD], F[MIX_MAXK][SAMPLE_MAXD * SAMPLE_MAXD], cum[MIX_MAXK];
    unsigned long long seed;
    long long offset, n, ld;
    double *out;
    int32_t *idx;
};
__global__ void sample_mixture_kernel(const __grid_constant__ MixturePar p) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    Rng rng;
    rng.ph.k0 = (uint32_t)p.seed;
    rng.ph.k1 = (uint32_t)(p.seed >> 32);
    const unsigned long long gt = (unsigned long long)(p.offset + t);
    rng.t_lo = (uint32_t)gt;
    rng.t_hi = (uint32_t)(gt >> 32);
    uint32_t r[4];
    rng.ph.gen(rng.t_lo, rng.t_hi, 0, (5u << 16) | 63u, r);
    const double u = u01(r[0], r[1]);
    int k = 0;
    while (k < p.n_comp - 1 && u >= p.cum[k]) ++k;
    double z[SAMPLE_MAXD];
    rng.normals<SAMPLE_MAXD>(0, 5, z);
    for (int i = 0; i < p.dim; ++i) {
        double s = 0.0;
        for (int j = 0; j < p.dim; ++j) s = fma(p.F[k][i * p.dim + j], z[j], s);
        p.out[(long long)i * p.ld + t] = p.mean[k][i] + s;
    }
    if (p.idx) p.idx[t] = k;
}
}  // namespace ssm

extern "C" int ssm_sample_mixture(int32_t dim, int32_t n_comp, const double *means, const double *factors, const double *alphas,
                                  uint64_t seed, int64_t offset, double *out, int32_t *idx, int64_t n, int64_t ld, void *stream) {
    if (!means || !factors || !alphas || !out || dim < 1 || dim > SAMPLE_MAXD || n_comp < 1 || n_comp > MIX_MAXK || n < 0 || ld < n) {
        set_error("ssm_sample_mixture: bad arguments (dim <= %d, components <= %d)", SAMPLE_MAXD, MIX_MAXK);
        return SSM_E_INVALID;
    }
    if (n == 0) return SSM_OK;
    MixturePar p;
    memset(&p, 0, sizeof(p));
    p.dim = dim; p.n_comp = n_comp;
    double c = 0.0;
    for (int k = 0; k < n_comp; ++k) {
        if (!(alphas[k] >= 0.0)) { set_error("ssm_sample_mixture: negative mixing proportion"); return SSM_E_INVALID; }
        for (int i = 0; i < dim; ++i) p.mean[k][i] = means[k * dim + i];
        for (int i = 0; i < dim * dim; ++i) p.F[k][i] = factors[k * dim * dim + i];
        c += alphas[k];
        p.cum[k] = c;
    }
    if (!(fabs(c - 1.0) < 1e-9)) { set_error("ssm_sample_mixture: mixing proportions must sum to one (got %g)", c); return SSM_E_INVALID; }
    p.seed = seed; p.offset = offset; p.n = n; p.ld = ld; p.out = out; p.idx = idx;
    sample_mixture_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
    if (cudaGetLastError() != cudaSuccess) { set_error("ssm_sample_mixture: CUDA error: %s", cudaGetErrorString(cudaPeekAtLastError())); return SSM_E_CUDA; }
    return SSM_OK;
}
