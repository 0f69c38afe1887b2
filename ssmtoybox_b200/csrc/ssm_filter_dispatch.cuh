// Host-side dispatch of one (dynamics, measurement) model pair over point-set type, transform kind
// and filter family.  Each model pair is compiled in its own translation unit (ssm_filter_*.cu) so
// the instantiations build in parallel.
#pragma once
#include <string.h>

#include <vector>

#include "ssm_filter.cuh"
#ifdef SSM_PAIR_MODEL
#include "ssm_filter_pair.cuh"
#endif


#ifndef SSM_PAIR_TPB
#define SSM_PAIR_TPB 64      // trajectories per CTA of the warp-pair kernel (CTA = 2 x TPB threads)
#endif
#ifndef SSM_PAIR_MINB
#define SSM_PAIR_MINB 4
#endif
#ifndef SSM_PAIR_DEFAULT
#define SSM_PAIR_DEFAULT 0
#endif

namespace ssm {

void set_error(const char *fmt, ...);

// SSM_PAIR=0 / 1 selects the forward-pass mapping at run time (developer switch for A/B measurements)
inline bool pair_enabled() {
    const char *e = getenv("SSM_PAIR");
    return e ? atoi(e) != 0 : (SSM_PAIR_DEFAULT != 0);
}

template <class Dyn, class Obs, int PTS, int KIND, int FAMILY, int THREADS, int MINB, bool SCORE = false>
int dispatch_npts(const FilterLaunch &L, const HostTfInfo &id, const HostTfInfo &io) {
    constexpr int D = Dyn::DX;
    constexpr int N = (PTS == PTS_AXIS_C) ? 2 * D + 1 : 2 * D;
    return launch_filter_const<Dyn, Obs, PTS, N, KIND, FAMILY, THREADS, MINB, SCORE>(L, id, io);
}

template <class Dyn, class Obs, int THREADS, int MINB>
int launch_filter_generic(const FilterLaunch &L, int kind, int family);

// TPQ on the fast path is a BQ transform with folded weights.  The reference adds the data-dependent model variance
//   mv (nu - 2 + fx K^-1 fx^T) / (nu - 2 + N)                       bqmod.py:1155-1160, bqmtran.py:414-415
// as a FULL E x E matrix whenever the transform object has dim_out = 1 (I_out is 1 x 1 and broadcasts; ssinf.py:550 builds
// every TPQ filter that way), so the covariance is one quadratic form plus a constant:
//   fx Wc fx^T - mf mf^T + c (fx K^-1 fx^T) + c (nu - 2) 11^T  =  fx (Wc + c sym(K^-1)) fx^T - mf mf^T + c (nu - 2) 11^T,
//   c = mv / (nu - 2 + N).
// The device needs the lower triangle only, and a symmetric form sees the symmetric part of K^-1 (the reference's
// cho_solve inverse is symmetric up to its own rounding).  One dense sum instead of two: the TPQ forward pass costs what
// the GPQ one does (coordinated turn, 121 952 x 500: 31.7 -> 15.9 ms).  Diagonal-only variance (dim_out = E > 1) and
// generic point sets keep the two separate sums on the runtime-N path.
struct TpFold {
    std::vector<double> Wc, mv;
    ssm_transform tf;
};
inline bool tp_foldable(const ssm_transform &tf) {
    return tf.kind == SSM_TF_TP && tf.iK && tf.model_var && tf.Wc && (tf.tp_full_matrix || tf.dim_out == 1);
}
inline void tp_fold(const ssm_transform &tf, TpFold &o) {
    const int N = tf.n_pts, E = tf.dim_out;
    const double tp_a = tf.nu - 2.0, tp_b = 1.0 / (tf.nu - 2.0 + (double)N), mv0 = tf.model_var[0];
    const double c = tp_b * mv0;
    o.Wc.resize((size_t)N * N);
    o.mv.assign((size_t)E * E, (tp_a * tp_b) * mv0);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) o.Wc[(size_t)i * N + j] = tf.Wc[i * N + j] + c * (0.5 * (tf.iK[i * N + j] + tf.iK[j * N + i]));
    o.tf = tf;
    o.tf.kind = SSM_TF_BQ;
    o.tf.Wc = o.Wc.data();
    o.tf.model_var = o.mv.data();
    o.tf.iK = nullptr;
}

template <class Dyn, class Obs, int THREADS, int MINB>
int dispatch_filter_model(const FilterLaunch &L) {
    const ssm_desc &d = *L.desc;
    const ssm_transform &a = d.tf_dyn, &b = d.tf_obs;
    // a model with non-additive noise is integrated over the augmented vector [x; noise] (ssinf.py:271-272, 282-283)
    constexpr int DD = Dyn::ADDITIVE ? Dyn::DX : Dyn::DX + Dyn::DQ, DO = Obs::ADDITIVE ? Dyn::DX : Dyn::DX + Obs::DY;
    if (a.dim_in != DD || a.dim_out != Dyn::DX || b.dim_in != DO || b.dim_out != Obs::DY) {
        set_error("transform dimensions do not match the model (dim_in = dim_state, + dim_noise when the noise is non-additive): "
                  "dynamics %d -> %d (expected %d -> %d), measurement %d -> %d (expected %d -> %d)",
                  a.dim_in, a.dim_out, DD, Dyn::DX, b.dim_in, b.dim_out, DO, Obs::DY);
        return SSM_E_INVALID;
    }
    if (a.kind != b.kind) { set_error("dynamics and measurement transforms must be of the same kind"); return SSM_E_UNSUPPORTED; }
    const HostTfInfo id = classify_points(a), io = classify_points(b);
    const int kind = a.kind, fam = d.family;
    if constexpr (Dyn::ADDITIVE && Obs::ADDITIVE) {
        if (kind == SSM_TF_TP && id.pts != PTS_GENERIC && id.pts == io.pts && a.n_pts == b.n_pts && tp_foldable(a) && tp_foldable(b) &&
            wc_symmetric(a) && wc_symmetric(b) && SSM_SYM_WC) {
            TpFold fa, fb;
            tp_fold(a, fa);
            tp_fold(b, fb);
            ssm_desc dd = d;
            dd.tf_dyn = fa.tf;
            dd.tf_obs = fb.tf;
            FilterLaunch Lf = L;
            Lf.desc = &dd;
            return dispatch_filter_model<Dyn, Obs, THREADS, MINB>(Lf);  // the parameter block is filled before the launch returns
        }
    }
    if (L.buf.x_truth) {
        // scoring forward pass (ssm_filter_scores): instantiated for additive models, UT-type point sets ([0 | cI | -cI]),
        // Gaussian family; everything else reports SSM_E_UNSUPPORTED and the caller scores the stored moments instead
        if constexpr (Dyn::ADDITIVE && Obs::ADDITIVE) {
            if (id.pts == PTS_AXIS_C && io.pts == PTS_AXIS_C && a.n_pts == b.n_pts && fam == SSM_FAMILY_GAUSS && wc_symmetric(a) && wc_symmetric(b)) {
                if (kind == SSM_TF_SP) return dispatch_npts<Dyn, Obs, PTS_AXIS_C, SSM_TF_SP, SSM_FAMILY_GAUSS, THREADS, MINB, true>(L, id, io);
                // compact sums + scoring, no predictive-moment stores: four CTAs per SM at 128 registers beat three at 168
                // (coordinated-turn BSQ sweep point, 10^6 x 100: 33.7 against 35.2 ms; with predictive moments it is
                // the other way round)
                if (kind == SSM_TF_BQ && weights_reflective(a, id) && weights_reflective(b, io))
                    return dispatch_npts<Dyn, Obs, PTS_AXIS_C, SSM_TF_BQR, SSM_FAMILY_GAUSS, THREADS, (MINB == 3 ? 4 : MINB), true>(L, id, io);
                if (kind == SSM_TF_BQ) return dispatch_npts<Dyn, Obs, PTS_AXIS_C, SSM_TF_BQ, SSM_FAMILY_GAUSS, THREADS, MINB, true>(L, id, io);
            }
        }
        set_error("ssm_filter_scores: in-kernel scoring is compiled for additive models, [0 | cI | -cI] point sets, symmetric covariance weights and the Gaussian family");
        return SSM_E_UNSUPPORTED;
    }
    if constexpr (!Dyn::ADDITIVE || !Obs::ADDITIVE) {
        // augmented transforms run on the runtime-N path (weights in global memory) only
        if (fam == SSM_FAMILY_STUDENT) { set_error("non-additive noise is implemented for the Gaussian family only"); return SSM_E_UNSUPPORTED; }
        if (!Dyn::ADDITIVE && (!d.q_mean || !d.q_cov || d.dq != Dyn::DQ)) { set_error("non-additive dynamics: q_mean, q_cov (dq = %d) required", Dyn::DQ); return SSM_E_INVALID; }
        if (!Obs::ADDITIVE && !d.r_mean) { set_error("non-additive measurement model: r_mean required"); return SSM_E_INVALID; }
        return launch_filter_generic<Dyn, Obs, THREADS, MINB>(L, kind, fam);
    } else {
    const bool same = id.pts == io.pts && a.n_pts == b.n_pts;
    // asymmetric covariance weights (never produced by the reference, bq/bqmod.py:519-521): dense rows of the runtime-N path
    // TPQ that did not fold (diagonal-only model variance with dim_out > 1): two separate sums, runtime-N path
    if (!same || id.pts == PTS_GENERIC || !wc_symmetric(a) || !wc_symmetric(b) || kind == SSM_TF_TP) return launch_filter_generic<Dyn, Obs, THREADS, MINB>(L, kind, fam);
#ifdef SSM_PAIR_MODEL
    // warp pair per 32 trajectories (ssm_filter_pair.cuh): UT-type point sets, BQ transforms, Gaussian family
    if (id.pts == PTS_AXIS_C && fam == SSM_FAMILY_GAUSS && kind == SSM_TF_BQ && pair_enabled()) {   // folded TPQ arrives as BQ
        const int rc = launch_filter_pair<Dyn, Obs, SSM_TF_BQ, SSM_PAIR_TPB, SSM_PAIR_MINB>(L, id, io);
        if (rc != SSM_E_UNSUPPORTED) return rc;
    }
#endif
    // reflection-invariant BQ weights (the package's own, symmetrised when built): compact sums, see moment_transform
    if (id.pts == PTS_AXIS_C && kind == SSM_TF_BQ && fam == SSM_FAMILY_GAUSS && weights_reflective(a, id) && weights_reflective(b, io))
        return dispatch_npts<Dyn, Obs, PTS_AXIS_C, SSM_TF_BQR, SSM_FAMILY_GAUSS, THREADS, MINB>(L, id, io);
#define SSM_CASE(P, K, F)                                                        \
    if (id.pts == P && kind == K && fam == F) return dispatch_npts<Dyn, Obs, P, K, F, THREADS, MINB>(L, id, io);
    SSM_CASE(PTS_AXIS_C, SSM_TF_SP, SSM_FAMILY_GAUSS)
    SSM_CASE(PTS_AXIS_C, SSM_TF_BQ, SSM_FAMILY_GAUSS)
    SSM_CASE(PTS_AXIS, SSM_TF_SP, SSM_FAMILY_GAUSS)
    SSM_CASE(PTS_AXIS, SSM_TF_BQ, SSM_FAMILY_GAUSS)
    SSM_CASE(PTS_AXIS_C, SSM_TF_SP, SSM_FAMILY_STUDENT)
    SSM_CASE(PTS_AXIS, SSM_TF_SP, SSM_FAMILY_STUDENT)
    // Student filters with BQ transforms on fully-symmetric degree-3 points: GPQStudent (research/tpq/tpq_base.py:41-91),
    // StudentProcessStudent (ssinf.py:778-833); other point sets take the runtime-N path
    SSM_CASE(PTS_AXIS_C, SSM_TF_BQ, SSM_FAMILY_STUDENT)
    if (fam == SSM_FAMILY_STUDENT) return launch_filter_generic<Dyn, Obs, THREADS, MINB>(L, kind, fam);
#undef SSM_CASE
    set_error("unsupported transform kind %d / family %d", kind, fam);
    return SSM_E_UNSUPPORTED;
    }
}

// ---- generic point sets: runtime N <= GEN_CAP, weights staged in device memory ------------------
template <int D, int E>
inline int fill_tf_global(TfGlobal<D, E> &o, const ssm_transform &tf, const HostTfInfo &info, double *dev, double *host,
                          size_t &off) {
    fill_tf_common(o, tf, info);
    const int N = tf.n_pts;
    auto put = [&](const double *src, size_t cnt, const double *&dst) {
        if (src) memcpy(host + off, src, cnt * sizeof(double));
        else memset(host + off, 0, cnt * sizeof(double));
        dst = dev + off;
        off += cnt;
    };
    put(tf.wm, N, o.wm_);
    {   // diagonal of Wc
        for (int i = 0; i < N; ++i) host[off + i] = tf.Wc[i * N + i];
        o.wc_ = dev + off;
        off += N;
    }
    // a sigma-point rule reads the diagonal of Wc only (staged above): no dense N x N copies for it (243^2 ... 3125^2 doubles)
    if (tf.kind == SSM_TF_SP) { o.Wc_ = o.wc_; o.Wcc_ = o.wc_; o.iK_ = o.wc_; }
    else {
        put(tf.Wc, (size_t)N * N, o.Wc_);
        put(tf.Wcc, (size_t)D * N, o.Wcc_);
        put(tf.kind == SSM_TF_TP ? tf.iK : nullptr, (size_t)N * N, o.iK_);
    }
    put(tf.points, (size_t)D * N, o.U_);
    return SSM_OK;
}

// doubles fill_tf_global stages for one transform
inline size_t tf_global_count(const ssm_transform &tf, int D) {
    const size_t N = (size_t)tf.n_pts;
    return 2 * N + (tf.kind == SSM_TF_SP ? 0 : 2 * N * N + (size_t)D * N) + (size_t)D * N;
}

// capacity of the runtime-N path: BQ / TP transforms keep their function values per thread (GEN_CAP); sigma-point
// rules beyond that are streamed (GEN_CAP_STREAM)
inline bool tf_global_fits(const ssm_transform &tf) {
    return tf.n_pts >= 1 && tf.n_pts <= (tf.kind == SSM_TF_SP ? GEN_CAP_STREAM : GEN_CAP);
}

template <class Dyn, class Obs, int KIND, int FAMILY, int THREADS, int MINB>
int launch_filter_global(const FilterLaunch &L) {
    constexpr int DX = Dyn::DX, DY = Obs::DY;
    constexpr int DD = Dyn::ADDITIVE ? DX : DX + Dyn::DQ, DO = Obs::ADDITIVE ? DX : DX + DY;  // transform input dimensions
    using TfD = TfGlobal<DD, DX>;
    using TfO = TfGlobal<DO, DY>;
    using Par = FilterPar<DX, DY, TfD, TfO>;
    const ssm_desc &d = *L.desc;
    const int Na = d.tf_dyn.n_pts, Nb = d.tf_obs.n_pts;
    if (!tf_global_fits(d.tf_dyn) || !tf_global_fits(d.tf_obs)) {
        set_error("generic point sets support at most %d points (sigma-point rules: %d); got %d / %d", GEN_CAP, GEN_CAP_STREAM, Na, Nb);
        return SSM_E_UNSUPPORTED;
    }
    if (!stride_fits<DX>(L.buf.n_steps, L.buf.ld)) {
        set_error("n_steps * ld = %lld elements per component: this model addresses components with a 32-bit stride (< 2^32); run the trajectories in chunks", (long long)L.buf.n_steps * L.buf.ld);
        return SSM_E_UNSUPPORTED;
    }
    const size_t cnt = tf_global_count(d.tf_dyn, DD) + tf_global_count(d.tf_obs, DO);
    double *host = (double *)malloc(cnt * sizeof(double));
    double *dev = nullptr;
    if (scratch_alloc((void **)&dev, cnt * sizeof(double), L.stream) != cudaSuccess) { free(host); set_error("cudaMallocAsync failed"); return SSM_E_CUDA; }
    Par p;
    memset(&p, 0, sizeof(p));
    size_t off = 0;
    HostTfInfo gi{PTS_GENERIC, 0.0};
    fill_tf_global(p.tf_dyn, d.tf_dyn, gi, dev, host, off);
    fill_tf_global(p.tf_obs, d.tf_obs, gi, dev, host, off);
    // pageable-source async copy: staged before the call returns, so `host` can be freed below
    cudaMemcpyAsync(dev, host, off * sizeof(double), cudaMemcpyHostToDevice, L.stream);
    for (int i = 0; i < 4; ++i) p.dyn_par[i] = d.dyn_par[i];
    for (int i = 0; i < 8; ++i) p.obs_par[i] = d.obs_par[i];
    for (int i = 0; i < DX; ++i) p.m0[i] = d.m0[i];
    pack_lower<DX>(d.P0, p.P0);
    pack_lower<DX>(d.GQG, p.GQG);
    pack_lower<DY>(d.R, p.R);
    if (!Dyn::ADDITIVE) {
        for (int i = 0; i < Dyn::DQ; ++i) p.q_mean[i] = d.q_mean[i];
        pack_lower<Dyn::DQ>(d.q_cov, p.q_cov);
    }
    if (!Obs::ADDITIVE)
        for (int i = 0; i < DY; ++i) p.r_mean[i] = d.r_mean[i];
    p.dof = d.dof; p.x0_dof = d.x0_dof; p.q_dof = d.q_dof; p.r_dof = d.r_dof;
    p.s0 = (d.family == SSM_FAMILY_STUDENT) ? (d.dof - 2.0) / d.dof : 1.0;
    p.fixed_dof = d.fixed_dof;
    p.b = L.buf;
    const long long blocks = (L.buf.n_traj + THREADS - 1) / THREADS;
    double *ttab = make_time_tab<Dyn>(L.buf, L.stream);
    p.b.time_tab = ttab;
    filter_kernel<Dyn, Obs, PTS_GENERIC, 0, KIND, FAMILY, Par, THREADS, MINB, false><<<(unsigned)blocks, THREADS, 0, L.stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (ttab) cudaFreeAsync(ttab, L.stream);
    if (e == cudaSuccess && filter_nan_fill(L.buf, DX, L.stream) != SSM_OK) e = cudaErrorUnknown;
    cudaFreeAsync(dev, L.stream);
    free(host);
    return e == cudaSuccess ? SSM_OK : SSM_E_CUDA;
}

template <class Dyn, class Obs, int THREADS, int MINB>
int launch_filter_generic(const FilterLaunch &L, int kind, int family) {
    if (family == SSM_FAMILY_STUDENT) {
        if (kind == SSM_TF_SP) return launch_filter_global<Dyn, Obs, SSM_TF_SP, SSM_FAMILY_STUDENT, THREADS, MINB>(L);
        if (kind == SSM_TF_BQ) return launch_filter_global<Dyn, Obs, SSM_TF_BQ, SSM_FAMILY_STUDENT, THREADS, MINB>(L);
        if (kind == SSM_TF_TP) return launch_filter_global<Dyn, Obs, SSM_TF_TP, SSM_FAMILY_STUDENT, THREADS, MINB>(L);
        set_error("unsupported transform kind %d", kind);
        return SSM_E_UNSUPPORTED;
    }
    if (kind == SSM_TF_SP) return launch_filter_global<Dyn, Obs, SSM_TF_SP, SSM_FAMILY_GAUSS, THREADS, MINB>(L);
    if (kind == SSM_TF_BQ) return launch_filter_global<Dyn, Obs, SSM_TF_BQ, SSM_FAMILY_GAUSS, THREADS, MINB>(L);
    if (kind == SSM_TF_TP) return launch_filter_global<Dyn, Obs, SSM_TF_TP, SSM_FAMILY_GAUSS, THREADS, MINB>(L);
    set_error("unsupported transform kind %d", kind);
    return SSM_E_UNSUPPORTED;
}

}  // namespace ssm
