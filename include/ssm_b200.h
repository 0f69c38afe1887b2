/*
 * ssm_b200.h -- C ABI of the B200-native (sm_100a) Monte-Carlo sigma-point / Bayesian-quadrature
 * Kalman filtering library behind the SSMToybox Python API.
 *
 * The reference (jacobnzw/SSMToybox, pure Python) has no FFI; the entry points below are what a
 * binding for its hot path would call.  Each one names the reference interface it replaces
 * (file:line relative to /root/reference/ssmtoybox/).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Descriptor structs and the small arrays they point to live in
 *    HOST memory and are copied at call time; all bulk buffers are DEVICE pointers to fp64 data
 *    allocated by the caller.  The library never allocates persistent memory and never frees
 *    caller memory.
 *  - Bulk layout is structure-of-arrays  [component][time step][trajectory]  with the trajectory
 *    index fastest and a caller-supplied leading dimension ld >= n_traj (elements), i.e. element
 *    (c, k, t) of an array with C components and N steps lives at  (c * N + k) * ld + t.
 *    Matrices are flattened row-major into the component index (c = row * dim + col).
 *    This is exactly a C-contiguous numpy array of shape (dim, N, M) / (dim, dim, N, M) -- the
 *    shapes the reference's research drivers assemble (research/gpq/icinco_demo.py:115-123).
 *  - Every call is asynchronous on the given CUDA stream (cudaStream_t passed as void*; NULL =
 *    default stream) and re-entrant; there is no global mutable state.
 *  - Return value: 0 on success, negative SSM_E_* on invalid arguments / unsupported
 *    configurations / CUDA errors (ssm_last_error() gives a thread-local message).  Numerical
 *    failures of individual trajectories are NOT errors: they are reported in status[traj]
 *    (0 = ok, otherwise (k << 8) | code, k = 1-based time step, code = SSM_FAIL_*), the
 *    trajectory is frozen and its remaining outputs are filled with NaN.  The reference raises
 *    numpy.linalg.LinAlgError / ValueError at the same places (mtran.py:139, bq/bqmtran.py:98,
 *    ssinf.py:321, 342).
 */
#ifndef SSM_B200_H
#define SSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSM_ABI_VERSION 5

/* ---- error codes ------------------------------------------------------------------------- */
#define SSM_OK             0
#define SSM_E_INVALID     -1   /* bad argument (null pointer, size, dimension mismatch)           */
#define SSM_E_UNSUPPORTED -2   /* model / transform combination without a device implementation  */
#define SSM_E_CUDA        -3   /* CUDA runtime error, see ssm_last_error()                        */

/* ---- per-trajectory numerical failure codes (low byte of status[]) ------------------------- */
#define SSM_FAIL_CHOL_DYN        1  /* input covariance of the dynamics transform not PD  (mtran.py:139, bqmtran.py:98) */
#define SSM_FAIL_CHOL_OBS        2  /* input covariance of the measurement transform not PD                           */
#define SSM_FAIL_CHOL_GAIN       3  /* measurement covariance not PD in cho_factor        (ssinf.py:321, 724)         */
#define SSM_FAIL_NONFINITE_GAIN  4  /* inf/NaN reaching cho_factor/cho_solve (scipy ValueError, ssinf.py:321)        */
#define SSM_FAIL_CHOL_SMOOTH     5  /* predictive covariance not PD in the smoother       (ssinf.py:342)              */

/* ---- model ids (ssmod.py) ------------------------------------------------------------------ */
#define SSM_DYN_UNGM      1  /* UNGMTransition.dyn_fcn              ssmod.py:268-269  par: -                      */
#define SSM_DYN_PENDULUM  2  /* Pendulum2DTransition.dyn_fcn        ssmod.py:357-358  par[0] = dt                 */
#define SSM_DYN_REENTRY   3  /* ReentryVehicle2DTransition.dyn_fcn  ssmod.py:530-564  par[0] = dt                 */
#define SSM_DYN_COORDTURN 4  /* CoordinatedTurnTransition.dyn_fcn   ssmod.py:675-690  par[0] = dt                 */
#define SSM_DYN_REENTRY1D 5  /* ReentryVehicle1DTransition.dyn_fcn  ssmod.py:418-421  par[0] = dt                 */
#define SSM_DYN_UNGMNA    6  /* UNGMNATransition.dyn_fcn (non-additive noise) ssmod.py:299-300                  */
#define SSM_DYN_CONSTVEL  7  /* ConstantVelocity.dyn_fcn            ssmod.py:839-846  par[0] = dt                 */
#define SSM_DYN_CTRS      8  /* ConstantTurnRateSpeed.dyn_fcn (non-additive noise) ssmod.py:755-774 par[0] = dt   */
#define SSM_OBS_UNGM      1  /* UNGMMeasurement.meas_fcn            ssmod.py:1060-1061                            */
#define SSM_OBS_PENDULUM  2  /* Pendulum2DMeasurement.meas_fcn      ssmod.py:1114-1115                            */
#define SSM_OBS_RADAR     3  /* Radar2DMeasurement.meas_fcn         ssmod.py:1227-1252 par[0..1] = radar_loc      */
#define SSM_OBS_UNGMNA    5  /* UNGMNAMeasurement.meas_fcn (non-additive noise) ssmod.py:1085-1086              */
#define SSM_OBS_BEARING   6  /* BearingMeasurement.meas_fcn, 4 sensors ssmod.py:1189-1195 par[2i], par[2i+1] = sensor i */
#define SSM_OBS_RANGE     4  /* RangeMeasurement.meas_fcn           ssmod.py:1146-1148 par[0..1] = (sx, sy)       */

/* ---- moment-transform kinds ---------------------------------------------------------------- */
#define SSM_TF_SP 1  /* sigma-point rule, centred form, diagonal Wc   SigmaPointTransform.apply mtran.py:105-149   */
#define SSM_TF_BQ 2  /* GPQ / BSQ, un-centred form, dense Wc + model variance  BQTransform.apply bqmtran.py:60-223 */
#define SSM_TF_TP 3  /* TPQ: BQ + data-dependent variance  StudentTProcessTransform._covariance bqmtran.py:394-415 */

/* ---- filter families ----------------------------------------------------------------------- */
#define SSM_FAMILY_GAUSS   1  /* GaussianInference   ssinf.py:215-344 */
#define SSM_FAMILY_STUDENT 2  /* StudentianInference ssinf.py:555-740 */

/* One moment transform: unit points and quadrature weights (host arrays, row-major).
 * Replaces the attributes the reference's transform objects carry (mtran.py:226-232,
 * bqmtran.py:55-58, 306-310): unit_sp / model.points, wm, Wc, Wcc, model.model_var, model.iK. */
typedef struct ssm_transform {
    int32_t kind;            /* SSM_TF_*                                                        */
    int32_t dim_in;          /* D                                                               */
    int32_t dim_out;         /* E                                                               */
    int32_t n_pts;           /* N                                                               */
    const double *points;    /* (D, N) unit sigma points                                        */
    const double *wm;        /* (N)                                                             */
    const double *Wc;        /* (N, N); SSM_TF_SP reads the diagonal only                       */
    const double *Wcc;       /* (D, N); BQ / TP only                                            */
    const double *model_var; /* BQ: (E, E) matrix ADDED to the covariance (already multiplied by
                                I_out, bqmtran.py:198); TP: 1 value, the GP model variance      */
    const double *iK;        /* (N, N) inverse kernel matrix; TP only (bqmod.py:1155-1158)      */
    double nu;               /* TP degrees of freedom (always the model default 4.0) */
    int32_t tp_full_matrix;  /* TP: 1 = add the full E x E matrix (I_out is 1x1, ssinf.py:550)  */
    int32_t reserved;
} ssm_transform;

/* A filter lowered to plain data.  Replaces the object graph filter -> transform -> model ->
 * kernel of ssinf.py:233-247 (GaussianInference.__init__) / :589-624 (StudentianInference). */
typedef struct ssm_desc {
    int32_t dyn_model, obs_model;   /* SSM_DYN_*, SSM_OBS_*                                      */
    int32_t dx, dy;                 /* state / measurement dimension                             */
    double dyn_par[8], obs_par[8];  /* model parameters, see the model ids                       */
    int32_t n_state_index;          /* 0 = measurement uses the leading state components         */
    int32_t state_index[8];         /* MeasurementModel.state_index, ssmod.py:905                */
    int32_t family;                 /* SSM_FAMILY_*                                              */
    int32_t reserved;
    const double *m0;               /* (dx)      initial mean          ssinf.py:239             */
    const double *P0;               /* (dx, dx)  initial covariance (Student: scale matrix)     */
    const double *GQG;              /* (dx, dx)  G Q G^T               ssinf.py:279             */
    const double *R;                /* (dy, dy)  measurement noise cov ssinf.py:291             */
    /* Student family only (ssinf.py:589-624) */
    double dof, x0_dof, q_dof, r_dof;
    int32_t fixed_dof;
    int32_t reserved2;
    ssm_transform tf_dyn, tf_obs;
    /* Non-additive noise (TransitionModel / MeasurementModel.noise_additive == False): the transform of that model
     * works on the augmented vector [x; noise] (dim_in = dx + dq / dx + dy) with mean [m; q_mean] and covariance
     * blockdiag(P, q_cov), ssinf.py:271-272, 282-283; GQG / R are not added for such a model (:278, :290).
     * Ignored (may be NULL) for additive models. */
    const double *q_mean;           /* (dq)                                                      */
    const double *q_cov;            /* (dq, dq)                                                  */
    const double *r_mean;           /* (dy)      measurement noise mean; its covariance is R     */
    int32_t dq;                     /* process noise dimension                                   */
    int32_t reserved3;
} ssm_desc;

/* ---- library info -------------------------------------------------------------------------- */
int ssm_abi_version(void);
const char *ssm_last_error(void);

/* 1 when the forward pass (ssm_filter*) will evaluate this BQ transform with the compact sums of the
 * reflection-symmetric form, 0 when it takes the dense sums of bqmtran.py:175-223 as they stand.  The compact form
 * applies to point sets [0 | cI | -cI] whose weights are invariant, bit for bit, under every coordinate reflection
 * x_j -> -x_j (wm(j+) == wm(j-), Wc equal within each class of reflected index pairs, Wcc(d, .) zero except for
 * Wcc(d, d+) == -Wcc(d, d-)): what the formulas of bq/bqmod.py:495-523, 893-992 give in exact arithmetic.  Host-side
 * check, no CUDA call; a TPQ transform is judged by the folded BQ weights it runs with.  Environment SSM_REFL=0
 * switches the compact form off. */
int ssm_weights_reflective(const ssm_transform *tf);

/* ---- K2: fused forward pass -----------------------------------------------------------------
 * Replaces StateSpaceInference.forward_pass (ssinf.py:66-118) with _time_update (:254-295 /
 * :634-698), _measurement_update (:297-323 / :700-736), MomentTransform.apply and the model
 * functions, for n_traj independent trajectories at once.
 *   y          (dy, n_steps, ld)            measurements
 *   fi_mean    (dx, n_steps, ld)            filtered means, steps 1..N        (nullable)
 *   fi_cov     (dx*dx, n_steps, ld)         filtered covariances               (nullable)
 *   pr_mean    (dx, n_steps, ld)            predictive means                   (nullable)
 *   pr_cov     (dx*dx, n_steps, ld)         predictive covariances             (nullable)
 *   pr_xx_cov  (dx*dx, n_steps, ld)         Cov(x_k, x_{k-1}) (E x D)          (nullable)
 *   init_mean  (dx, ld), init_cov (dx*dx, ld)  per-trajectory initial moments; NULL = (m0, P0)
 *              for every trajectory (the reference after reset(), ssinf.py:249-252)
 *   last_mean / last_cov (dx, ld)/(dx*dx, ld)  moments after the last step (nullable); Student
 *              family: last_cov receives the filtered SCALE matrix x_smat_fi (ssinf.py:733)
 *   t_offset   (ld) int32 per-trajectory shift of the time index, nullable
 *   k0         time index of the first step (the reference passes time = k - 1, ssinf.py:104)
 *   status     (ld) int32, see above
 */
int ssm_filter(const ssm_desc *desc, const double *y,
               double *fi_mean, double *fi_cov,
               double *pr_mean, double *pr_cov, double *pr_xx_cov,
               const double *init_mean, const double *init_cov,
               double *last_mean, double *last_cov,
               const int32_t *t_offset, int32_t k0,
               int32_t *status, int64_t n_traj, int32_t n_steps, int64_t ld, void *stream);

/* Time-window form: the arrays have n_steps slots, only the steps [k_lo, k_hi) are processed (time index k0 + k).
 * Successive windows of the same arrays carry the filter state through last_mean / last_cov -> init_mean /
 * init_cov; for k_lo > 0 status[] must hold the status left by the preceding window (failed trajectories stay
 * failed, keep their status and are NaN-filled).  Results are bitwise identical to one ssm_filter call.  This is
 * what lets the host-streaming driver (ssmtoybox_b200/mc.py) start filtering as soon as the first time slice of
 * the measurements has reached the device, with every launch spanning all trajectories (the loops it replaces:
 * research/gpq/icinco_demo.py:115-125, research/gpq/gpq_tracking.py:52-57). */
int ssm_filter_window(const ssm_desc *desc, const double *y,
                      double *fi_mean, double *fi_cov,
                      double *pr_mean, double *pr_cov, double *pr_xx_cov,
                      const double *init_mean, const double *init_cov,
                      double *last_mean, double *last_cov,
                      const int32_t *t_offset, int32_t k0,
                      int32_t *status, int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld,
                      void *stream);

/* ssm_filter_window for consumers that read only the lower triangles of the symmetric outputs (ssm_smooth_scores does:
 * ssinf.py:342-344 needs P_k and P^-_{k+1} as symmetric matrices): the entries (row, column > row) of fi_cov and pr_cov
 * NEED NOT be written, and the compact-sum instantiation (ssm_weights_reflective) does not write them -- 20 of the 85
 * stores of a 5-D step, 536 instead of 696 bytes per trajectory-step; the other instantiations write full matrices as
 * ssm_filter_window does.  Everything else (layout, the entries that are written, status) is bit for bit what
 * ssm_filter_window produces. */
int ssm_filter_window_lower(const ssm_desc *desc, const double *y,
                            double *fi_mean, double *fi_cov,
                            double *pr_mean, double *pr_cov, double *pr_xx_cov,
                            const double *init_mean, const double *init_cov,
                            double *last_mean, double *last_cov,
                            const int32_t *t_offset, int32_t k0,
                            int32_t *status, int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld,
                            void *stream);

/* Scoring forward pass: ssm_filter_window with the error statistics of the FILTERED moments accumulated in-kernel, so a
 * filter-only Monte-Carlo run keeps no per-trajectory moment arrays (the loops of research/gpq/icinco_demo.py:115-125 and
 * research/bsq/bsq_tracking.py:300-337 keep none either): fi_mean / fi_cov may be NULL.  x_truth (dx, n_steps, ld);
 * stats rows [k_lo, k_hi) of (n_steps, ssm_scores_width(dx)) and rmse_acc (dx, ld) exactly as ssm_scores_phase1 on the
 * stored moments would fill them (bitwise); quad (n_steps, ld) = d' P^-1 d and dres (dx, n_steps, ld) = d = x - m per
 * scored unit for ssm_scores_phase2_res (both nullable).  No predictive moments (no smoother behind it).
 * Compiled for additive models, [0 | cI | -cI] point sets (UT, fully-symmetric degree 3) and the Gaussian family;
 * SSM_E_UNSUPPORTED otherwise (score the stored moments with ssm_scores_phase1 instead). */
int ssm_filter_scores(const ssm_desc *desc, const double *y, const double *x_truth, double *fi_mean, double *fi_cov,
                      double *stats, double *rmse_acc, double *quad, double *dres,
                      const double *init_mean, const double *init_cov, double *last_mean, double *last_cov,
                      const int32_t *t_offset, int32_t k0, int32_t *status,
                      int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* ---- K3: RTS smoother -----------------------------------------------------------------------
 * Replaces StateSpaceInference.backward_pass + GaussianInference._smoothing_update
 * (ssinf.py:120-147, 325-344), including the reference's index range (slots N and N-1 are never
 * smoothed).  Arrays as produced by ssm_filter (n_steps slots = steps 1..N).
 * In-kernel score accumulation (optional): with x_truth (dx, n_steps, ld) the kernel also accumulates the
 * phase-1 error statistics of the SMOOTHED moments while they are in registers -- stats (n_steps, W) and
 * rmse_acc (dx, ld, nullable) exactly as ssm_scores_phase1 would return them for (x_truth, sm_mean, sm_cov) --
 * which saves a full read pass.  x_truth = NULL: plain smoother.
 */
int ssm_smooth(int32_t dx, const double *fi_mean, const double *fi_cov,
               const double *pr_mean, const double *pr_cov, const double *pr_xx_cov,
               double *sm_mean, double *sm_cov, int32_t *status,
               const double *x_truth, double *stats, double *rmse_acc,
               int64_t n_traj, int32_t n_steps, int64_t ld, void *stream);

/* Time-window form (windows must be walked from the last to the first on one stream): the window [k_lo, k_hi) with
 * k_hi < n_steps continues the recursion from the smoothed moments the later window left in sm_mean / sm_cov and,
 * with x_truth, continues the per-trajectory sums in rmse_acc; it fills stats rows [k_lo, k_hi).  Bitwise identical
 * to one ssm_smooth call. */
int ssm_smooth_window(int32_t dx, const double *fi_mean, const double *fi_cov,
                      const double *pr_mean, const double *pr_cov, const double *pr_xx_cov,
                      double *sm_mean, double *sm_cov, int32_t *status,
                      const double *x_truth, double *stats, double *rmse_acc,
                      int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* Same pass with one more output for the two-phase scores: quad (n_steps, ld) receives d' P_s^-1 d of every scored
 * (step, trajectory) -- the first quadratic form of the log credibility ratio (utils.py:113-120), which the in-kernel
 * scoring computes anyway for the NLL -- so that ssm_scores_phase2_quad does not have to read the smoothed covariances
 * again (NaN where the covariance is not positive definite).  quad needs x_truth; quad = NULL: ssm_smooth_window. */
int ssm_smooth_quad(int32_t dx, const double *fi_mean, const double *fi_cov,
                    const double *pr_mean, const double *pr_cov, const double *pr_xx_cov,
                    double *sm_mean, double *sm_cov, int32_t *status,
                    const double *x_truth, double *stats, double *rmse_acc, double *quad,
                    int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* Score-only smoother: the same recursion and in-kernel phase-1 statistics, but the smoothed moments are NOT stored
 * (nothing reads them when only RMSE / NCI / NLL are wanted: the Monte-Carlo loops of research/gpq/icinco_demo.py:115-125
 * keep no per-trajectory arrays either).  Outputs per scored unit: quad (n_steps, ld) = d' P_s^-1 d and
 * dres (dx, n_steps, ld) = d = x - m_s, the two inputs of ssm_scores_phase2_res (both nullable).  Time windows walk
 * from the last to the first and hand the recursion over through carry (dx + dx (dx + 1) / 2, ld) (required when the
 * window is not the whole range).  status as in ssm_smooth.  Results bitwise equal to ssm_smooth_quad's statistics. */
int ssm_smooth_scores(int32_t dx, const double *fi_mean, const double *fi_cov,
                      const double *pr_mean, const double *pr_cov, const double *pr_xx_cov, int32_t *status,
                      const double *x_truth, double *stats, double *rmse_acc, double *quad, double *dres, double *carry,
                      int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* ---- K1: batched simulation -----------------------------------------------------------------
 * Replaces TransitionModel.simulate_discrete / simulate_continuous (ssmod.py:168-244),
 * MeasurementModel.simulate_measurements (ssmod.py:1011-1039) and the samplers
 * GaussRV.sample / StudentRV.sample (utils.py:618-619, 670-671).
 * Noise source: either injected arrays (x0, q, r device pointers; parity mode) or Philox4x32-10
 * keyed by (seed, global trajectory index) with Box-Muller normals coloured by the host-supplied
 * factors (rng->*_factor, A A^T = cov).
 */
typedef struct ssm_rng {
    uint64_t seed;
    int64_t traj_offset;        /* global index of trajectory 0 (sharding-invariant draws)       */
    const double *x0_mean;      /* (dx)                                                          */
    const double *x0_factor;    /* (dx, dx)  x0 = mean + F z                                     */
    const double *q_factor;     /* (dq, dq)                                                      */
    const double *r_factor;     /* (dy, dy)                                                      */
    double x0_dof, q_dof, r_dof; /* > 0: Student-t draws  n / sqrt(gamma(nu/2, 2/nu))  (utils.py:349-382); 0: Gaussian */
    int32_t dq;
    int32_t reserved;
} ssm_rng;

#define SSM_SIM_DISCRETE   1
#define SSM_SIM_CONTINUOUS 2  /* Euler-Maruyama with dt_cont; output drops x0 (ssmod.py:244) and keeps
                                 every sub-th state: output k = internal state k*sub + 1; injected q has
                                 (n_steps-1)*sub + 1 time slices, injected r n_steps                     */
#define SSM_SIM_MEASURE    3  /* measurements of a GIVEN state array x (simulate_measurements)   */

int ssm_simulate(const ssm_desc *desc, const ssm_rng *rng, int32_t mode, double dt_cont, int32_t sub,
                 const double *x0_inj, const double *q_inj, const double *r_inj,
                 double *x, double *y,
                 int64_t n_traj, int32_t n_steps, int64_t ld, void *stream);

/* ---- K5: batched Bayesian-quadrature weights ------------------------------------------------
 * Replaces GaussianProcessModel.bq_weights (bq/bqmod.py:495-523) with RBFGauss.eval /
 * eval_inv_dot / exp_x_kx / exp_x_xkx / exp_x_kxkx / exp_xy_kxy (bq/bqkern.py:96-120, 329-424)
 * and BayesSardModel.bq_weights (bq/bqmod.py:893-992, incl. utils.vandermonde :478-502 and the
 * polynomial expectations :635-797) for n_par kernel-parameter vectors at once (one CTA each).
 *   par      (n_par, D+1) host   [alpha, l_1..l_D]
 *   points   (D, N) host         unit sigma points (shared by the whole batch)
 *   mulind   (D, Q) host int32   Bayes-Sard multi-indices, NULL = plain GP weights
 * Outputs are DEVICE arrays: wm (n_par, N), Wc (n_par, N, N), Wcc (n_par, D, N),
 * iK (n_par, N, N), scal (n_par, 2) = [model_var, integral_var]; info (n_par) int32 != 0 when a
 * Cholesky factorisation failed.
 * precision: 0 = float64 arithmetic (the reference's own: error ~ eps cond(K) on wm / Wcc, eps cond(K)^2 on Wc,
 * i.e. rounding noise for the cond ~ 1e9 kernels of the reference's tracking scripts); 1 = double-double
 * arithmetic rounded once at the end (the correctly rounded value of the same formulas).
 */
int ssm_bq_weights(int32_t dim, int32_t n_pts, int32_t n_par, const double *par, const double *points,
                   const int32_t *mulind, int32_t n_basis,
                   double *wm, double *Wc, double *Wcc, double *iK, double *scal, int32_t *info,
                   int32_t precision, void *stream);

/* ---- K6: error statistics -------------------------------------------------------------------
 * Replaces utils.squared_error / mse_matrix / neg_log_likelihood / log_cred_ratio
 * (utils.py:18-148) and the reductions of research/gpq/icinco_demo.py:17-52.
 * Phase 1 accumulates, per time step k, over the trajectories with status == 0:
 *   stats[k, :] = [ sum SE (dx) | sum dx dx^T (dx*dx) | sum NLL | sum sqrt(sum_d SE) | count ]
 * (row length ssm_scores_width(dx)); rmse_acc (dx, ld) receives per-trajectory time-sums of SE.
 * The packed stats rows are what the multi-GPU path all-reduces (NCCL, one call).
 * Phase 2 takes the global per-step MSE matrices (dx*dx, n_steps) and accumulates the log
 * credibility ratio g = 10 (log10 d'P^-1 d - log10 d'MSE^-1 d):  lcr (n_steps, 2) =
 * [ sum_traj g | sum_traj |g| ]  (inclination indicator / non-credibility index numerators).
 * Both phases reduce in a fixed order (bitwise reproducible for a given trajectory count).
 */
int32_t ssm_scores_width(int32_t dx);
int ssm_scores_phase1(int32_t dx, const double *x, const double *mean, const double *cov,
                      const int32_t *status, double *stats, double *rmse_acc,
                      int64_t n_traj, int32_t n_steps, int64_t ld, void *stream);
int ssm_scores_phase2(int32_t dx, const double *x, const double *mean, const double *cov,
                      const int32_t *status, const double *mse, double *lcr,
                      int64_t n_traj, int32_t n_steps, int64_t ld, void *stream);
/* Time-window forms: rows [k_lo, k_hi) of stats / lcr (and of mse) only; phase 1 walks the windows first to last
 * (rmse_acc continues the per-trajectory sums when k_lo > 0).  The per-step MSE matrix of step k depends on step k
 * alone, so phase 2 of a window can follow its phase 1 (and the all-reduce of those rows) immediately. */
int ssm_scores_phase1_window(int32_t dx, const double *x, const double *mean, const double *cov,
                             const int32_t *status, double *stats, double *rmse_acc,
                             int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);
int ssm_scores_phase2_window(int32_t dx, const double *x, const double *mean, const double *cov,
                             const int32_t *status, const double *mse, double *lcr,
                             int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* Same two passes with per-trajectory time-sums next to the per-step sums: nll_acc (ld, nullable) = sum over the
 * window of the negative log-likelihood of each trajectory, lcr_acc (ld, nullable) = sum of its log credibility
 * ratios (continued when k_lo > 0; the caller zeroes them before a window that does not start at 0 -- e.g. [1, N)
 * for the research tables, which skip k = 0).  These are the per-simulation nllData / nciData of
 * evaluate_performance (research/gpq/icinco_demo.py:28-48, research/bsq/bsq_ungm.py:38-58) that feed bootstrap_var. */
int ssm_scores_phase1_traj(int32_t dx, const double *x, const double *mean, const double *cov,
                           const int32_t *status, double *stats, double *rmse_acc, double *nll_acc,
                           int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);
/* ... and with quad (n_steps, ld, nullable) = d' P^-1 d of every scored unit kept for ssm_scores_phase2_quad. */
int ssm_scores_phase1_quad(int32_t dx, const double *x, const double *mean, const double *cov,
                           const int32_t *status, double *stats, double *rmse_acc, double *nll_acc, double *quad,
                           int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);
int ssm_scores_phase2_traj(int32_t dx, const double *x, const double *mean, const double *cov,
                           const int32_t *status, const double *mse, double *lcr, double *lcr_acc,
                           int64_t n_traj, int32_t n_steps, int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);
/* Phase 2 from the quadratic forms stored by ssm_smooth_quad instead of the covariances: reads x, mean (2 dx doubles)
 * and quad (1 double) per unit instead of 2 dx + dx (dx + 1) / 2; bitwise the same result as ssm_scores_phase2_traj on
 * the smoothed moments. */
int ssm_scores_phase2_quad(int32_t dx, const double *x, const double *mean, const double *quad, const int32_t *status,
                           const double *mse, double *lcr, double *lcr_acc, int64_t n_traj, int32_t n_steps,
                           int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* Phase 2 from the stored errors d = x - m (dres, ssm_smooth_scores / ssm_filter_scores) and quadratic forms: dx + 1
 * doubles per unit; bitwise the same result as ssm_scores_phase2_quad. */
int ssm_scores_phase2_res(int32_t dx, const double *dres, const double *quad, const int32_t *status,
                          const double *mse, double *lcr, double *lcr_acc, int64_t n_traj, int32_t n_steps,
                          int32_t k_lo, int32_t k_hi, int64_t ld, void *stream);

/* ---- bootstrap variance of a sample mean -------------------------------------------------------
 * Replaces utils.bootstrap_var (utils.py:223-244): var[0] = population variance of the means of n_boot resamples
 * (with replacement, size n) of data (n, device).  means (n_boot, device) is caller-provided scratch that returns
 * the resample means.  Philox-keyed by (seed, resample index): deterministic for a given seed. */
int ssm_bootstrap_var(const double *data, int64_t n, int32_t n_boot, uint64_t seed, double *means, double *var,
                      void *stream);

/* ---- stand-alone moment transform / model evaluation -----------------------------------------
 * ssm_transform_apply replaces MomentTransform.apply(f, mean, cov, fcn_pars) as a public call
 * (mtran.py:105-149, bq/bqmtran.py:60-109) for n (mean, cov) pairs: mean (D, ld), cov (D*D, ld) ->
 * mean_f (E, ld), cov_f (E*E, ld), cov_fx (E*D, ld).  The integrand is a device model function:
 * which = 0 -> TransitionModel.dyn_eval of model SSM_DYN_*, which = 1 -> MeasurementModel.meas_eval of
 * model SSM_OBS_* with state_index (si0, si1) (ssmod.py:129-166, 960-1009).  par = model parameters
 * (host, 8 doubles), time = the integrand's time argument.
 * ssm_model_eval evaluates dyn_fcn / meas_fcn themselves at n points x (D, ld) with optional noise
 * (DQ, ld) (ssmod.py:268-269, 357-358, 530-564, 675-690, 1060-1061, 1114-1115, 1227-1252). */
int ssm_transform_apply(int32_t which, int32_t model, int32_t dim_state, int32_t si0, int32_t si1,
                        const double *par, const ssm_transform *tf, double time,
                        const double *mean, const double *cov, double *mean_f, double *cov_f, double *cov_fx,
                        int32_t *status, int64_t n, int64_t ld, void *stream);

/* The same transform (GPQ / BSQ kind: un-centred form, dense Wc, + model_var I) with one weight set PER COLUMN, all on
 * the device in the layout ssm_bq_weights writes: wm (n, N), Wc (n, N, N), Wcc (n, D, N), model_var (n) (nullable = 0).
 * points (D, N) on the host.  Replaces BQTransform.apply(f, mean, cov, fcn_par, kern_par) (bq/bqmtran.py:60-109 with
 * kern_par given: weights recomputed per call) for the batches of (trajectory, parameter vector) pairs that
 * MarginalInference evaluates (ssinf.py:1107-1185).  Additive models (the dispatch of ssm_model_eval). */
int ssm_transform_apply_batched(int32_t which, int32_t model, int32_t dim_state, int32_t si0, int32_t si1,
                                const double *par, int32_t n_pts, const double *points, const double *wm,
                                const double *Wc, const double *Wcc, const double *model_var, double time,
                                const double *mean, const double *cov, double *mean_f, double *cov_f, double *cov_fx,
                                int32_t *status, int64_t n, int64_t ld, void *stream);
int ssm_model_eval(int32_t which, int32_t model, int32_t dim_state, int32_t si0, int32_t si1,
                   const double *par, double time, const double *x, const double *noise, double *out,
                   int64_t n, int64_t ld, void *stream);

/* ---- RBF kernel and its Gaussian expectations (set-up path, one CTA) -------------------------
 * Replaces RBFGauss.eval (bq/bqkern.py:329-343; x1 (D, n1), x2 (D, n2) host, K (n1, n2) device) and
 * exp_x_kx / exp_x_xkx / exp_x_kxkx / exp_xy_kxy (bq/bqkern.py:345-424): q (N), R (D, N), Q (N, N),
 * kbar (1) device.  par = [alpha, l_1..l_D] host. */
int ssm_rbf_eval(int32_t dim, int32_t n1, int32_t n2, const double *par, const double *x1, const double *x2,
                 int32_t scaling, double *K, void *stream);
int ssm_rbf_expectations(int32_t dim, int32_t n_pts, const double *par, const double *points, int32_t scaling,
                         double *q, double *R, double *Q, double *kbar, void *stream);

/* Monte-Carlo counterpart for a standard Student-t density (RBFStudent, bq/bqkern.py:457-536; 2 * 10^6 samples in
 * the reference): q (N), R (D, N), Q (N, N), kbar (1) of the UNSCALED kernel from n_samples draws of t_dof(0, I),
 * Philox-keyed by (seed, sample index).  dim <= 8, n_pts <= 32. */
int ssm_rbf_student_expectations(int32_t dim, int32_t n_pts, const double *par, const double *points, double dof,
                                 int64_t n_samples, uint64_t seed, double *q, double *R, double *Q, double *kbar,
                                 void *stream);

/* ---- K5c: marginal likelihood of the integrand model (hyper-parameter fitting) -------------------
 * Replaces GaussianProcessModel.neg_log_marginal_likelihood (bq/bqmod.py:537-596; nu = 0) and
 * StudentTProcessModel.neg_log_marginal_likelihood (bq/bqmod.py:1191-1245; nu > 2) with RBFGauss.der_par
 * (bq/bqkern.py:426-436), the objective of Model.optimize (bq/bqmod.py:250-285), for n_par kernel LOG-parameter
 * vectors at once (optimiser restarts / parameter grids; one CTA each).
 *   log_par (n_par, D+1) host  log [alpha, l_1..l_D];  x_obs (D, N) host;  fcn_obs (N, E) host;
 *   jitter (N, N) host, added to the kernel matrix (the reference passes 1e-8 I), nullable.
 * Outputs (device): nlml (n_par), grad (n_par, D+1) -- as the reference defines it (der_par: alpha-derivative for
 * column 0, log-lengthscale derivatives for the others), info (n_par) int32 = 1 where the kernel matrix is not
 * positive definite (nlml / grad = NaN; numpy.linalg.LinAlgError in the reference).  dim <= 8, N <= 32, E <= 8. */
int ssm_gp_nlml(int32_t dim, int32_t n_pts, int32_t n_out, int32_t n_par, const double *log_par,
                const double *x_obs, const double *fcn_obs, const double *jitter, double nu, double *nlml,
                double *grad, int32_t *info, void *stream);

/* ---- stand-alone sampler ---------------------------------------------------------------------
 * Replaces GaussRV.sample / StudentRV.sample (utils.py:618-619, 670-671): out (dim, ld) =
 * mean + F z (dof = 0) or mean + F z / sqrt(gamma(dof/2, 2/dof)) (dof > 0), Philox-keyed by
 * (seed, offset + sample index). */
int ssm_sample(int32_t dim, const double *mean, const double *factor, double dof, uint64_t seed, int64_t offset,
               double *out, int64_t n, int64_t ld, void *stream);

/* Gaussian-mixture counterpart: utils.gauss_mixture (utils.py:261-301) and GaussianMixtureRV.sample
 * (research/tpq/tpq_base.py:13-31), the heavy-tailed data generators of the TPQ experiments.  means (K, dim),
 * factors (K, dim, dim) with F_k F_k^T = cov_k, alphas (K) host; out (dim, ld) device, idx (ld) int32 device
 * (component of each sample, nullable).  dim <= 8, K <= 4. */
int ssm_sample_mixture(int32_t dim, int32_t n_comp, const double *means, const double *factors, const double *alphas,
                       uint64_t seed, int64_t offset, double *out, int32_t *idx, int64_t n, int64_t ld, void *stream);

/* ---- strided copy of a trajectory range ---------------------------------------------------------
 * height rows of width bytes with row pitches dpitch / spitch (bytes): moves columns [a, b) of a host array laid
 * out [component][step][trajectory] into a compact device chunk (or back).  Asynchronous on the stream when the
 * host side is pinned.  Used by the host-streaming Monte-Carlo driver (research loops over y[..., i]). */
int ssm_memcpy2d(void *dst, uint64_t dpitch, const void *src, uint64_t spitch, uint64_t width, uint64_t height,
                 int32_t host_to_device, void *stream);

/* ---- FP64 FMA micro-benchmark (roofline denominator; MEASURED_PEAKS.json has no fp64 figure) --
 * Launches a dependent-chain-free DFMA loop; returns the number of FLOPs it executes in *flops.
 * The caller times it with CUDA events. */
int ssm_fp64_peak_kernel(int32_t n_blocks, int32_t n_iters, double *sink, double *flops, void *stream);

/* ---- device math probe -----------------------------------------------------------------------------
 * Evaluates the library's own fp64 routines at n points (device arrays): which = 0 exp(a), 1 sqrt(a), 2 1/sqrt(a),
 * 3 a / b, 4 atan2(a, b).  b is read for which >= 3 only.  Test hook: the forward pass calls these routines out of
 * line (one shared copy each) -- the parity suite measures their error in ulp against
 * numpy over the whole argument range. */
int ssm_math_probe(int32_t which, const double *a, const double *b, double *out, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SSM_B200_H */
