"""research/tpq/tpq_ungm.py `ungm_demo` (:39-174) as a one-call GPU workload: the UNGM experiment of the TPQ paper.

Data come from a UNGM whose process and measurement noises are two-component Gaussian mixtures (80 % nominal, 20 %
with 10x / 100x variance); the filters work on a Student-t state-space model (nu = 4) or a Gaussian one: UKF,
fully-symmetric Student filter and three Student-t process quadrature Student filters (TPQSF) with different
degrees of freedom of the t-process.  As in the reference, every BQ transform gets the SAME Monte-Carlo weights,
computed once from 10^6 Student-t samples (`rbf_student_mc_weights`, one device launch per transform instead of 1000
numpy batches); the scores are the time-resolved average RMSE and inclination indicator of `tpq_base.eval_perf_scores`
with bootstrap variances of their time means.

The reference's own driver no longer runs: `GaussianMixtureRV.sample` hands the (samples, indexes) pair of
`utils.gauss_mixture` to `np.moveaxis` (research/tpq/tpq_base.py:27-28), so there is no golden output to replay; the
test checks the driver's outputs for consistency instead (tests/test_gpu_research.py)."""
import numpy as np

from . import tpq_base
from .. import device as dv
from ..bq.bqmtran import BQTransform
from ..ssinf import UnscentedKalman, FullySymmetricStudent, StudentProcessStudent
from ..ssmod import UNGMTransition, UNGMMeasurement
from ..utils import GaussRV, StudentRV, GaussianMixtureRV


def ungm_demo(steps=250, mc_sims=100, x=None, z=None, mc_weight_samples=int(1e6), num_bs_samples=int(1e4)):
    # SYSTEM (data generator): dynamics and measurement                                   tpq_ungm.py:40-56
    x0_cov = 1.0
    q_cov_0, q_cov_1 = 10.0, 100.0
    r_cov_0, r_cov_1 = 0.01, 1.0
    zero_means = (np.zeros((1,)), np.zeros((1,)))
    gm_weights = np.array([0.8, 0.2])
    if x is None:
        q = GaussianMixtureRV(1, zero_means, (np.atleast_2d(q_cov_0), np.atleast_2d(q_cov_1)), gm_weights)
        dyn = UNGMTransition(GaussRV(1, cov=x0_cov), q)
        x = dyn.simulate_discrete(steps, mc_sims, device_out=True)
    if z is None:
        r = GaussianMixtureRV(1, zero_means, (np.atleast_2d(r_cov_0), np.atleast_2d(r_cov_1)), gm_weights)
        z = UNGMMeasurement(r, 1).simulate_measurements(x, device_out=True)

    # STUDENT STATE SPACE MODEL                                                            tpq_ungm.py:58-64
    nu = 4.0
    dyn = UNGMTransition(StudentRV(1, scale=(nu - 2) / nu * x0_cov, dof=nu), StudentRV(1, scale=((nu - 2) / nu) * q_cov_0, dof=nu))
    obs = UNGMMeasurement(StudentRV(1, scale=((nu - 2) / nu) * r_cov_0, dof=nu), dyn.dim_state)
    # GAUSSIAN SSM for the UKF                                                             tpq_ungm.py:66-71
    dyn_gauss = UNGMTransition(GaussRV(1, cov=x0_cov), GaussRV(1, cov=q_cov_0))
    obs_gauss = UNGMMeasurement(GaussRV(1, cov=r_cov_0), dyn.dim_state)

    par_dyn_tp = np.array([[3.0, 1.0]])                                                  # tpq_ungm.py:77-78
    par_obs_tp = np.array([[3.0, 3.0]])
    kappa = 0.0
    par_pt = {'kappa': kappa}
    filters = (                                                                          # tpq_ungm.py:92-107
        UnscentedKalman(dyn_gauss, obs_gauss, kappa=kappa),
        FullySymmetricStudent(dyn, obs, kappa=kappa, dof=4.0),
        StudentProcessStudent(dyn, obs, par_dyn_tp, par_obs_tp, dof=4.0, dof_tp=3.0, point_par=par_pt),
        StudentProcessStudent(dyn, obs, par_dyn_tp, par_obs_tp, dof=4.0, dof_tp=10.0, point_par=par_pt),
        StudentProcessStudent(dyn, obs, par_dyn_tp, par_obs_tp, dof=4.0, dof_tp=500.0, point_par=par_pt),
    )
    itpq = [i for i, f in enumerate(filters) if isinstance(f, StudentProcessStudent)][0]

    # one set of Monte-Carlo weights per transform, assigned to every BQ filter            tpq_ungm.py:110-126
    weights = {}
    for which in ('tf_dyn', 'tf_obs'):
        tf = getattr(filters[itpq], which)
        wm, wc, wcc, Q = tpq_base.rbf_student_mc_weights(tf.model.points, tf.model.kernel, mc_weight_samples, 1000)
        weights[which] = (wm, wc, wcc, Q)
        for f in filters:
            t = getattr(f, which)
            if isinstance(t, BQTransform):
                t.wm, t.Wc, t.Wcc = wm, wc, wcc
                t.Q = Q

    mf, Pf = tpq_base.run_filters(filters, z)                                            # tpq_ungm.py:135
    rmse_avg, lcr_avg = tpq_base.eval_perf_scores(x, mf, Pf)                             # tpq_ungm.py:138
    var_rmse_avg, var_lcr_avg = np.zeros((len(filters),)), np.zeros((len(filters),))     # tpq_ungm.py:141-146
    import torch
    for fi in range(len(filters)):
        var_rmse_avg[fi] = float(dv.bootstrap_var(torch.as_tensor(rmse_avg[:, fi], device='cuda'), num_bs_samples, seed=2 * fi))
        var_lcr_avg[fi] = float(dv.bootstrap_var(torch.as_tensor(lcr_avg[:, fi], device='cuda'), num_bs_samples, seed=2 * fi + 1))
    f_label = [f.__class__.__name__ for f in filters]
    table = np.array([rmse_avg.mean(axis=0), np.sqrt(var_rmse_avg), lcr_avg.mean(axis=0), np.sqrt(var_lcr_avg)]).T   # tpq_ungm.py:166-169
    return {'x': x, 'z': z, 'rmse_avg': rmse_avg, 'lcr_avg': lcr_avg, 'var_rmse_avg': var_rmse_avg, 'var_lcr_avg': var_lcr_avg,
            'labels': f_label, 'columns': ['MEAN_RMSE', 'STD(MEAN_RMSE)', 'MEAN_INC', 'STD(MEAN_INC)'], 'table': table,
            'weights': weights, 'n_failed': [int(torch.isnan(m[0, -1]).sum()) if isinstance(m, torch.Tensor) else int(np.isnan(m[0, -1]).sum())
                                             for m in (mf if isinstance(mf, list) else [mf[..., i] for i in range(len(filters))])]}
