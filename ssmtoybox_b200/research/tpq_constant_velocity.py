"""research/tpq/tpq_constant_velocity.py `constant_velocity_radar_demo` (:12-143) as a one-call GPU workload: tracking of an
object moving with constant velocity from radar measurements with glint noise (a two-component Gaussian mixture: 15 % of
the measurements come with 100x / 40x the nominal variance), filtered on a Student-t state-space model by the
fully-symmetric Student filter and a Student-t process quadrature Student filter (TPQSF) whose BQ transforms get
Monte-Carlo weights computed once from 2 * 10^6 Student-t samples.

The reference's own driver no longer runs -- its process-noise covariance `G.T.dot(Q).dot(G)` multiplies a (2, 4) by a
(2, 2) matrix (:24-28), it reads `f.tf_meas` where the filters have `tf_obs` (:99-105), and `GaussianMixtureRV.sample`
breaks (research/tpq/tpq_base.py:27-28) -- so this is the experiment as intended: ConstantVelocity's own 2-D acceleration
noise with covariance Q through its noise gain (ssmod.py:831-846), radar on the position components [0, 2].  There is no
golden output to replay; the test checks the driver's outputs for consistency (tests/test_gpu_research.py)."""
import numpy as np
import torch

from . import tpq_base
from .. import device as dv
from ..bq.bqmtran import BQTransform
from ..ssinf import FullySymmetricStudent, StudentProcessStudent
from ..ssmod import ConstantVelocity, Radar2DMeasurement
from ..utils import GaussRV, StudentRV, GaussianMixtureRV


def constant_velocity_radar_demo(steps=100, mc_sims=100, x=None, z=None, mc_weight_samples=int(2e6), num_bs_samples=int(1e4)):
    # SYSTEM (data generator)                                                              tpq_constant_velocity.py:16-37
    dt = 0.5
    Q = np.diag([50.0, 5.0])
    R0 = np.diag([50, 0.4e-6])
    R1 = np.diag([5000, 1.6e-5])        # glint (outlier) covariance
    glint_prob = 0.15
    if x is None:
        m0 = np.array([10000, 300, 1000, -40], dtype=float)
        P0 = np.diag([100 ** 2, 10 ** 2, 100 ** 2, 10 ** 2])
        x = ConstantVelocity(GaussRV(4, m0, P0), GaussRV(2, cov=Q), dt).simulate_discrete(steps, mc_sims, device_out=True)
    if z is None:
        r = GaussianMixtureRV(2, (np.zeros(2), np.zeros(2)), (R0, R1), np.array([1 - glint_prob, glint_prob]))
        z = Radar2DMeasurement(r, 4, state_index=[0, 2]).simulate_measurements(x, device_out=True)

    # STUDENT STATE SPACE MODEL                                                            tpq_constant_velocity.py:39-51
    m0 = np.array([10175, 295, 980, -35], dtype=float)
    P0 = np.diag([100 ** 2, 10 ** 2, 100 ** 2, 10 ** 2])
    x0_dof = 1000.0
    dyn = ConstantVelocity(StudentRV(4, m0, ((x0_dof - 2) / x0_dof) * P0, x0_dof),
                           StudentRV(2, scale=((x0_dof - 2) / x0_dof) * Q, dof=x0_dof), dt)
    r_dof = 4.0
    obs = Radar2DMeasurement(StudentRV(2, scale=((r_dof - 2) / r_dof) * R0, dof=r_dof), dyn.dim_state, state_index=[0, 2])

    par_dyn_tp = np.array([[0.05, 100, 100, 100, 100]], dtype=float)                     # tpq_constant_velocity.py:60-62
    par_obs_tp = np.array([[0.005, 10, 100, 10, 100]], dtype=float)
    kappa = 0.0
    par_pt = {'kappa': kappa}
    filters = (                                                                          # tpq_constant_velocity.py:76-85
        FullySymmetricStudent(dyn, obs, kappa=kappa, dof=4.0),
        StudentProcessStudent(dyn, obs, par_dyn_tp, par_obs_tp, dof=4.0, dof_tp=4.0, point_par=par_pt),
    )
    itpq = [i for i, f in enumerate(filters) if isinstance(f, StudentProcessStudent)][0]
    weights = {}                                                                         # tpq_constant_velocity.py:88-105
    for which in ('tf_dyn', 'tf_obs'):
        tf = getattr(filters[itpq], which)
        wm, wc, wcc, Qk = tpq_base.rbf_student_mc_weights(tf.model.points, tf.model.kernel, mc_weight_samples, 1000)
        weights[which] = (wm, wc, wcc, Qk)
        for f in filters:
            t = getattr(f, which)
            if isinstance(t, BQTransform):
                t.wm, t.Wc, t.Wcc = wm, wc, wcc
                t.Q = Qk

    mf, Pf = tpq_base.run_filters(filters, z)                                            # :108
    xd = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x), device='cuda')

    def sub(idx):                                                                        # :111-115, position / velocity blocks
        i = torch.as_tensor(idx, device='cuda')
        ms = [m[i].contiguous() for m in (mf if isinstance(mf, list) else [torch.as_tensor(np.ascontiguousarray(mf[..., a]), device='cuda') for a in range(len(filters))])]
        Ps = [P[i][:, i].contiguous() for P in (Pf if isinstance(Pf, list) else [torch.as_tensor(np.ascontiguousarray(Pf[..., a]), device='cuda') for a in range(len(filters))])]
        return tpq_base.eval_perf_scores(xd[i].contiguous(), ms, Ps)
    pos_rmse, pos_lcr = sub([0, 2])
    vel_rmse, vel_lcr = sub([1, 3])
    rmse_avg, lcr_avg = tpq_base.eval_perf_scores(x, mf, Pf)
    var_rmse_avg, var_lcr_avg = np.zeros((len(filters),)), np.zeros((len(filters),))     # :118-123
    for fi in range(len(filters)):
        var_rmse_avg[fi] = float(dv.bootstrap_var(torch.as_tensor(rmse_avg[:, fi], device='cuda'), num_bs_samples, seed=2 * fi))
        var_lcr_avg[fi] = float(dv.bootstrap_var(torch.as_tensor(lcr_avg[:, fi], device='cuda'), num_bs_samples, seed=2 * fi + 1))
    f_label = [f.__class__.__name__ for f in filters]
    table = np.array([rmse_avg.mean(axis=0), np.sqrt(var_rmse_avg), lcr_avg.mean(axis=0), np.sqrt(var_lcr_avg)]).T   # :152-156
    ms = mf if isinstance(mf, list) else [mf[..., a] for a in range(len(filters))]
    return {'x': x, 'z': z, 'rmse_avg': rmse_avg, 'lcr_avg': lcr_avg, 'var_rmse_avg': var_rmse_avg, 'var_lcr_avg': var_lcr_avg,
            'pos_rmse': pos_rmse, 'pos_lcr': pos_lcr, 'vel_rmse': vel_rmse, 'vel_lcr': vel_lcr, 'steps': steps, 'mc_sims': mc_sims,
            'par_dyn_tp': par_dyn_tp, 'par_obs_tp': par_obs_tp, 'labels': f_label,
            'columns': ['MEAN_RMSE', 'STD(MEAN_RMSE)', 'MEAN_INC', 'STD(MEAN_INC)'], 'table': table, 'weights': weights,
            'n_failed': [int(torch.isnan(m[0, -1]).sum()) if isinstance(m, torch.Tensor) else int(np.isnan(m[0, -1]).sum()) for m in ms]}
