"""research/tpq/tpq_base.py `run_filters` (:175-192) and `eval_perf_scores` (:154-172) as batched GPU workloads."""
import numpy as np
import torch

from . import scoring
from ..ssinf import StudentianInference
from ..mtran import FullySymmetricStudentTransform
from ..bq.bqmtran import GaussianProcessTransform
from ..bq.bqkern import RBFStudent
from ..utils import GaussianMixtureRV  # noqa: F401  (tpq_base.py:13-31)


class GPQStudent(StudentianInference):
    """Student filter with GPQ transforms on fully-symmetric points and Student-density kernel expectations
    (tpq_base.py:41-91).  The reference's constructor predates the dim_out argument of GaussianProcessTransform
    (it passes kern_par in its place, bqmtran.py:285-310) and no longer runs; the intended transforms are built here."""

    def __init__(self, dyn, obs, kern_par_dyn, kern_par_obs, point_hyp=None, dof=4.0, fixed_dof=True):
        _, _, q_dof = dyn.noise_rv.get_stats()
        _, _, r_dof = obs.noise_rv.get_stats()
        point_hyp = dict() if point_hyp is None else point_hyp
        point_hyp_dyn, point_hyp_obs = dict(point_hyp, dof=q_dof), dict(point_hyp, dof=r_dof)
        t_dyn = GaussianProcessTransform(dyn.dim_in, dyn.dim_state, kern_par_dyn, 'rbf-student', 'fs', point_hyp_dyn)
        t_obs = GaussianProcessTransform(obs.dim_in, obs.dim_out, kern_par_obs, 'rbf-student', 'fs', point_hyp_obs)
        super(GPQStudent, self).__init__(dyn, obs, t_dyn, t_obs, dof, fixed_dof)


class FSQStudent(StudentianInference):
    """Student filter on fully-symmetric rules with the noise dofs in the point sets (tpq_base.py:94-105)."""

    def __init__(self, dyn, obs, degree=3, kappa=None, dof=4.0, fixed_dof=True):
        _, _, q_dof = dyn.noise_rv.get_stats()
        _, _, r_dof = obs.noise_rv.get_stats()
        t_dyn = FullySymmetricStudentTransform(dyn.dim_in, degree, kappa, q_dof)
        t_obs = FullySymmetricStudentTransform(obs.dim_in, degree, kappa, r_dof)
        super(FSQStudent, self).__init__(dyn, obs, t_dyn, t_obs, dof, fixed_dof)


def rbf_student_mc_weights(x, kern, num_samples, num_batch=None, seed=0):
    """tpq_base.py:108-151: Monte-Carlo BQ weights (wm, Wc, Wcc, Q) of the RBF kernel under a standard Student
    density, from ONE device launch over num_samples draws (num_batch is accepted and ignored: the batches exist in
    the reference only to bound host memory).  Wc is NOT symmetrised, as in the reference."""
    assert isinstance(kern, RBFStudent)
    k = RBFStudent(kern.dim, kern.par, dof=kern.dof, num_samples=num_samples, seed=seed)
    iK = k.eval_inv_dot(k.par, x, scaling=False)
    q, R, Q, _ = k._expect(k.par, x, False)
    return q.dot(iK), iK.dot(Q).dot(iK), R.dot(iK), Q


def run_filters(filters, z):
    """Filtered means (xD, steps, mc, n_filt) and covariances (xD, xD, steps, mc, n_filt) of every filter on the
    measurements z (zD, steps, mc): one batched launch per filter, each trajectory started from the initial moments
    (the reference calls reset() after every simulation, tpq_base.py:188-189).  numpy in -> numpy out; a CUDA tensor
    in -> lists of per-filter device tensors (no 5-D host array is assembled)."""
    import torch
    res = scoring.run_all(filters, z, smooth=False)
    if isinstance(z, torch.Tensor) and z.is_cuda:
        return [r['mean_f'] for r in res], [r['cov_f'] for r in res]
    mf = np.stack([r['mean_f'].cpu().numpy() for r in res], axis=-1)
    Pf = np.stack([r['cov_f'].cpu().numpy() for r in res], axis=-1)
    return mf, Pf


def eval_perf_scores(x, mf, Pf):
    """tpq_base.py:154-172: RMSE (norm of the state error) and inclination indicator (log credibility ratio against
    the per-step MSE matrix + 1e-6 I), both averaged over the simulations -> two (steps, n_filt) arrays."""
    xd = scoring.to_device(x)
    xD = xd.shape[0]
    mfs, Pfs = scoring._split_algs(mf, 4), scoring._split_algs(Pf, 5)
    rmse, lcr = [], []
    for m, P in zip(mfs, Pfs):
        # a trajectory whose filter failed (NaN-filled from the failing step on) is left out of the averages; the
        # reference has no such case: its run_filters would have stopped with the exception
        status = torch.isnan(m[0, -1]).to(torch.int32)
        r = scoring.score_pass(xd, m, P, status if bool(status.any()) else None, skip_first=False, reg=1e-6 * np.eye(xD))
        rmse.append((r['stats'][:, xD + xD * xD + 1] / r['count']).cpu().numpy())
        lcr.append((r['lcr'][:, 0] / r['count']).cpu().numpy())
    return np.stack(rmse, axis=1), np.stack(lcr, axis=1)
