"""research/gpq/icinco_demo.py as batched GPU workloads: `tables` (:81-168) and `hypers_demo` (:171-224)."""
import numpy as np
import pandas as pd

from ..ssinf import CubatureKalman, UnscentedKalman, GaussHermiteKalman, GaussianProcessKalman
from ..ssmod import UNGMTransition, UNGMMeasurement
from ..utils import GaussRV
from . import scoring
from .scoring import evaluate_performance  # noqa: F401  (same name and signature as icinco_demo.py:17)


def _ungm():
    dyn = UNGMTransition(GaussRV(1, cov=np.atleast_2d(5.0)), GaussRV(1, cov=np.atleast_2d(10.0)))
    obs = UNGMMeasurement(GaussRV(1), 1)
    return dyn, obs


def algorithms(dyn, obs):
    """The 14 filters / smoothers of icinco_demo.py:92-111 (seven classical rules, seven GPQ counterparts)."""
    kern_par_sr = np.array([[1.0, 0.3 * dyn.dim_in]])
    kern_par_ut = np.array([[1.0, 3.0 * dyn.dim_in]])
    kern_par_gh = np.array([[1.0, 0.1 * dyn.dim_in]])
    return (
        CubatureKalman(dyn, obs),
        UnscentedKalman(dyn, obs),
        GaussHermiteKalman(dyn, obs),
        GaussHermiteKalman(dyn, obs),
        GaussHermiteKalman(dyn, obs),
        GaussHermiteKalman(dyn, obs),
        GaussHermiteKalman(dyn, obs),
        GaussianProcessKalman(dyn, obs, kern_par_sr, kern_par_sr, points='sr'),
        GaussianProcessKalman(dyn, obs, kern_par_ut, kern_par_ut, points='ut'),
        GaussianProcessKalman(dyn, obs, kern_par_sr, kern_par_sr, points='gh', point_hyp={'degree': 5}),
        GaussianProcessKalman(dyn, obs, kern_par_gh, kern_par_gh, points='gh', point_hyp={'degree': 7}),
        GaussianProcessKalman(dyn, obs, kern_par_gh, kern_par_gh, points='gh', point_hyp={'degree': 10}),
        GaussianProcessKalman(dyn, obs, kern_par_gh, kern_par_gh, points='gh', point_hyp={'degree': 15}),
        GaussianProcessKalman(dyn, obs, kern_par_gh, kern_par_gh, points='gh', point_hyp={'degree': 20}),
    )


def tables(steps=500, sims=100, x=None, z=None, bootstrap_variance=True, num_bs_samples=10000):
    """icinco_demo.py:81-168.  x (1, steps, sims), z (1, steps, sims): optional data (the reference draws them from
    numpy's global MT19937 stream, which the device cannot replay: by default they are simulated with Philox).
    Returns the same dict of six pandas tables (rows SR, UT, GH-5 ... GH-20; columns Classical, Bayesian and their
    `2 std`)."""
    dyn, obs = _ungm()
    if x is None:
        x = dyn.simulate_discrete(steps, mc_sims=sims, device_out=True)
    if z is None:
        z = obs.simulate_measurements(x, device_out=True)
    res = scoring.run_all(algorithms(dyn, obs), z)
    sc = evaluate_performance(x, [r['mean_f'] for r in res], [r['cov_f'] for r in res], [r['mean_s'] for r in res],
                              [r['cov_s'] for r in res], bootstrap_variance, num_bs_samples, status=[r['status'] for r in res])
    mean = sc[:6]
    std = sc[6:] if bootstrap_variance else [np.zeros((len(res), 1))] * 6
    row_labels = ['SR', 'UT', 'GH-5', 'GH-7', 'GH-10', 'GH-15', 'GH-20']
    col_labels = ['Classical', 'Bayesian', 'Classical (2std)', 'Bayesian (2std)']
    keys = ('filter_RMSE', 'filter_NCI', 'filter_NLL', 'smoother_RMSE', 'smoother_NCI', 'smoother_NLL')
    return {k: pd.DataFrame(np.hstack((m.reshape(2, 7).T, s.reshape(2, 7).T)), index=row_labels, columns=col_labels)
            for k, m, s in zip(keys, mean, std)}


def hypers_demo(lscale=None, steps=500, mc=100, x=None, z=None, carry_over=False):
    """icinco_demo.py:171-224 without the plot: RMSE / NCI / NLL of the GPQ Kalman filter (UT points, kappa = 0)
    against the kernel length-scale.
    carry_over=False (default): every trajectory starts from the model's initial moments and all of them run in one
    launch per length-scale.  carry_over=True reproduces the reference literally: it never calls reset(), so every
    trajectory starts from the previous trajectory's last posterior (icinco_demo.py:195-196) -- a
    serial chain over the simulations, run here as `mc` single-trajectory calls."""
    if lscale is None:
        lscale = [1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1, 3, 1e1, 3e1, 1e2]
    dyn, obs = _ungm()
    if x is None:
        x = dyn.simulate_discrete(steps, mc_sims=mc, device_out=True)
    if z is None:
        z = obs.simulate_measurements(x, device_out=True)
    xd, zd = scoring.to_device(x), scoring.to_device(z)
    dim, N, M = xd.shape
    rmse, nci, nll = [], [], []
    for el in lscale:
        ker_par = np.array([[1.0, el * dyn.dim_in]])
        f = GaussianProcessKalman(dyn, obs, ker_par, ker_par, points='ut', point_hyp={'kappa': 0.0})
        if carry_over:
            zh = zd.cpu().numpy()
            mf, Pf = np.zeros((dim, N, M)), np.zeros((dim, dim, N, M))
            for s in range(M):
                mf[..., s], Pf[..., s] = f.forward_pass(zh[..., s])
            mf, Pf, st = scoring.to_device(mf), scoring.to_device(Pf), None
        else:
            mf, Pf = f.forward_pass(zd)
            st = f.status
        r = scoring.score_pass(xd, mf, Pf, st, skip_first=False)
        n_ok = r['ok'].sum()
        rmse.append(scoring._mean_ok(r['rmse_data'], r['ok']))
        nci.append((r['lcr'][:, 0].sum() / (N * n_ok)).reshape(1))
        nll.append((r['stats'][:, dim + dim * dim].sum() / (N * n_ok)).reshape(1))
    stack = lambda v: np.stack([t.cpu().numpy() for t in v], axis=-1)  # noqa: E731
    return {'el': lscale, 'rmse': stack(rmse), 'nci': stack(nci), 'neg_log_likelihood': stack(nll)}
