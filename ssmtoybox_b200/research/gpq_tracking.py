"""research/gpq/gpq_tracking.py as batched GPU workloads: `reentry_gpq_demo` (:9-111, the 5-D reentry vehicle -- the
benchmark configuration of this repository) and `reentry_simple_gpq_demo` (:114-314, vertically falling body with a
range sensor).  Plots are left out; the arrays behind them are returned."""
import numpy as np
import torch

from ..ssinf import GaussianProcessKalman, UnscentedKalman
from ..ssmod import ReentryVehicle1DTransition, RangeMeasurement, ReentryVehicle2DTransition, Radar2DMeasurement
from ..utils import GaussRV
from . import scoring


def _sub(x, m, P, status, idx):
    """scores of the sub-vector idx of the state: (rmse_vs_time, inc_vs_time), both (steps,)"""
    i = torch.as_tensor(idx, device=x.device)
    r = scoring.score_pass(x[i].contiguous(), m[i].contiguous(), P[i][:, i].contiguous(), status, skip_first=False)
    d = len(idx)
    return (r['stats'][:, d + d * d + 1] / r['count']).cpu().numpy(), (r['lcr'][:, 0] / r['count']).cpu().numpy()


def reentry_gpq_demo(mc_sims=20, duration=200, x=None, y=None, alg=None):
    """gpq_tracking.py:9-111: GPQKF (RBF, UT points) against the UKF on the reentry vehicle.  Returns the position
    RMSE against time and the inclination indicator of the first four states against time, each (steps, 2)
    [GPQKF, UKF], with their time averages (the two numbers the script prints).
    alg: optional replacement for the (GPQKF, UKF) pair, e.g. with quadrature weights assigned from outside -- the
    script's kernel parameters make cond(K) ~ 1e9, so its float64 weights are rounding noise and
    differ from the double-double weights computed here."""
    disc_tau = 0.1
    m0 = np.array([6500.4, 349.14, -1.8093, -6.7967, 0.6932])
    Q = np.diag([2.4064e-5, 2.4064e-5, 0])
    sys = ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0])), GaussRV(3, cov=Q), dt=disc_tau)
    obs = Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([sys.R0, 0]))
    if x is None:
        x = sys.simulate_continuous(duration=duration, dt=disc_tau, mc_sims=mc_sims, device_out=True)
        y = obs.simulate_measurements(x, device_out=True)
    m0 = np.array([6500.4, 349.14, -1.8093, -6.7967, 0])
    dyn = ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 1])), GaussRV(3, cov=disc_tau * Q), dt=disc_tau)
    hdyn = np.array([[1.0, 25, 25, 25, 25, 25]])
    hobs = np.array([[1.0, 25, 25, 1e4, 1e4, 1e4]])
    if alg is None:
        alg = (GaussianProcessKalman(dyn, obs, hdyn, hobs, kernel='rbf', points='ut'), UnscentedKalman(dyn, obs))
    xd = scoring.to_device(x)
    res = scoring.run_all(alg, y, smooth=False)
    pos = [_sub(xd, r['mean_f'], r['cov_f'], r['status'], [0, 1])[0] for r in res]
    inc = [_sub(xd, r['mean_f'], r['cov_f'], r['status'], [0, 1, 2, 3])[1] for r in res]   # lcr of x[:4], :80-84
    pos, inc = np.stack(pos, axis=1), np.stack(inc, axis=1)
    return {'models': (dyn, obs), 'pos_rmse_vs_time': pos, 'inc_ind_vs_time': inc, 'avg_rmse': pos.mean(axis=0), 'avg_inc': inc.mean(axis=0),
            'n_failed': [int((r['status'] != 0).sum().item()) for r in res]}


def reentry_simple_gpq_demo(dur=30, tau=0.1, mc=100, x=None, y=None, alg=None):
    """gpq_tracking.py:114-314: altitude / velocity / ballistic coefficient of a falling body from range measurements.
    Returns per-state RMSE and inclination indicator against time, each (steps, 2) [GPQKF, UKF], and the average RMSE
    the script prints."""
    P0 = np.diag([0.0929, 1.4865, 1e-4])
    sys = ReentryVehicle1DTransition(GaussRV(3, np.array([90, 6, 1.5]), P0), GaussRV(3, cov=np.zeros((3, 3))), dt=tau)
    obs = RangeMeasurement(GaussRV(1, cov=np.array([[0.03048 ** 2]])), 3)
    if x is None:
        x = sys.simulate_continuous(dur, mc_sims=mc, device_out=True)
        x = x[..., (x >= 0).all(dim=0).all(dim=0)].contiguous()     # only non-divergent trajectories (:145)
        y = obs.simulate_measurements(x, device_out=True)
    dyn = ReentryVehicle1DTransition(GaussRV(3, np.array([90, 6, 1.7]), P0), GaussRV(3, cov=np.zeros((3, 3))), dt=tau)
    kpar_dyn_ut = np.array([[0.5, 10, 10, 10]])
    kpar_obs_ut = np.array([[0.5, 15, 20, 20]])
    if alg is None:
        alg = (GaussianProcessKalman(dyn, obs, kpar_dyn_ut, kpar_obs_ut, kernel='rbf', points='ut'), UnscentedKalman(dyn, obs))
    xd = scoring.to_device(x)
    res = scoring.run_all(alg, y, smooth=False)
    out = {'models': (dyn, obs)}
    for name, i in (('pos', 0), ('vel', 1), ('theta', 2)):
        sc = [_sub(xd, r['mean_f'], r['cov_f'], r['status'], [i]) for r in res]
        out[name + '_rmse_vs_time'] = np.stack([s[0] for s in sc], axis=1)
        out[name + '_inc_vs_time'] = np.stack([s[1] for s in sc], axis=1)
    full = [_sub(xd, r['mean_f'], r['cov_f'], r['status'], [0, 1, 2])[0] for r in res]
    out['avg_rmse'] = np.stack(full, axis=1).mean(axis=0)             # np.sqrt(error2.sum(axis=0)).mean(axis=(0, 1)), :313
    out['n_failed'] = [int((r['status'] != 0).sum().item()) for r in res]
    return out
