"""research/bsq/bsq_tracking.py `reentry_demo` (:223-349) as a batched GPU workload: reentry vehicle tracking with
three Bayes-Sard quadrature Kalman filters (different expected model variances) and the UKF."""
from collections import OrderedDict

import numpy as np
import torch

from ..ssinf import BayesSardKalman, UnscentedKalman
from ..ssmod import ReentryVehicle2DTransition, Radar2DMeasurement
from ..utils import GaussRV
from . import scoring


def reentry_models():
    """Truth model, measurement model and the mis-specified filter model (bsq_tracking.py:230-261)."""
    m0 = np.array([6500, 350, -1.8, -6.8, 0.7])
    P0 = np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0])
    sys = ReentryVehicle2DTransition(GaussRV(5, m0, P0), GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 0])))
    obs = Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([6374, 0.0]))
    m0 = np.array([6500, 350, -1.1, -6.1, 0.7])
    P0 = np.diag([1e-6, 1e-6, 1e-6, 1e-6, 1])
    dyn = ReentryVehicle2DTransition(GaussRV(5, m0, P0), GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 1e-6])), dt=0.1)
    return sys, obs, dyn


def reentry_algorithms(dyn, obs):
    """bsq_tracking.py:263-281: the model variances are assigned after construction."""
    par_dyn = np.array([[1.0, 1, 1, 1, 1, 1]])
    par_obs = np.array([[1.0, 0.9, 0.9, 1e4, 1e4, 1e4]])
    mul_ut = np.hstack((np.zeros((dyn.dim_in, 1)), np.eye(dyn.dim_in), 2 * np.eye(dyn.dim_in))).astype(int)
    alg = OrderedDict({
        'bsqkf': BayesSardKalman(dyn, obs, par_dyn, par_obs, mul_ut, mul_ut, points='ut'),
        'bsqkf_2e-6': BayesSardKalman(dyn, obs, par_dyn, par_obs, mul_ut, mul_ut, points='ut'),
        'bsqkf_2e-7': BayesSardKalman(dyn, obs, par_dyn, par_obs, mul_ut, mul_ut, points='ut'),
        'ukf': UnscentedKalman(dyn, obs, beta=0.0),
    })
    alg['bsqkf'].tf_dyn.model.model_var = np.diag([0.0002, 0.0002, 0.0002, 0.0002, 0.0002])
    alg['bsqkf'].tf_obs.model.model_var = 0 * np.eye(2)
    alg['bsqkf_2e-6'].tf_dyn.model.model_var = 2e-6 * np.eye(5)
    alg['bsqkf_2e-6'].tf_obs.model.model_var = 0 * np.eye(2)
    alg['bsqkf_2e-7'].tf_dyn.model.model_var = 2e-7 * np.eye(5)
    alg['bsqkf_2e-7'].tf_obs.model.model_var = 0 * np.eye(2)
    return alg


def _block_scores(x, m, P, status, idx):
    """RMSE and inclination indicator against time of the sub-vector idx of the state (bsq_tracking.py:311-337):
    the sub-blocks of the estimates scored against the same sub-block of the MSE matrix."""
    i = torch.as_tensor(idx, device=x.device)
    xs, ms = x[i].contiguous(), m[i].contiguous()
    Ps = P[i][:, i].contiguous()
    r = scoring.score_pass(xs, ms, Ps, status, skip_first=False)
    d = len(idx)
    cnt = r['count']
    return (r['stats'][:, d + d * d + 1] / cnt).cpu().numpy(), (r['lcr'][:, 0] / cnt).cpu().numpy()


def reentry_demo(dur=200, mc_sims=100, x=None, y=None, keep_arrays=False, weights=None):
    """bsq_tracking.py:223-349 without the file output and the plots.  x (5, steps, mc), y (2, steps, mc): optional
    data (already sub-sampled); by default the truth is simulated on the device (Euler-Maruyama, dt = 0.05, every
    second point kept).  Returns the reference's result dict: 'alg_str' and, for 'state', 'position', 'velocity',
    'parameter', the arrays 'rmse' and 'inc' of shape (steps, n_alg); with keep_arrays also 'x', 'mean', 'cov'
    (lists of device tensors per algorithm).  weights: see below."""
    tau, disc_tau = 0.05, 0.1
    sys, obs, dyn = reentry_models()
    if x is None:
        x = sys.simulate_continuous(duration=dur, dt=tau, mc_sims=mc_sims, device_out=True)
        y = obs.simulate_measurements(x, device_out=True)
        x, y = x[:, ::2, :].contiguous(), y[:, ::2, :].contiguous()
    xd, yd = scoring.to_device(x), scoring.to_device(y)
    alg = reentry_algorithms(dyn, obs)
    if weights is not None:
        # quadrature weights assigned from outside (the pattern of research/tpq/tpq_ungm.py:114-124), e.g. the values of a
        # particular reference run: weights = {'dyn': (wm, Wc, Wcc), 'obs': (wm, Wc, Wcc)} for the three BSQ filters
        for name, a in alg.items():
            if name.startswith('bsqkf'):
                a.tf_dyn.wm, a.tf_dyn.Wc, a.tf_dyn.Wcc = weights['dyn']
                a.tf_obs.wm, a.tf_obs.Wc, a.tf_obs.Wcc = weights['obs']
    res = scoring.run_all(list(alg.values()), yd, smooth=False)
    parts = {'state': [0, 1, 2, 3, 4], 'position': [0, 1], 'velocity': [2, 3], 'parameter': [4]}
    out = {'duration': dur, 'disc_tau': disc_tau, 'alg_str': list(alg.keys())}
    for name, idx in parts.items():
        sc = [_block_scores(xd, r['mean_f'], r['cov_f'], r['status'], idx) for r in res]
        out[name] = {'rmse': np.stack([s[0] for s in sc], axis=1), 'inc': np.stack([s[1] for s in sc], axis=1)}
    out['n_failed'] = [int((r['status'] != 0).sum().item()) for r in res]
    if keep_arrays:
        out.update(x=xd, mean=[r['mean_f'] for r in res], cov=[r['cov_f'] for r in res])
    return out
