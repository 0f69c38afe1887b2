"""BASELINE configuration C5: NCI calibration sweep of the Bayes-Sard quadrature Kalman filter.

The reference calibrates the BSQ filters by assigning the expected model variance from outside and reading the
non-credibility index of a Monte-Carlo batch (research/bsq/bsq_tracking.py:266-281, 311-337; set-ups
ssmtoybox/tests/test_ssinf.py:42-51 (pendulum), :65-78 (coordinated turn) with the radar of
research/tpq/synthetic.py:2126; filter construction ssmtoybox/tests/test_ssinf.py:195-203: unit kernel parameters,
UT points, multi-index [0 | I | 2I]).  Here one sweep point = one generator-driven run of mc.monte_carlo_scores:
simulate -> BSQ filter with in-kernel scoring, chunk by chunk, no per-trajectory moment arrays; under torchrun the
trajectories are sharded over the ranks and the statistics all-reduced once per phase.
"""
import time

import numpy as np
import torch

from .. import mc
from ..ssinf import BayesSardKalman
from ..ssmod import Pendulum2DTransition, Pendulum2DMeasurement, CoordinatedTurnTransition, Radar2DMeasurement
from ..utils import GaussRV

N_STEPS = 100          # trajectory length of the reference's fixtures (simulate_discrete(100), tests/test_ssinf.py:49, 78)


def mul_ut(d):
    """Multi-index of the polynomial mean [0 | I | 2I] (tests/test_ssinf.py:198-201)."""
    return np.hstack((np.zeros((d, 1)), np.eye(d), 2 * np.eye(d))).astype(int)


def pendulum_model():
    """tests/test_ssinf.py:42-51"""
    dt = 0.01
    x0 = GaussRV(2, mean=np.array([1.5, 0]), cov=0.01 * np.eye(2))
    q = GaussRV(2, cov=0.01 * np.array([[(dt ** 3) / 3, (dt ** 2) / 2], [(dt ** 2) / 2, dt]]))
    dyn = Pendulum2DTransition(x0, q, dt=dt)
    obs = Pendulum2DMeasurement(GaussRV(1, cov=np.array([[0.1]])), dyn.dim_state)
    return dyn, obs


def coordinated_turn_model(dt=0.1):
    """Dynamics of tests/test_ssinf.py:65-78 with the radar of research/tpq/synthetic.py:2126 (state_index [0, 2])."""
    m0 = np.array([1000, 300, 1000, 0, np.deg2rad(-3.0)])
    P0 = np.diag([100, 10, 100, 10, 0.1])
    rho_1, rho_2 = 0.1, 1.75e-4
    A = np.array([[dt ** 3 / 3, dt ** 2 / 2], [dt ** 2 / 2, dt]])
    Q = np.zeros((5, 5))
    Q[:2, :2], Q[2:4, 2:4], Q[4, 4] = rho_1 * A, rho_1 * A, rho_2 * dt
    dyn = CoordinatedTurnTransition(GaussRV(5, m0, P0), GaussRV(5, cov=Q), dt=dt)
    obs = Radar2DMeasurement(GaussRV(2, cov=np.diag([100, 10e-6])), 5, state_index=[0, 2])
    return dyn, obs


MODELS = {'pendulum': pendulum_model, 'coordturn': coordinated_turn_model}
# expected-model-variance grid of the dynamics transform (the measurement transform's is set to 0 like
# bsq_tracking.py:277-281); None = the value the BSQ weights themselves give (no assignment)
MODEL_VAR = {'pendulum': (None, 1e-1, 1e-2, 1e-3, 1e-4, 0.0), 'coordturn': (None, 1e-1, 1e-2, 1e-3, 1e-4, 0.0)}


def build_filter(model, model_var=None):
    dyn, obs = MODELS[model]()
    d = dyn.dim_in
    kp = np.atleast_2d(np.ones(d + 1))
    alg = BayesSardKalman(dyn, obs, kp, kp, mul_ut(d), mul_ut(d), points='ut')
    if model_var is not None:       # research/bsq/bsq_tracking.py:276-281
        alg.tf_dyn.model.model_var = float(model_var) * np.eye(dyn.dim_state)
        alg.tf_obs.model.model_var = 0.0 * np.eye(obs.dim_out)
    return alg


def bsq_nci_sweep(model='pendulum', mc_sims=(1000,), model_var=None, n_steps=N_STEPS, seed=0, chunk=1 << 18, comm=None,
                  smooth=False, verbose=False):
    """NCI / RMSE / NLL of the BSQ filter on `model` for every (trajectory count, expected model variance) pair.
    Returns a list of dicts: model, mc_sims, model_var, nci, abs_nci, rmse (dx,), nll, n_failed, seconds,
    traj_steps_per_s (whole job: all ranks), kept_bytes (this rank)."""
    grid = MODEL_VAR[model] if model_var is None else tuple(model_var)
    rows = []
    for M in mc_sims:
        for mv in grid:
            alg = build_filter(model, mv)
            torch.cuda.synchronize()
            if comm is not None:
                comm.barrier()
            t0 = time.perf_counter()
            r = mc.monte_carlo_scores(alg, int(M), n_steps, seed=seed, chunk=chunk, smooth=smooth, comm=comm)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if comm is not None:
                dt = comm.allreduce_max(dt)
            row = dict(model=model, mc_sims=int(M), model_var=mv, nci=r['nci'], abs_nci=r['abs_nci'], rmse=np.asarray(r['rmse']).tolist(),
                       nll=r['nll'], n_failed=int(r['n_failed']), seconds=dt, traj_steps_per_s=int(M) * n_steps / dt,
                       kept_bytes=r['kept_bytes'])
            rows.append(row)
            if verbose and (comm is None or comm.rank == 0):
                print('{model:10s} M={mc_sims:<9d} model_var={model_var!s:8s} NCI {nci:+8.4f}  RMSE {r0:9.4f}  NLL {nll:10.4f}  '
                      'failed {n_failed:<7d} {seconds:7.3f} s  {traj_steps_per_s:.3e} traj-steps/s'.format(r0=row['rmse'][0], **row), flush=True)
    return rows
