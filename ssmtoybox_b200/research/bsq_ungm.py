"""research/bsq/bsq_ungm.py `tables` (:91-186) as a batched GPU workload."""
import numpy as np
import pandas as pd

from ..ssinf import UnscentedKalman, GaussHermiteKalman, GaussianProcessKalman, BayesSardKalman
from ..ssmod import UNGMTransition, UNGMMeasurement
from ..utils import GaussRV
from . import scoring
from .scoring import evaluate_performance  # noqa: F401  (bsq_ungm.py:27)


def algorithms(dyn, obs):
    """Classical, GPQ and BSQ filters on UT / GH-5 / GH-7 points (bsq_ungm.py:100-122)."""
    par_ut = np.array([[3.0, 0.3]])
    par_gh5 = np.array([[5.0, 0.6]])
    par_gh7 = np.array([[3.0, 0.4]])
    mulind_ut = np.array([[0, 1, 2]])
    mulind_gh = lambda degree: np.atleast_2d(np.arange(degree))  # noqa: E731
    return (
        UnscentedKalman(dyn, obs, alpha=1.0, beta=0.0),
        GaussHermiteKalman(dyn, obs, deg=5),
        GaussHermiteKalman(dyn, obs, deg=7),
        GaussianProcessKalman(dyn, obs, par_ut, par_ut, kernel='rbf', points='ut', point_hyp={'alpha': 1.0}),
        GaussianProcessKalman(dyn, obs, par_gh5, par_gh5, kernel='rbf', points='gh', point_hyp={'degree': 5}),
        GaussianProcessKalman(dyn, obs, par_gh7, par_gh7, kernel='rbf', points='gh', point_hyp={'degree': 7}),
        BayesSardKalman(dyn, obs, par_ut, par_ut, mulind_ut, mulind_ut, points='ut', point_hyp={'alpha': 1.0}),
        BayesSardKalman(dyn, obs, par_gh5, par_gh5, mulind_gh(5), mulind_gh(5), points='gh', point_hyp={'degree': 5}),
        BayesSardKalman(dyn, obs, par_gh7, par_gh7, mulind_gh(7), mulind_gh(7), points='gh', point_hyp={'degree': 7}),
    )


def tables(steps=500, mc=100, x=None, z=None, bootstrap_variance=True, num_bs_samples=10000):
    """bsq_ungm.py:91-186: six tables (rows UT, GH-5, GH-7; columns Classical, GPQ, BSQ and their `2 std`)."""
    dyn = UNGMTransition(GaussRV(1, cov=5.0), GaussRV(1, cov=10.0))
    obs = UNGMMeasurement(GaussRV(1, cov=1.0), 1)
    if x is None:
        x = dyn.simulate_discrete(steps, mc, device_out=True)
    if z is None:
        z = obs.simulate_measurements(x, device_out=True)
    res = scoring.run_all(algorithms(dyn, obs), z)
    sc = evaluate_performance(x, [r['mean_f'] for r in res], [r['cov_f'] for r in res], [r['mean_s'] for r in res],
                              [r['cov_s'] for r in res], bootstrap_variance, num_bs_samples, status=[r['status'] for r in res])
    mean = sc[:6]
    std = sc[6:] if bootstrap_variance else [np.zeros((len(res), 1))] * 6
    row_labels = ['UT', 'GH-5', 'GH-7']
    n = len(row_labels)
    col_labels = ['Classical', 'GPQ', 'BSQ', 'Classical (2std)', 'GPQ (2std)', 'BSQ (2std)']
    keys = ('filter_RMSE', 'filter_NCI', 'filter_NLL', 'smoother_RMSE', 'smoother_NCI', 'smoother_NLL')
    return {k: pd.DataFrame(np.hstack((m.reshape(3, n).T, s.reshape(3, n).T)), index=row_labels, columns=col_labels)
            for k, m, s in zip(keys, mean, std)}
