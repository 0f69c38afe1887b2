"""research/tpq/tpq_base.py `run_filters` (:175-192) and `eval_perf_scores` (:154-172) as batched GPU workloads."""
import numpy as np

from . import scoring


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
        r = scoring.score_pass(xd, m, P, None, skip_first=False, reg=1e-6 * np.eye(xD))
        rmse.append((r['stats'][:, xD + xD * xD + 1] / r['count']).cpu().numpy())
        lcr.append((r['lcr'][:, 0] / r['count']).cpu().numpy())
    return np.stack(rmse, axis=1), np.stack(lcr, axis=1)
