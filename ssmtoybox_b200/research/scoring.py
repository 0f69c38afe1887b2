"""Device-side evaluation shared by the research drivers.

`evaluate_performance` is the function of the same name in research/gpq/icinco_demo.py:17-71 and
research/bsq/bsq_ungm.py:27-84 (identical there): simulation-average of the time-averaged RMSE, NCI and NLL of
filter and smoother for several algorithms, optionally with twice the bootstrap standard deviation.  The Python
loops over (step, algorithm, simulation) become two reduction passes per algorithm (K6) plus one bootstrap kernel
per score; the per-simulation data that the reference keeps in (1, step, sim, alg) arrays are the per-trajectory
accumulators of ssm_scores_phase1_traj / ssm_scores_phase2_traj.
"""
import numpy as np
import torch

from .. import device as dv
from .._lib import lib


def to_device(a):
    if isinstance(a, torch.Tensor):
        return a.to(device='cuda', dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device='cuda')


def run_all(algorithms, z, smooth=True):
    """The filter / smoother loop of icinco_demo.py:115-125, bsq_ungm.py:131-137, tpq_base.py:175-192: one batched
    forward (and backward) pass per algorithm over all trajectories of z (dy, N, M), each trajectory starting from
    the model's initial moments (= the reference's reset() after every simulation).
    Returns a list of dicts of device tensors: mean_f, cov_f, (mean_s, cov_s,) status."""
    zd = to_device(z)
    out = []
    for alg in algorithms:
        alg.reset()
        mf, Pf = alg.forward_pass(zd)
        r = dict(mean_f=mf, cov_f=Pf, status=alg.status)
        if smooth:
            ms, Ps = alg.backward_pass()
            r.update(mean_s=ms, cov_s=Ps, status=alg.status)
        alg.reset()
        out.append(r)
    return out


def score_pass(x, m, P, status=None, mse=None, skip_first=True, reg=None):
    """Both reduction phases for one set of estimates.  skip_first: the per-trajectory NLL / NCI sums leave out
    k = 0 like the loops `for k in range(1, num_step)` of icinco_demo.py:32 (the time MEAN still divides by N).
    mse: per-step MSE matrices (dx, dx, N) to use in the credibility ratio instead of this set's own (the reference
    scores the smoother against the FILTER's MSE matrix, icinco_demo.py:35); reg is added to them
    (tpq_base.py:161).
    Returns device tensors: stats (N, W), lcr (N, 2), mse (dx, dx, N), rmse_data (dx, M), nll_data (M,),
    nci_data (M,), ok (M,) bool."""
    dx, N, M = x.shape
    W = lib.ssm_scores_width(dx)
    kw = dict(dtype=torch.float64, device=x.device)
    stats, acc = torch.empty((N, W), **kw), torch.empty((dx, M), **kw)
    nll, nci, lcr = torch.zeros(M, **kw), torch.zeros(M, **kw), torch.zeros((N, 2), **kw)
    k1 = 1 if skip_first else 0
    if k1:
        dv.scores_phase1(x, m, P, status, window=(0, 1), out=(stats, acc))
    quad = torch.empty((N, M), **kw)       # d' P^-1 d per unit, kept by the first pass for the second one
    if N > k1:
        dv.scores_phase1(x, m, P, status, window=(k1, N), out=(stats, acc), nll_acc=nll, quad=quad)
    cnt = stats[:, -1]
    own = (stats[:, dx:dx + dx * dx] / cnt[:, None]).T.reshape(dx, dx, N).contiguous()
    used = own if mse is None else mse
    if reg is not None:
        used = used + to_device(reg)[:, :, None]
    if N > k1:
        dv.scores_phase2(x, m, P, used, status, window=(k1, N), out=lcr, lcr_acc=nci, quad=quad)
    ok = torch.ones(M, dtype=torch.bool, device=x.device) if status is None else (status == 0)
    return dict(stats=stats, lcr=lcr, mse=own, rmse_data=torch.sqrt(acc / N), nll_data=nll / N, nci_data=nci / N, ok=ok,
                count=cnt)


def _mean_ok(data, ok):
    """mean over the trajectories that completed (last axis)"""
    return torch.where(ok, data, torch.zeros_like(data)).sum(dim=-1) / ok.sum()


def _split_algs(a, ndim_alg):
    """(…, A) array or list of per-algorithm arrays -> list of device tensors"""
    if isinstance(a, (list, tuple)):
        return [to_device(v) for v in a]
    if a.ndim != ndim_alg:
        raise ValueError('expected an array with a trailing algorithm axis')
    return [to_device(a[..., i]) for i in range(a.shape[-1])]


def evaluate_performance(x, mean_f, cov_f, mean_s, cov_s, bootstrap_variance=True, num_bs_samples=10000, status=None,
                         seed=0):
    """icinco_demo.py:17-71 / bsq_ungm.py:27-84.  x (dim, N, M); mean_* (dim, N, M, A), cov_* (dim, dim, N, M, A) as
    numpy / torch arrays, or lists of A per-algorithm arrays (device tensors are used in place).
    status: optional list of A (M,) int32 device tensors; failed trajectories are left out of every average (the
    reference would have stopped with an exception).
    Returns rmseMean_f, nciMean_f, nllMean_f, rmseMean_s, nciMean_s, nllMean_s [, the six `2 std` arrays] as numpy
    arrays shaped like the reference's: means (A, dim) / (A, 1), standard deviations (A, 1)."""
    xd = to_device(x)
    mf, Pf, ms, Ps = _split_algs(mean_f, 4), _split_algs(cov_f, 5), _split_algs(mean_s, 4), _split_algs(cov_s, 5)
    A = len(mf)
    dim = xd.shape[0]
    names = ('rmse_data', 'nci_data', 'nll_data')
    means = [[None] * A for _ in range(6)]
    stds = [[None] * A for _ in range(6)]
    for a in range(A):
        st = None if status is None else status[a]
        f = score_pass(xd, mf[a], Pf[a], st)
        s = score_pass(xd, ms[a], Ps[a], st, mse=f['mse'])  # smoother NCI against the FILTER's MSE matrix (Q13)
        for j, r in enumerate((f, s)):
            for i, nm in enumerate(names):
                means[3 * j + i][a] = _mean_ok(r[nm], r['ok']).reshape(-1)
                if bootstrap_variance:
                    data = r[nm].reshape(-1, r[nm].shape[-1])
                    if data.shape[0] != 1:  # np.random.choice(data, ...) in utils.py:238 needs 1-D data
                        raise ValueError('a must be 1-dimensional: bootstrap_var works on scalar states only (utils.py:236-240)')
                    data = data[0][r['ok']]
                    stds[3 * j + i][a] = 2.0 * torch.sqrt(dv.bootstrap_var(data, num_bs_samples, seed=seed + 6 * a + 3 * j + i)).reshape(1)
    out = [torch.stack(m).cpu().numpy().reshape(A, -1) for m in means]
    if bootstrap_variance:
        out += [torch.stack(s).cpu().numpy().reshape(A, 1) for s in stds]
    return tuple(out)
