"""Random variables and performance scores (mirror of ssmtoybox/utils.py for the hot path).

Arrays follow the reference's convention (D, N, M, ...): D dimension, N time steps, M Monte-Carlo
trajectories.  All arithmetic runs on the GPU through the C-ABI library (K6 score kernels, Philox
sampler); there is no CPU fallback.  Reference lines: RandomVariable / GaussRV / StudentRV
utils.py:580-674, multivariate_t :349-382, squared_error :18-38, mse_matrix :41-64,
log_cred_ratio :67-120, neg_log_likelihood :123-148.
"""
import ctypes as C
from abc import ABCMeta, abstractmethod

import numpy as np
import torch

from . import _lib, device as dv
from ._lib import lib

_seed_state = {'seed': 0, 'calls': 0}


def seed(s):
    """Seed of the device Philox streams (the counterpart of np.random.seed for this package)."""
    _seed_state['seed'], _seed_state['calls'] = int(s), 0


def next_stream_seed():
    """A fresh 64-bit Philox key derived from the package seed; every sampling call consumes one."""
    _seed_state['calls'] += 1
    return (_seed_state['seed'] * 0x9E3779B97F4A7C15 + _seed_state['calls'] * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF


class RandomVariable(metaclass=ABCMeta):
    @abstractmethod
    def sample(self, size):
        pass

    @abstractmethod
    def get_stats(self):
        pass

    def _sample(self, size, dof, device_out=False):
        shape = (size,) if np.isscalar(size) else tuple(size)
        n = int(np.prod(shape))
        mean = dv._c(self.mean)
        factor = dv._cov_factor(self.cov if dof == 0.0 else self.scale)
        out = torch.empty((self.dim, n), dtype=torch.float64, device='cuda')
        rc = lib.ssm_sample(self.dim, dv._ptr(mean), dv._ptr(factor), float(dof), C.c_uint64(next_stream_seed()), 0,
                            dv._p(out), n, n, dv._stream())
        _lib.check(rc, 'ssm_sample')
        out = out.reshape((self.dim,) + shape)
        return out if device_out else out.cpu().numpy()


class GaussRV(RandomVariable):
    """Gaussian random variable (utils.py:580-623).  sample(size) -> (dim,) + size."""

    def __init__(self, dim, mean=None, cov=None):
        if mean is None:
            mean = np.zeros((dim, ))
        mean = np.atleast_1d(mean)
        if cov is None:
            cov = np.eye(dim)
        cov = np.atleast_2d(cov)
        self.dim = dim
        self.mean = mean
        self.cov = cov

    def sample(self, size, device_out=False):
        return self._sample(size, 0.0, device_out)

    def get_stats(self):
        return self.mean, self.cov


class StudentRV(RandomVariable):
    """Student's t random variable with scale matrix and dof (utils.py:626-674)."""

    def __init__(self, dim, mean=None, scale=None, dof=3.0):
        if mean is None:
            mean = np.zeros((dim,))
        mean = np.atleast_1d(mean)
        if scale is None:
            scale = np.eye(dim)
        scale = np.atleast_2d(scale)
        if dof <= 2.0:
            dof = 3.0
        self.dim = dim
        self.mean = mean
        self.scale = scale
        self.dof = dof

    def sample(self, size, device_out=False):
        return self._sample(size, self.dof, device_out)

    def get_stats(self):
        return self.mean, self.scale, self.dof


class GaussianMixtureRV(RandomVariable):
    """Gaussian-mixture random variable, the heavy-tailed data generator of the TPQ experiments (mirror of
    research/tpq/tpq_base.py:13-31; its sample() breaks on tuple sizes in the reference because utils.gauss_mixture
    returns a (samples, indexes) pair -- here sample(size) -> (dim,) + size for int and tuple sizes)."""

    def __init__(self, dim, means, covs, alphas):
        if len(means) != len(covs) or len(covs) != len(alphas):
            raise ValueError('Same number of means, covariances and mixture weights needs to be supplied!')
        if not np.isclose(np.sum(alphas), 1.0):
            raise ValueError('Mixture weights must sum to unity!')
        self.dim = dim
        self.means = means
        self.covs = covs
        self.alphas = alphas

    def sample(self, size, device_out=False):
        shape = (size,) if np.isscalar(size) else tuple(size)
        out, _ = _mixture(self.means, self.covs, self.alphas, int(np.prod(shape)), self.dim)
        out = out.reshape((self.dim,) + shape)
        return out if device_out else out.cpu().numpy()

    def get_stats(self):
        return self.means, self.covs, self.alphas


def _mixture(means, covs, alphas, n, dim):
    K = len(alphas)
    mu = dv._c(np.stack([np.atleast_1d(np.asarray(m, dtype=np.float64)).reshape(dim) for m in means]))
    F = dv._c(np.stack([dv._cov_factor(np.atleast_2d(c)) for c in covs]))
    al = dv._c(np.asarray(alphas, dtype=np.float64).reshape(K))
    out = torch.empty((dim, n), dtype=torch.float64, device='cuda')
    idx = torch.empty((n,), dtype=torch.int32, device='cuda')
    rc = lib.ssm_sample_mixture(dim, K, dv._ptr(mu), dv._ptr(F), dv._ptr(al), C.c_uint64(next_stream_seed()), 0,
                                dv._p(out), dv._p(idx), n, n, dv._stream())
    _lib.check(rc, 'ssm_sample_mixture')
    return out, idx


def gauss_mixture(means, covs, alphas, size):
    """Samples of a Gaussian mixture and the component each one came from -> (n, dim), (n,) (utils.py:261-301)."""
    if len(means) != len(covs) or len(covs) != len(alphas):
        raise ValueError('means, covs and alphas need to have the same length.')
    n = int(np.prod(size))
    out, idx = _mixture(means, covs, alphas, n, len(np.atleast_1d(means[0])))
    return out.T.cpu().numpy(), idx.cpu().numpy().astype(int)


def multivariate_t(mean, scale, nu, size):
    """Samples of a multivariate Student's t-distribution -> (size, dim) (utils.py:349-382)."""
    mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
    return StudentRV(mean.shape[0], mean, scale, nu).sample(int(size)).T


def bootstrap_var(data, samples=1000):
    """Bootstrap estimate of the variance of the sample mean (utils.py:223-244), resampled on the device."""
    d = torch.as_tensor(np.ascontiguousarray(np.asarray(data, dtype=np.float64).squeeze()), device='cuda')
    return float(dv.bootstrap_var(d, int(samples), seed=next_stream_seed()))


# ------------------------------------------------------------------------------------------------
# scores
# ------------------------------------------------------------------------------------------------
def _dev(a):
    if isinstance(a, torch.Tensor):
        return a.to(device='cuda', dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device='cuda')


def squared_error(x, m):
    """(x - m)**2, broadcasting like the reference (utils.py:18-38)."""
    xd, md = _dev(x), _dev(m)
    out = (xd - md) ** 2
    return out if isinstance(x, torch.Tensor) else out.cpu().numpy()


def mse_matrix(x, m):
    """Sample MSE matrix over MC simulations: x (d, 1) or (d, M), m (d, M) -> (d, d) (utils.py:41-64)."""
    m = np.asarray(m, dtype=np.float64)
    d, M = m.shape
    x = np.broadcast_to(np.asarray(x, dtype=np.float64).reshape(d, -1), (d, M))
    xd, md = _dev(x[:, None, :]), _dev(m[:, None, :])
    cov = torch.eye(d, dtype=torch.float64, device='cuda')[:, :, None, None].expand(d, d, 1, M).contiguous()
    stats, _ = dv.scores_phase1(xd, md, cov, want_rmse_acc=False)
    st = stats.cpu().numpy()[0]
    return st[d:d + d * d].reshape(d, d) / M


def neg_log_likelihood(x, m, P):
    """0.5 (log|P| + (x-m)' P^-1 (x-m) + d log 2 pi) for one estimate (utils.py:123-148)."""
    x, m = np.atleast_1d(np.asarray(x, dtype=np.float64)), np.atleast_1d(np.asarray(m, dtype=np.float64))
    d = x.shape[0]
    P = np.asarray(P, dtype=np.float64).reshape(d, d)
    stats, _ = dv.scores_phase1(_dev(x.reshape(d, 1, 1)), _dev(m.reshape(d, 1, 1)), _dev(P.reshape(d, d, 1, 1)),
                                want_rmse_acc=False)
    return float(stats.cpu().numpy()[0, d + d * d])


def log_cred_ratio(x, m, P, MSE):
    """10 (log10 dx'P^-1dx - log10 dx'MSE^-1dx) for one estimate (utils.py:67-120)."""
    x, m = np.atleast_1d(np.asarray(x, dtype=np.float64)), np.atleast_1d(np.asarray(m, dtype=np.float64))
    d = x.shape[0]
    P = np.asarray(P, dtype=np.float64).reshape(d, d)
    MSE = np.asarray(MSE, dtype=np.float64).reshape(d, d)
    lcr = dv.scores_phase2(_dev(x.reshape(d, 1, 1)), _dev(m.reshape(d, 1, 1)), _dev(P.reshape(d, d, 1, 1)),
                           _dev(MSE.reshape(d, d, 1)))
    return float(lcr.cpu().numpy()[0, 0])


def evaluate_performance(x, mean, cov, status=None, comm=None, to_host=True, phase1=None, quad=None):
    """Batched RMSE / NCI / NLL of one filter over all trajectories, aggregated exactly like
    research/gpq/icinco_demo.py:17-52 (RMSE = trajectory-mean of sqrt(time-mean SE); NCI and NLL skip
    k = 0 but divide by N), computed on the device in two reduction phases.
    x, mean (dx, N, M); cov (dx, dx, N, M); status (M,) int32 or None (failed trajectories are excluded).
    comm: optional ssmtoybox_b200.dist.Communicator -- trajectories are then sharded over ranks and the
    packed statistics are all-reduced (one NCCL call per phase).
    phase1: optional (stats, rmse_acc) already accumulated in-kernel by the smoother (device.smooth_backward(...,
    x_truth=x)); the first reduction pass over the arrays is then skipped.
    quad: optional (N, M) array of d' P^-1 d stored by the same call with want_quad=True; the second phase then does not
    read the covariances again.
    Returns dict(rmse (dx,), nci, nll, inc (inclination), mse (dx,dx,N), rmse_vs_time (N,), n_ok); device
    tensors instead of numpy / floats when to_host=False (no synchronisation)."""
    xd, md, Pd = _dev(x), _dev(mean), _dev(cov)
    dx, N, M = xd.shape
    if phase1 is not None:
        stats, acc = phase1
    else:   # the first pass keeps d' P^-1 d per unit, so the second one does not read the covariances again
        quad = torch.empty((N, M), dtype=torch.float64, device=xd.device)
        stats, acc = dv.scores_phase1(xd, md, Pd, status, quad=quad)
    return finish_scores(stats, acc, status, lambda mse: dv.scores_phase2(xd, md, Pd, mse, status, quad=quad),
                         comm=comm, to_host=to_host)


def finish_scores(stats, acc, status, second_phase, comm=None, to_host=True):
    """Scores from the first-phase statistics of this rank: stats (N, W) packed rows, acc (dx, M) per-trajectory
    time-sums of the squared error, status (M,) or None.  second_phase(mse (dx, dx, N)) -> (N, 2) sums of the log
    credibility ratio, called once the GLOBAL per-step MSE matrix is known (research/gpq/icinco_demo.py:34-40).
    One all-reduce per phase when comm is given."""
    N, W = stats.shape
    dx, M = acc.shape
    ok = torch.ones(M, dtype=torch.bool, device=stats.device) if status is None else (status == 0)
    # per-trajectory sqrt(time-mean SE), summed over the trajectories that completed
    rm = torch.where(ok[None, :], torch.sqrt(acc / N), torch.zeros_like(acc)).sum(dim=1)
    pack = torch.cat([stats.reshape(-1), rm])
    if comm is not None:
        pack = comm.allreduce_sum(pack)
    st = pack[:N * W].reshape(N, W)
    rm = pack[N * W:]
    cnt = st[:, -1]
    n_ok = cnt[0]
    mse = (st[:, dx:dx + dx * dx] / cnt[:, None]).T.reshape(dx, dx, N).contiguous()
    lcr = second_phase(mse)
    if comm is not None:
        lcr = comm.allreduce_sum(lcr)
    out = dict(rmse=(rm / n_ok), nll=st[1:, dx + dx * dx].sum() / (N * n_ok), inc=lcr[1:, 0].sum() / (N * n_ok),
               nci=lcr[1:, 0].sum() / (N * n_ok), abs_nci=lcr[1:, 1].sum() / (N * n_ok), mse=mse,
               rmse_vs_time=st[:, dx + dx * dx + 1] / cnt, n_ok=n_ok)
    if not to_host:
        return out
    return {k: (v.cpu().numpy() if v.ndim else float(v)) for k, v in out.items()}


def evaluate_scored(sc, comm=None, to_host=True):
    """Scores from the outputs of a scoring pass that kept no moment arrays (device.smooth_scores / filter_scored):
    sc = dict(stats, rmse_acc, quad (N, M), dres (dx, N, M), status).  Same result as evaluate_performance on the
    arrays that pass did not store."""
    return finish_scores(sc['stats'], sc['rmse_acc'], sc['status'],
                         lambda mse: dv.scores_phase2_res(sc['dres'], sc['quad'], mse, sc['status']), comm=comm, to_host=to_host)
