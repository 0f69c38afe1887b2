"""Sigma-point moment transforms (mirror of ssmtoybox/mtran.py:11-578 for the rules on the hot path).

Point sets and classical weights are small host-side constant tables; `apply` runs on the GPU through
ssm_transform_apply when the integrand is the dyn_eval / meas_eval of a device model.  Arbitrary
Python callables cannot run on the device and there is no CPU fallback: they raise NotImplementedError.
"""
import ctypes as C
from abc import ABCMeta, abstractmethod

import numpy as np
import torch
from numpy.polynomial.hermite_e import hermegauss, hermeval

from . import _lib, device as dv
from ._lib import lib


def _integrand_of(f):
    """Map a bound dyn_eval / meas_eval / dyn_fcn-style method to (which, model id, dim_state, si, par, dim_out)."""
    from .ssmod import TransitionModel, MeasurementModel
    owner = getattr(f, '__self__', None)
    name = getattr(f, '__name__', '')
    if isinstance(owner, TransitionModel) and name == 'dyn_eval' and owner._device_id is not None:
        return 0, owner._device_id, owner.dim_state, (0, 0), owner._par(), owner.dim_state
    if isinstance(owner, MeasurementModel) and name == 'meas_eval' and owner._device_id is not None:
        return 1, owner._device_id, owner.dim_state, owner._si(), owner._par(), owner.dim_out
    raise NotImplementedError(
        'MomentTransform.apply runs on the GPU only: the integrand must be the dyn_eval / meas_eval method of a '
        'state-space model with a device implementation (got {!r}); there is no CPU fallback'.format(f))


def _apply_device(tf_dict, f, mean, cov, fcn_pars):
    """Shared implementation of SigmaPointTransform.apply / BQTransform.apply."""
    which, model_id, dim_state, si, par, dim_out = _integrand_of(f)
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    batched = mean.ndim == 2
    m2 = np.ascontiguousarray(mean.reshape(mean.shape[0], -1))
    D, n = m2.shape
    c2 = np.ascontiguousarray(cov.reshape(D, D, -1))
    keep = []
    d = dict(tf_dict)
    d['t_dim_out_fn'] = dim_out
    t = dv._lower_transform(d, 't_', keep)
    mt, ct = torch.as_tensor(m2, device='cuda'), torch.as_tensor(c2, device='cuda')
    kw = dict(dtype=torch.float64, device='cuda')
    mf, cf, cfx = torch.empty((dim_out, n), **kw), torch.empty((dim_out, dim_out, n), **kw), torch.empty((dim_out, D, n), **kw)
    status = torch.empty((n,), dtype=torch.int32, device='cuda')
    time = float(np.asarray(fcn_pars).reshape(-1)[0]) if fcn_pars is not None and np.size(fcn_pars) else 0.0
    p = (C.c_double * 8)(*(list(par) + [0.0] * (8 - len(par))))
    rc = lib.ssm_transform_apply(which, model_id, dim_state, si[0], si[1], p, C.byref(t), time, dv._p(mt), dv._p(ct),
                                 dv._p(mf), dv._p(cf), dv._p(cfx), dv._p(status), n, n, dv._stream())
    _lib.check(rc, 'ssm_transform_apply')
    st = status.cpu().numpy()
    if not batched and st[0] != 0:
        raise np.linalg.LinAlgError('Matrix is not positive definite')  # numpy.linalg.cholesky, mtran.py:139
    mf, cf, cfx = mf.cpu().numpy(), cf.cpu().numpy(), cfx.cpu().numpy()
    if batched:
        return mf, cf, cfx
    return mf[:, 0], cf[:, :, 0], cfx[:, :, 0]


class MomentTransform(metaclass=ABCMeta):
    """Base class for all moment transforms (mtran.py:11-46)."""

    @abstractmethod
    def apply(self, f, mean, cov, fcn_pars, tf_pars=None):
        pass


class SigmaPointTransform(MomentTransform):
    """Base class of sigma-point transforms (mtran.py:102-149): x = m + chol(P) U, centred moments."""

    def apply(self, f, mean, cov, fcn_pars, tf_pars=None):
        d = {'t_kind': 'sp', 't_points': self.unit_sp, 't_wm': self.wm, 't_Wc': self.Wc}
        return _apply_device(d, f, mean, cov, fcn_pars)

    def _tf_dict(self, prefix):
        return {prefix + 'kind': 'sp', prefix + 'points': self.unit_sp, prefix + 'wm': self.wm, prefix + 'Wc': self.Wc}


class SphericalRadialTransform(SigmaPointTransform):
    """Spherical-radial (cubature) rule, 2*dim points (mtran.py:152-204)."""

    def __init__(self, dim):
        self.wm = self.weights(dim)
        self.Wc = np.diag(self.wm)
        self.unit_sp = self.unit_sigma_points(dim)

    @staticmethod
    def weights(dim):
        return (1 / (2.0 * dim)) * np.ones(2 * dim)

    @staticmethod
    def unit_sigma_points(dim):
        c = np.sqrt(dim)
        return np.hstack((c * np.eye(dim), -c * np.eye(dim)))


class UnscentedTransform(SigmaPointTransform):
    """Unscented transform, 2*dim+1 points (mtran.py:207-293)."""

    def __init__(self, dim, kappa=None, alpha=1.0, beta=2.0):
        self.wm, self.wc = self.weights(dim, kappa=kappa, alpha=alpha, beta=beta)
        self.Wm = np.diag(self.wm)
        self.Wc = np.diag(self.wc)
        self.unit_sp = self.unit_sigma_points(dim, kappa=kappa, alpha=alpha)

    @staticmethod
    def unit_sigma_points(dim, kappa=None, alpha=1.0):
        kappa = np.max([3.0 - dim, 0.0]) if kappa is None else kappa
        lam = alpha ** 2 * (dim + kappa) - dim
        c = np.sqrt(dim + lam)
        return np.hstack((np.zeros((dim, 1)), c * np.eye(dim), -c * np.eye(dim)))

    @staticmethod
    def weights(dim, kappa=None, alpha=1.0, beta=2.0):
        kappa = np.max([3.0 - dim, 0.0]) if kappa is None else kappa
        lam = alpha ** 2 * (dim + kappa) - dim
        wm = 1.0 / (2.0 * (dim + lam)) * np.ones(2 * dim + 1)
        wc = wm.copy()
        wm[0] = lam / (dim + lam)
        wc[0] = wm[0] + (1 - alpha ** 2 + beta)
        return wm, wc


def _cartesian(arrays):
    grids = np.meshgrid(*arrays, indexing='ij')
    return np.stack([g.reshape(-1) for g in grids], axis=1)


def _factorial(n):
    r = 1.0
    for k in range(2, int(n) + 1):
        r *= k
    return r


class GaussHermiteTransform(SigmaPointTransform):
    """Gauss-Hermite rule, degree**dim points (mtran.py:296-360).  The weights are NOT hermegauss'
    (mtran.py:334-336)."""

    def __init__(self, dim, degree=3):
        self.degree = degree
        self.wm = self.weights(dim, degree)
        self.Wc = np.diag(self.wm)
        self.unit_sp = self.unit_sigma_points(dim, degree)

    @staticmethod
    def weights(dim, degree=3):
        x, w = hermegauss(degree)
        w = _factorial(degree) / (degree ** 2 * hermeval(x, [0] * (degree - 1) + [1]) ** 2)
        return np.prod(_cartesian([w] * dim), axis=1)

    @staticmethod
    def unit_sigma_points(dim, degree=3):
        x, w = hermegauss(degree)
        return _cartesian([x] * dim).T


class FullySymmetricStudentTransform(SigmaPointTransform):
    """Fully symmetric rule for Student-t inputs, degrees 3 and 5 (mtran.py:363-578)."""

    _supported_degrees_ = [3, 5]

    def __init__(self, dim, degree=3, kappa=None, dof=4):
        self.degree, self.kappa, self.dof = degree, kappa, dof
        self.wm = self.weights(dim, degree, kappa, dof)
        self.Wc = np.diag(self.wm)
        self.unit_sp = self.unit_sigma_points(dim, degree, kappa, dof)

    @staticmethod
    def weights(dim, degree=3, kappa=None, dof=4.0):
        if degree not in FullySymmetricStudentTransform._supported_degrees_:
            degree = 3
        kappa = np.max([3.0 - dim, 0.0]) if kappa is None else kappa
        dof = np.max((dof, degree))
        if degree == 3:
            w = 1 / (2 * (dim + kappa)) * np.ones(2 * dim + 1)
            w[0] = kappa / (dim + kappa)
            return w
        I2 = dof / (dof - 2)
        I22 = dof ** 2 / ((dof - 2) * (dof - 4))
        I4 = 3 * I22
        A0 = 1 - dim * (I2 / I4) ** 2 * (I4 - 0.5 * (dim - 1) * I22)
        A1 = 0.5 * (I2 / I4) ** 2 * (I4 - (dim - 1) * I22)
        A11 = 0.25 * (I2 / I4) ** 2 * I22
        return np.hstack((A0, A1 * np.ones(2 * dim), A11 * np.ones(2 * dim * (dim - 1))))

    @staticmethod
    def unit_sigma_points(dim, degree=3, kappa=None, dof=4.0):
        if degree not in FullySymmetricStudentTransform._supported_degrees_:
            degree = 3
        kappa = np.max([3.0 - dim, 0.0]) if kappa is None else kappa
        dof = np.max((dof, degree))
        if degree == 3:
            I2 = dof / (dof - 2)
            u = np.sqrt(I2 * (dim + kappa))
            return u * np.hstack((np.zeros((dim, 1)), np.eye(dim), -np.eye(dim)))
        I2 = dof / (dof - 2)
        I4 = 3 * dof ** 2 / ((dof - 2) * (dof - 4))
        u = np.sqrt(I4 / I2)
        sp0 = FullySymmetricStudentTransform.symmetric_set(dim, [])
        sp1 = FullySymmetricStudentTransform.symmetric_set(dim, [u])
        sp2 = FullySymmetricStudentTransform.symmetric_set(dim, [u, u])
        return np.hstack((sp0, sp1, sp2))

    @staticmethod
    def symmetric_set(dim, gen):
        """Fully symmetric point set of a generator with one or two equal entries (mtran.py:522-578):
        [] -> origin; [u] -> +-u e_i; [u, u] -> (+-u e_i +- u e_j), i < j, in the reference's order."""
        gen = list(np.atleast_1d(gen)) if np.size(gen) else []
        if len(gen) == 0:
            return np.zeros((dim, 1))
        eye = np.eye(dim)
        cols = []
        if len(gen) == 1:
            for i in range(dim):
                cols += [gen[0] * eye[i], -gen[0] * eye[i]]
        elif len(gen) == 2 and abs(gen[0] - gen[1]) < np.spacing(1.0):
            for i in range(dim):
                for j in range(i + 1, dim):
                    for s in (1.0, -1.0):
                        v = gen[0] * eye[i] + s * gen[1] * eye[j]
                        cols += [v, -v]
        else:
            raise NotImplementedError('generators with unequal or more than two entries are not used by the '
                                      'degree 3 / 5 rules')
        if not cols:
            return np.empty((dim, 0))
        return np.stack(cols, axis=1)
