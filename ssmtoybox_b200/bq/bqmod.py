"""Integrand models of Bayesian quadrature (mirror of ssmtoybox/bq/bqmod.py: Model :15-106, 340-423,
GaussianProcessModel :426-523, BayesSardModel :599-992, StudentTProcessModel :1055-1160).

bq_weights runs on the GPU (ssm_bq_weights, one CTA per kernel-parameter vector), and so does the objective of the
hyper-parameter fit (neg_log_marginal_likelihood -> ssm_gp_nlml, batched over log-parameter vectors); the optimiser
itself is scipy.optimize.minimize on the host, as in the reference (bqmod.py:250-285).  predict and plotting are
outside the hot path."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .. import device as dv
from ..mtran import SphericalRadialTransform, UnscentedTransform, GaussHermiteTransform, FullySymmetricStudentTransform
from .bqkern import RBFGauss, RBFStudent


# ------------------------------------------------------------------------------------------------
# arithmetic of the quadrature weights
# ------------------------------------------------------------------------------------------------
# 'dd' (default): double-double evaluation rounded once -- the correctly rounded value of the reference's FORMULAS for
#   every conditioning;  'float64': the reference's own arithmetic (float64 kernel matrix, Cholesky inverse, products).
# Where K is well conditioned the two agree to rounding.  For cond(K) ~ 1e9 (the reference's reentry tracking
# hyper-parameters, research/gpq/gpq_tracking.py:41-44) the float64 covariance weights are dominated by rounding noise
# (eps * cond^2 = O(1)): 'float64' then reproduces the reference's arithmetic, not its bits -- only assigning tf.wm / Wc /
# Wcc from a reference run does that (tests/test_gpu_weights_envelope.py bounds the difference).
_PRECISION = [__import__('os').environ.get('SSM_BQ_PRECISION', 'dd')]


def set_weight_precision(precision):
    """Package-level switch for every bq_weights evaluation that follows: 'dd' or 'float64'.  Returns the old value."""
    if precision not in ('dd', 'float64'):
        raise ValueError("weight precision must be 'dd' or 'float64'")
    old, _PRECISION[0] = _PRECISION[0], precision
    return old


def get_weight_precision():
    return _PRECISION[0]


class weight_precision(object):
    """with weight_precision('float64'): alg = GaussianProcessKalman(...)   # weights in the reference's arithmetic"""

    def __init__(self, precision):
        self.precision = precision

    def __enter__(self):
        self.old = set_weight_precision(self.precision)
        return self

    def __exit__(self, *exc):
        set_weight_precision(self.old)
        return False


# structure of the quadrature weights
# ------------------------------------------------------------------------------------------------
# Point sets of the form [0 | c I | -c I] (UT, fully-symmetric degree 3) are mapped onto themselves by every coordinate
# reflection x_j -> -x_j, and so are the RBF kernel, the polynomial bases closed under x_j -> -x_j and the integration
# density: in exact arithmetic wm(j+) = wm(j-), Wc and K^-1 commute with every reflection, and Wcc(d, .) vanishes
# except for Wcc(d, d+) = -Wcc(d, d-).  Computed weights have this structure up to rounding only (the reference's float64
# weights for its reentry hyper-parameters: 1e-12 ... 1e-9 where zeros belong, next to entries of 0.22).  With the switch
# on (default, 'dd' precision only) the computed weights are projected onto the invariant subspace -- averaged over each
# class of entries that are equal in exact arithmetic, exact zeros where they belong -- which removes rounding noise,
# never adds any, and lets the forward pass use the compact sums of the reflection-symmetric form (ssm_filter.cuh,
# SSM_TF_BQR: ~45 % fewer multiply-adds per moment transform).  'float64' precision keeps the reference's arithmetic as is.
_SYMMETRY = [__import__('os').environ.get('SSM_BQ_SYMMETRIZE', '1') != '0']


def set_weight_symmetry(on):
    """Package-level switch: project 'dd' weights on axis-symmetric point sets onto their exact reflection structure."""
    old, _SYMMETRY[0] = _SYMMETRY[0], bool(on)
    return old


def get_weight_symmetry():
    return _SYMMETRY[0]


def reflective_axis_set(points):
    """c > 0 when points (D, N) are [0 | c I | -c I] bit for bit (the layout the forward pass calls PTS_AXIS_C), else None."""
    pts = np.asarray(points, dtype=np.float64)
    D, N = pts.shape
    if N != 2 * D + 1 or not pts[0, 1] > 0.0:
        return None
    c = pts[0, 1]
    want = np.hstack([np.zeros((D, 1)), c * np.eye(D), -c * np.eye(D)])
    return float(c) if np.array_equal(pts, want) else None


def symmetrize_reflective(points, w, rtol=1e-6):
    """Project one weight set (dict with wm (N,), Wc (N, N), Wcc (D, N), optionally iK (N, N)) onto the reflection-invariant
    structure of an axis-symmetric point set.  Returns a new dict, or `w` itself when the point set is not of that form
    or the weights are further than rtol (relative to the largest entry of each array) from the structure -- a kernel
    or basis without the symmetry is left alone."""
    if reflective_axis_set(points) is None:
        return w
    D = np.asarray(points).shape[0]
    N = 2 * D + 1
    mirror = np.r_[0, np.arange(1, D + 1) + D, np.arange(1, D + 1)]

    def average(A):
        # (A + P_j A P_j^T) / 2 for every axis j in turn: the reflections commute, so the product of these projectors is
        # the average over the whole group; (x + y) / 2 is the same bits for both members of a pair, and a pair made
        # equal by axis j stays equal through the steps of the other axes
        A = 0.5 * (A + A.T)
        for j in range(D):
            perm = np.arange(N)
            perm[[1 + j, 1 + D + j]] = perm[[1 + D + j, 1 + j]]
            A = 0.5 * (A + A[np.ix_(perm, perm)])
        return A

    out = dict(w)
    wm = np.asarray(w['wm'], dtype=np.float64)
    out['wm'] = 0.5 * (wm + wm[mirror])
    out['Wc'] = average(np.asarray(w['Wc'], dtype=np.float64))
    Wcc = np.asarray(w['Wcc'], dtype=np.float64)
    S = np.zeros_like(Wcc)
    for d in range(D):
        v = 0.5 * (Wcc[d, 1 + d] - Wcc[d, 1 + D + d])
        S[d, 1 + d], S[d, 1 + D + d] = v, -v
    out['Wcc'] = S
    if w.get('iK') is not None:
        out['iK'] = average(np.asarray(w['iK'], dtype=np.float64))
    for k in ('wm', 'Wc', 'Wcc') + (('iK',) if w.get('iK') is not None else ()):
        a, b = np.asarray(w[k], dtype=np.float64), out[k]
        if not np.all(np.isfinite(a)) or np.abs(a - b).max() > rtol * np.abs(a).max():
            return w
    return out


def n_sum_k(n, k):
    """All n-tuples of non-negative integers summing to k, in the reference's column order (utils.py:459-475)."""
    assert k >= 0
    if k == 0:
        return np.zeros((n, 1), dtype=int)
    if k == 1:
        return np.eye(n, dtype=int)
    a = n_sum_k(n, k - 1)
    eye = np.eye(n, dtype=int)
    temp = np.zeros((n, (n * (1 + n) // 2) - 1), dtype=int)
    tind = 0
    for i in range(n - 1):
        for j in range(i, n):
            temp[:, tind] = a[:, i] + eye[:, j]
            tind += 1
    return np.hstack((temp, a[:, n - 1:] + eye[:, -1, None]))


class Model(object):
    """Kernel + point set (bqmod.py:15-106)."""
    _supported_points_ = ['sr', 'ut', 'gh', 'fs']
    _supported_kernels_ = ['rbf', 'rbf-student']

    def __init__(self, dim, kern_par, kern_str, point_str, point_par, estimate_par):
        self.kernel = Model.get_kernel(dim, kern_str, kern_par)
        self.points = Model.get_points(dim, point_str, point_par)
        self.estimate_par = estimate_par
        self.str_pts = point_str
        self.str_pts_par = str(point_par)
        self.dim_in, self.num_pts = self.points.shape
        self.eye_d, self.eye_n = np.eye(self.dim_in), np.eye(self.num_pts)
        self.q, self.Q, self.R, self.iK = None, None, None, None
        self.model_var = None
        self.integral_var = None

    @staticmethod
    def get_points(dim, points, point_par):
        """(bqmod.py:340-382)"""
        points = points.lower()
        if points not in Model._supported_points_:
            raise ValueError('Points {} not supported. Supported points are {}.'.format(points, Model._supported_points_))
        if point_par is None:
            point_par = {}
        if points == 'sr':
            return SphericalRadialTransform.unit_sigma_points(dim)
        elif points == 'ut':
            return UnscentedTransform.unit_sigma_points(dim, **point_par)
        elif points == 'gh':
            return GaussHermiteTransform.unit_sigma_points(dim, **point_par)
        return FullySymmetricStudentTransform.unit_sigma_points(dim, **point_par)

    @staticmethod
    def get_kernel(dim, kernel, par):
        """(bqmod.py:384-423); 'rq' is not on the hot path."""
        kernel = kernel.lower()
        if kernel == 'rbf-student':
            return RBFStudent(dim, par)     # dof is the class default 4.0, as in the reference (bqmod.py:421)
        if kernel != 'rbf':
            raise NotImplementedError("kernel '{}' has no device implementation ('rbf', 'rbf-student')".format(kernel))
        return RBFGauss(dim, par)

    def _weights(self, par, mulind=None):
        par = self.kernel.get_parameters(par)
        if isinstance(self.kernel, RBFStudent):
            return self._weights_mc(par)
        w = dv.bq_weights(par[:1], self.points, mulind, precision=get_weight_precision())
        if int(w['info'][0]) != 0:
            raise np.linalg.LinAlgError('kernel matrix is not positive definite (info = {})'.format(int(w['info'][0])))
        if get_weight_symmetry() and get_weight_precision() == 'dd':
            w1 = symmetrize_reflective(self.points, {k: w[k][0] for k in ('wm', 'Wc', 'Wcc', 'iK')})
            for k in ('wm', 'Wc', 'Wcc', 'iK'):
                w[k] = w1[k][None]
        return w


def _weights_mc(self, par):
    """bq_weights (bqmod.py:495-523) from Monte-Carlo kernel expectations: the inverse kernel matrix comes from the
    weights kernel (double-double), q / Q / R / kbar from ssm_rbf_student_expectations, the products are the
    reference's own numpy expressions."""
    k = self.kernel
    p1 = np.array(par[:1], dtype=np.float64)
    p1[0, 0] = 1.0
    iK = dv.bq_weights(p1, self.points, precision=get_weight_precision())['iK'][0]    # eval_inv_dot(par, x, scaling=False)
    q, Q, R = k.exp_x_kx(par, self.points), k.exp_x_kxkx(par, par, self.points), k.exp_x_xkx(par, self.points)
    w_c = iK.dot(Q).dot(iK)
    if not np.array_equal(w_c, w_c.T):
        w_c = 0.5 * (w_c + w_c.T)
    model_var = k.exp_x_kxx(par) * (1 - np.trace(Q.dot(iK)))
    integral_var = k.exp_xy_kxy(par) - q.T.dot(iK).dot(q)
    self.q, self.Q, self.R = q, Q, R
    return dict(wm=q.dot(iK)[None], Wc=w_c[None], Wcc=R.dot(iK)[None], iK=iK[None], model_var=np.array([model_var]),
                integral_var=np.array([integral_var]), info=np.zeros(1, dtype=np.int32))


Model._weights_mc = _weights_mc


def _nlml_batch(self, log_par, fcn_obs, x_obs, jitter=None):
    """Negative log marginal likelihood and its gradient for a BATCH of kernel log-parameter vectors in one launch
    (ssm_gp_nlml): log_par (n_par, D+1) -> nlml (n_par,), grad (n_par, D+1), info (n_par,).  The batched form of
    neg_log_marginal_likelihood (bqmod.py:537-596 / 1191-1245); not positive-definite kernel matrices give NaN and
    info = 1 instead of an exception."""
    lp = dv._c(np.atleast_2d(log_par))
    x = dv._c(x_obs)
    y = dv._c(np.asarray(fcn_obs, dtype=np.float64).reshape(x.shape[1], -1))
    D, N = x.shape
    E, n_par = y.shape[1], lp.shape[0]
    if lp.shape[1] != D + 1:
        raise ValueError('log_par must have {} columns (log alpha, log l_1..l_D)'.format(D + 1))
    jit = None if jitter is None else dv._c(np.broadcast_to(np.asarray(jitter, dtype=np.float64), (N, N)))
    kw = dict(dtype=torch.float64, device='cuda')
    nlml, grad = torch.empty(n_par, **kw), torch.empty((n_par, D + 1), **kw)
    info = torch.empty(n_par, dtype=torch.int32, device='cuda')
    nu = float(self.nu) if isinstance(self, StudentTProcessModel) else 0.0
    rc = _lib.lib.ssm_gp_nlml(D, N, E, n_par, dv._ptr(lp), dv._ptr(x), dv._ptr(y), dv._ptr(jit) if jit is not None else None, nu,
                              dv._p(nlml), dv._p(grad), dv._p(info), dv._stream())
    _lib.check(rc, 'ssm_gp_nlml')
    return nlml.cpu().numpy(), grad.cpu().numpy(), info.cpu().numpy()


def _neg_log_marginal_likelihood(self, log_par, fcn_obs, x_obs, jitter):
    """-> (nlml, gradient) for one log-parameter vector (bqmod.py:537-596; Student-t process: :1191-1245)."""
    v, g, info = self.nlml_batch(np.asarray(log_par, dtype=np.float64).reshape(1, -1), fcn_obs, x_obs, jitter)
    if info[0] != 0:
        raise np.linalg.LinAlgError('kernel matrix is not positive definite')     # scipy cho_factor, bqmod.py:578
    return float(v[0]), g[0]


def _optimize(self, log_par_0, fcn_obs, x_obs, method='BFGS', **kwargs):
    """Model.optimize (bqmod.py:250-285): scipy.optimize.minimize over the kernel log-parameters with the device
    objective and gradient."""
    from scipy.optimize import minimize
    jitter = 1e-8 * np.eye(np.asarray(x_obs).shape[1])
    return minimize(self.neg_log_marginal_likelihood, np.asarray(log_par_0, dtype=np.float64).reshape(-1),
                    args=(fcn_obs, x_obs, jitter), method=method, jac=True, **kwargs)


def _optimize_multistart(self, log_par_0, fcn_obs, x_obs, method='BFGS', **kwargs):
    """Batched extension: the objective of EVERY start in log_par_0 (n_start, D+1) is evaluated in one launch, the
    starts are ranked, and the best `n_refine` (default 1) are refined with optimize().  Returns the best result and
    the initial objective values."""
    n_refine = int(kwargs.pop('n_refine', 1))
    lp0 = np.atleast_2d(np.asarray(log_par_0, dtype=np.float64))
    jitter = 1e-8 * np.eye(np.asarray(x_obs).shape[1])
    v0, _, info = self.nlml_batch(lp0, fcn_obs, x_obs, jitter)
    v0 = np.where(info == 0, v0, np.inf)
    best = None
    for i in np.argsort(v0)[:n_refine]:
        if not np.isfinite(v0[i]):
            continue
        r = self.optimize(lp0[i], fcn_obs, x_obs, method=method, **kwargs)
        if best is None or r.fun < best.fun:
            best = r
    return best, v0


Model.nlml_batch = _nlml_batch
Model.neg_log_marginal_likelihood = _neg_log_marginal_likelihood
Model.optimize = _optimize
Model.optimize_multistart = _optimize_multistart


class GaussianProcessModel(Model):
    """GP model of the integrand (bqmod.py:426-523)."""

    def __init__(self, dim, kern_par, kern_str, point_str, point_par=None, estimate_par=False):
        super(GaussianProcessModel, self).__init__(dim, kern_par, kern_str, point_str, point_par, estimate_par)

    def bq_weights(self, par, *args):
        """-> (wm, Wc, Wcc, model_var, integral_var) (bqmod.py:495-523)."""
        w = self._weights(par)
        self.iK = w['iK'][0]
        self.model_var = float(w['model_var'][0])
        self.integral_var = float(w['integral_var'][0])
        return w['wm'][0], w['Wc'][0], w['Wcc'][0], self.model_var, self.integral_var

    def exp_model_variance(self, par, *args):
        return float(self._weights(par)['model_var'][0])

    def integral_variance(self, par, *args):
        return float(self._weights(par)['integral_var'][0])


class BayesSardModel(Model):
    """GP model with polynomial prior mean (bqmod.py:599-992)."""

    def __init__(self, dim, kern_par, multi_ind=2, point_str='ut', point_par=None, estimate_par=False):
        super(BayesSardModel, self).__init__(dim, kern_par, 'rbf', point_str, point_par, estimate_par)
        if type(multi_ind) is int:
            self.mulind = np.hstack([n_sum_k(dim, td) for td in range(multi_ind + 1)])
        elif type(multi_ind) is np.ndarray:
            self.mulind = multi_ind
        else:
            raise ValueError('Multi-index error: multi-index has to be either int or ndarray')

    def bq_weights(self, par, multi_ind=None):
        """-> (wm, Wc, Wcc, model_var, integral_var) (bqmod.py:893-992)."""
        if multi_ind is None:
            multi_ind = self.mulind
        if not hasattr(multi_ind, 'shape'):
            # the reference forwards the raw int and fails with AttributeError
            raise AttributeError("'int' object has no attribute 'shape'")
        if multi_ind.shape[1] > self.num_pts:
            raise ValueError('Number of basis functions needs to be lower than or equal to the number of points.'
                             'You supplied {:d} basis functions and {:d} points.'.format(multi_ind.shape[1], self.num_pts))
        w = self._weights(par, multi_ind)
        self.iK = w['iK'][0]
        self.model_var = float(w['model_var'][0])
        self.integral_var = float(w['integral_var'][0])
        return w['wm'][0], w['Wc'][0], w['Wcc'][0], self.model_var, self.integral_var

    def exp_model_variance(self, par, mulind=None):
        return float(self._weights(par, self.mulind if mulind is None else mulind)['model_var'][0])

    def integral_variance(self, par, mulind=None):
        return float(self._weights(par, self.mulind if mulind is None else mulind)['integral_var'][0])


class StudentTProcessModel(GaussianProcessModel):
    """Student's t-process model (bqmod.py:1055-1160); weights are the GP weights, the model variance
    is scaled by the data inside the filter kernel."""

    def __init__(self, dim, kern_par, kern_str, point_str, point_par=None, estimate_par=False, nu=4.0):
        super(StudentTProcessModel, self).__init__(dim, kern_par, kern_str, point_str, point_par, estimate_par)
        nu = 3.0 if nu < 2 else nu
        self.nu = nu

    def exp_model_variance(self, par, *args):
        """scale * gp_emv with scale = (nu - 2 + F iK F') / (nu - 2 + N) (bqmod.py:1132-1160); evaluated
        inside the moment-transform kernel on the hot path, here for a single set of observations."""
        fcn_obs = np.squeeze(np.asarray(args[0], dtype=np.float64))
        scale = (self.nu - 2 + fcn_obs.dot(self.iK).dot(fcn_obs.T)) / (self.nu - 2 + self.num_pts)
        return scale * self.model_var
