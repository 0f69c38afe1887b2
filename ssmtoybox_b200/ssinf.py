"""Local filters and smoothers (mirror of ssmtoybox/ssinf.py for the hot path).

Class names, constructor signatures, public attributes and the forward_pass / backward_pass / reset
protocol follow the reference (StateSpaceInference ssinf.py:19-212, GaussianInference :215-344,
StudentianInference :555-740 and the concrete filters cited on each class).  The time loop, both moment
transforms, the model functions and the measurement update run inside ONE CUDA kernel launch for all
trajectories (ssm_filter); the RTS smoother is a second launch (ssm_smooth).

Batched extension (new): forward_pass accepts data of shape (dy, N, M) -- the array the reference's
research drivers loop over (research/gpq/icinco_demo.py:115-123) -- and returns (dx, N, M),
(dx, dx, N, M).  numpy in -> numpy out; torch CUDA tensors in -> torch CUDA tensors out (no host
round trip).  Every trajectory of a batch starts from the model's initial moments, like the reference
after reset(); a single-trajectory call without reset() continues from the last posterior like the
reference does (ssinf.py:93, 245).

Failures: a single-trajectory call raises numpy.linalg.LinAlgError / ValueError where the reference
does (mtran.py:139, bqmtran.py:98, ssinf.py:321); a batched call never raises for numerical failures:
`self.status` (M,) holds 0 or (step << 8 | code) and the failed trajectory's outputs are NaN from the
failing step on.
"""
import warnings
from abc import ABCMeta

import numpy as np
import torch

from . import _lib, device as dv
from .bq.bqmtran import GaussianProcessTransform, StudentTProcessTransform, BayesSardTransform
from .mtran import MomentTransform, SphericalRadialTransform, UnscentedTransform, GaussHermiteTransform, \
    FullySymmetricStudentTransform
from .ssmod import TransitionModel, MeasurementModel
from .utils import StudentRV

_DEFAULT_OBS = {1: ('UNGMMeasurement', []), 2: ('Pendulum2DMeasurement', []), 3: ('Radar2DMeasurement', []),
                4: ('Radar2DMeasurement', [0, 2]), 5: ('RangeMeasurement', []), 6: ('UNGMNAMeasurement', []),
                7: ('Radar2DMeasurement', []), 8: ('Radar2DMeasurement', [])}
# (measurement model, state_index[, dim_state]) -> (transition model, dx, dq); the entry with dim_state wins
_DEFAULT_DYN = {('UNGMMeasurement', ()): ('UNGMTransition', 1, 1), ('Pendulum2DMeasurement', ()): ('Pendulum2DTransition', 2, 2),
                ('Radar2DMeasurement', ()): ('ReentryVehicle2DTransition', 5, 3),
                ('Radar2DMeasurement', (0, 1)): ('ReentryVehicle2DTransition', 5, 3),
                ('Radar2DMeasurement', (0, 2)): ('CoordinatedTurnTransition', 5, 5),
                ('Radar2DMeasurement', (), 4): ('ConstantVelocity', 4, 2),
                ('Radar2DMeasurement', (0, 1), 4): ('ConstantVelocity', 4, 2),
                ('Radar2DMeasurement', (0, 2), 4): ('ConstantVelocity', 4, 2),
                ('BearingMeasurement', (0, 2)): ('CoordinatedTurnTransition', 5, 5),
                ('UNGMNAMeasurement', ()): ('UNGMNATransition', 1, 1),
                ('RangeMeasurement', ()): ('ReentryVehicle1DTransition', 3, 3),
                ('RangeMeasurement', (0,)): ('ReentryVehicle1DTransition', 3, 3)}


def _sp_stub(dim, prefix):
    pts = UnscentedTransform.unit_sigma_points(dim)
    wm, wc = UnscentedTransform.weights(dim)
    return {prefix + 'kind': 'sp', prefix + 'points': pts, prefix + 'wm': wm, prefix + 'Wc': np.diag(wc)}


def lower_models(dyn, obs):
    """Lower a (transition, measurement) model pair for the simulators; a missing side is filled with
    the matching default model (it is not evaluated)."""
    d = {}
    if dyn is not None:
        d.update(dyn._desc())
    if obs is not None:
        d.update(obs._desc())
    if obs is None:
        name, si = _DEFAULT_OBS[dyn._device_id]
        dy = 2 if name == 'Radar2DMeasurement' else 1
        d.update({'obs_name': name, 'state_index': si, 'radar_loc': [0.0, 0.0], 'r_cov': np.eye(dy)})
    if dyn is None:
        key = (type(obs).__name__, tuple(obs.state_index) if obs.state_index is not None else ())
        if key + (obs.dim_state,) in _DEFAULT_DYN:
            key = key + (obs.dim_state,)
        if key not in _DEFAULT_DYN:
            raise NotImplementedError('no device implementation for {} with state_index {}'.format(*key[:2]))
        name, dx, dq = _DEFAULT_DYN[key]
        d.update({'dyn_name': name, 'dyn_dt': 0.1, 'G': np.eye(dx, dq), 'm0': np.zeros(dx), 'P0': np.eye(dx),
                  'q_cov': np.eye(dq)})
    dx = np.asarray(d['m0']).shape[0]
    d.update(_sp_stub(dx, 'dyn_'))
    d.update(_sp_stub(dx, 'obs_'))
    # StudentRV models: the simulators draw multivariate-t noise (make_rng reads the per-variable dofs)
    d['sample_student'] = any(k in d for k in ('x0_dof', 'q_dof', 'r_dof'))
    return dv.lower(d), d


class StateSpaceInference(metaclass=ABCMeta):
    """Base class of the local filters / smoothers (ssinf.py:19-212)."""

    def __init__(self, mod_dyn, mod_obs, tf_dyn, tf_obs):
        assert isinstance(mod_dyn, TransitionModel) and isinstance(mod_obs, MeasurementModel)
        self.mod_dyn = mod_dyn
        self.mod_obs = mod_obs
        assert isinstance(tf_dyn, MomentTransform) and isinstance(tf_obs, MomentTransform)
        self.tf_dyn = tf_dyn
        self.tf_obs = tf_obs
        self.flags = {'filtered': False, 'smoothed': False}
        self.x_mean_pr, self.x_cov_pr, = None, None
        self.x_mean_sm, self.x_cov_sm = None, None
        self.xx_cov, self.xy_cov = None, None
        self.pr_mean, self.pr_cov, self.pr_xx_cov = None, None, None
        self.fi_mean, self.fi_cov = None, None
        self.sm_mean, self.sm_cov = None, None
        self.D, self.N = None, None
        self.status = None
        self._fwd, self._sm, self._lazy = None, None, {}   # device arrays of the last passes
        self._carry = None        # per-trajectory (mean, cov) carried to the next call when reset() is skipped

    def get_flag(self, key):
        return self.flags[key]

    def set_flag(self, key, value):
        self.flags[key] = value

    # -- lowering -----------------------------------------------------------------------------------
    def _describe(self):
        """Flat description of the CURRENT state of the filter object (read at forward_pass time, because
        research code mutates transforms and model variances after construction)."""
        d = {}
        d.update(self.mod_dyn._desc())
        d.update(self.mod_obs._desc())
        d.update(self.tf_dyn._tf_dict('dyn_'))
        d.update(self.tf_obs._tf_dict('obs_'))
        d['m0'], d['P0'] = self.x0_mean, self.x0_cov
        return d

    def _forward_device(self, y, store_pred=True):
        dy, N, M = y.shape
        low = dv.lower(self._describe())
        if low.dy != dy:
            raise ValueError('data has {} rows, the measurement model has dim_out = {}'.format(dy, low.dy))
        init_mean = init_cov = None
        if self._carry is not None and self._carry[0].shape[-1] == M:
            init_mean, init_cov = self._carry
        fwd = dv.filter_forward(low, y, store_pred=store_pred, init_mean=init_mean, init_cov=init_cov, want_last=True)
        fwd['init_mean'] = init_mean if init_mean is not None else \
            torch.as_tensor(np.asarray(self.x0_mean, dtype=np.float64), device=y.device)[:, None].expand(-1, M)
        fwd['init_cov'] = init_cov if init_cov is not None else \
            torch.as_tensor(np.asarray(self.x0_cov, dtype=np.float64), device=y.device)[:, :, None].expand(-1, -1, M)
        self._low = low
        return fwd

    @staticmethod
    def _raise_for_status(st):
        code = st & 0xFF
        if code == _lib.FAIL_NONFINITE_GAIN:
            raise ValueError('array must not contain infs or NaNs')  # scipy.linalg.cho_factor, ssinf.py:321
        raise np.linalg.LinAlgError('Matrix is not positive definite (time step {})'.format(st >> 8))

    def forward_pass(self, data):
        """Filtering.  data (dy, N) -> (dx, N), (dx, dx, N); or batched (dy, N, M) -> (dx, N, M),
        (dx, dx, N, M) (ssinf.py:66-118)."""
        is_torch = isinstance(data, torch.Tensor)
        y = data if is_torch else torch.as_tensor(np.ascontiguousarray(np.asarray(data, dtype=np.float64)), device='cuda')
        single = y.ndim == 2
        if single:
            y = y[:, :, None]
        y = y.to(dtype=torch.float64).contiguous()
        if not y.is_cuda:
            y = y.cuda()
        self.D, self.N = y.shape[0], y.shape[1]
        fwd = self._forward_device(y)
        self._fwd = fwd
        self.status = fwd['status']
        # carry the last posterior to the next call (the reference keeps x_mean_fi / x_cov_fi, ssinf.py:93)
        self._carry = (fwd['last_mean'], fwd['last_cov'])
        self._single = single
        if single:
            st = int(fwd['status'][0].item())
            if st != 0:
                self._raise_for_status(st)
        self._sm = None
        self._publish_forward(fwd, single, is_torch)
        self.set_flag('filtered', True)
        return self._out(fwd['fi_mean']), self._out(fwd['fi_cov'])

    def _publish_forward(self, fwd, single, is_torch):
        """Expose the reference's public attributes.  The arrays with N+1 time slots (slot 0 = initial
        moments, ssinf.py:85-96) are assembled lazily on first access: the hot path returns views of the
        kernel's own N-slot output arrays and never pays for the copy."""
        self._lazy = {}
        self._is_torch = is_torch

        def out(t):
            if single:
                t = t[..., 0]
            return t if is_torch else t.cpu().numpy()
        self._out = out
        self.x_mean_fi, self.x_cov_fi = out(fwd['fi_mean'][:, -1]), out(fwd['fi_cov'][:, :, -1])
        self.x_mean_pr, self.x_cov_pr = out(fwd['pr_mean'][:, -1]), out(fwd['pr_cov'][:, :, -1])
        self.xx_cov = out(fwd['pr_xx_cov'][:, :, -1])
        self.x_mean_sm, self.x_cov_sm = self.x_mean_fi, self.x_cov_fi  # ssinf.py:117
        if not is_torch:
            self.status = fwd['status'].cpu().numpy()

    def _with_slot0(self, name, key, is_cov, init_key=None, src=None):
        """(N+1)-slot array of the reference: initial moments followed by the kernel output."""
        if name not in self._lazy:
            src = self._fwd if src is None else src
            init = self._fwd['init_cov' if is_cov else 'init_mean']
            t = torch.cat([init[:, :, None, :] if is_cov else init[:, None, :], src[key]], dim=2 if is_cov else 1)
            self._lazy[name] = self._out(t)
        return self._lazy[name]

    fi_mean = property(lambda self: None if self._fwd is None else self._with_slot0('fi_mean', 'fi_mean', False),
                       lambda self, v: None)
    fi_cov = property(lambda self: None if self._fwd is None else self._with_slot0('fi_cov', 'fi_cov', True),
                      lambda self, v: None)
    pr_mean = property(lambda self: None if self._fwd is None else self._with_slot0('pr_mean', 'pr_mean', False),
                       lambda self, v: None)
    pr_cov = property(lambda self: None if self._fwd is None else self._with_slot0('pr_cov', 'pr_cov', True),
                      lambda self, v: None)
    pr_xx_cov = property(lambda self: None if self._fwd is None else self._with_slot0('pr_xx_cov', 'pr_xx_cov', True),
                         lambda self, v: None)
    def _obs_pred(self, idx):
        """y_mean_pr / y_cov_pr / xy_cov of the LAST time step (ssinf.py:281-294): the measurement transform of the
        last predictive state moments, evaluated on first access (ssm_transform_apply; the fused forward pass keeps
        them in registers).  Gaussian family, additive or non-additive measurement noise."""
        if self._fwd is None or isinstance(self, StudentianInference):
            return None
        if 'obs_pred' not in self._lazy:
            from scipy.linalg import block_diag
            fwd = self._fwd
            m = fwd['pr_mean'][:, -1].cpu().numpy()           # (dx, M)
            P = fwd['pr_cov'][:, :, -1].cpu().numpy()         # (dx, dx, M)
            ok = (fwd['status'] == 0).cpu().numpy()
            dx, M = m.shape
            if not self.mod_obs.noise_additive:               # ssinf.py:282-283
                m = np.vstack((m, np.repeat(np.asarray(self.r_mean, dtype=np.float64).reshape(-1, 1), M, axis=1)))
                P = np.stack([block_diag(P[..., i], self.r_cov) for i in range(M)], axis=-1)
            mm, PP = m.copy(), P.copy()
            mm[:, ~ok] = 0.0                                  # failed trajectories: NaN moments, not transformed
            PP[:, :, ~ok] = np.eye(PP.shape[0])[:, :, None]
            time = np.atleast_1d(float(self.N - 1))           # both transforms of step k get time k - 1 (ssinf.py:104)
            ym, yc, xy = self.tf_obs.apply(self.mod_obs.meas_eval, mm, PP, time)
            if self.mod_obs.noise_additive:
                yc = yc + np.asarray(self.r_cov, dtype=np.float64)[:, :, None]      # ssinf.py:290-291
            xy = xy[:, :dx]                                   # ssinf.py:293
            for a in (ym, yc, xy):
                a[..., ~ok] = np.nan
            outs = []
            for a in (ym, yc, xy):
                t = torch.as_tensor(np.ascontiguousarray(a), device=fwd['fi_mean'].device)
                outs.append(self._out(t))
            self._lazy['obs_pred'] = outs
        return self._lazy['obs_pred'][idx]

    y_mean_pr = property(lambda self: self._obs_pred(0), lambda self, v: None)
    y_cov_pr = property(lambda self: self._obs_pred(1), lambda self, v: None)
    xy_cov = property(lambda self: self._obs_pred(2), lambda self, v: None)
    sm_mean = property(lambda self: None if getattr(self, '_sm', None) is None else
                       self._with_slot0('sm_mean', 'sm_mean', False, src=self._sm), lambda self, v: None)
    sm_cov = property(lambda self: None if getattr(self, '_sm', None) is None else
                      self._with_slot0('sm_cov', 'sm_cov', True, src=self._sm), lambda self, v: None)

    def backward_pass(self):
        """Smoothing over the stored forward pass (ssinf.py:120-147)."""
        assert self.get_flag('filtered')  # require filtered state
        fwd = self._fwd
        sm = dv.smooth_backward(self._low.dx, fwd)
        if self._single:
            st = int(sm['status'][0].item())
            if st != 0:
                self._raise_for_status(st)
        self._sm = sm
        self._lazy.pop('sm_mean', None), self._lazy.pop('sm_cov', None)
        self.status = sm['status'] if self._is_torch else sm['status'].cpu().numpy()
        self.set_flag('smoothed', True)
        return self._out(sm['sm_mean']), self._out(sm['sm_cov'])

    def reset(self):
        """Reset internal variables and flags (ssinf.py:149-158)."""
        self.x_mean_pr, self.x_cov_pr = None, None
        self.x_mean_sm, self.x_cov_sm = None, None
        self.xx_cov, self.xy_cov = None, None
        self.pr_mean, self.pr_cov, self.pr_xx_cov = None, None, None
        self.fi_mean, self.fi_cov = None, None
        self.sm_mean, self.sm_cov = None, None
        self.D, self.N = None, None
        self.flags = {'filtered': False, 'smoothed': False}
        self._fwd, self._sm, self._lazy, self._carry, self.status = None, None, {}, None, None


class GaussianInference(StateSpaceInference):
    """Gaussian filters and smoothers (ssinf.py:215-344)."""

    def __init__(self, mod_dyn, mod_obs, tf_dyn, tf_obs):
        assert isinstance(mod_dyn, TransitionModel) and isinstance(mod_obs, MeasurementModel)
        self.x0_mean, self.x0_cov = mod_dyn.init_rv.get_stats()
        self.q_mean, self.q_cov = mod_dyn.noise_rv.get_stats()
        self.r_mean, self.r_cov = mod_obs.noise_rv.get_stats()
        self.G = mod_dyn.noise_gain
        self.x_mean_fi, self.x_cov_fi = self.x0_mean, self.x0_cov
        super(GaussianInference, self).__init__(mod_dyn, mod_obs, tf_dyn, tf_obs)

    def reset(self):
        self.x_mean_fi, self.x_cov_fi = self.x0_mean, self.x0_cov
        super(GaussianInference, self).reset()


class CubatureKalman(GaussianInference):
    """Cubature Kalman filter and smoother (ssinf.py:360-366)."""

    def __init__(self, dyn, obs):
        tf = SphericalRadialTransform(dyn.dim_in)
        th = SphericalRadialTransform(obs.dim_in)
        super(CubatureKalman, self).__init__(dyn, obs, tf, th)


class UnscentedKalman(GaussianInference):
    """Unscented Kalman filter and smoother (ssinf.py:369-386)."""

    def __init__(self, dyn, obs, kappa=None, alpha=1.0, beta=2.0):
        tf = UnscentedTransform(dyn.dim_in, kappa=kappa, alpha=alpha, beta=beta)
        th = UnscentedTransform(obs.dim_in, kappa=kappa, alpha=alpha, beta=beta)
        super(UnscentedKalman, self).__init__(dyn, obs, tf, th)


class GaussHermiteKalman(GaussianInference):
    """Gauss-Hermite Kalman filter and smoother (ssinf.py:389-402)."""

    def __init__(self, dyn, obs, deg=3):
        tf = GaussHermiteTransform(dyn.dim_in, degree=deg)
        th = GaussHermiteTransform(obs.dim_in, degree=deg)
        super(GaussHermiteKalman, self).__init__(dyn, obs, tf, th)


class GaussianProcessKalman(GaussianInference):
    """Gaussian process quadrature Kalman filter and smoother (ssinf.py:405-451)."""

    def __init__(self, dyn, obs, kern_par_dyn, kern_par_obs, kernel='rbf', points='ut', point_hyp=None):
        t_dyn = GaussianProcessTransform(dyn.dim_in, dyn.dim_state, kern_par_dyn, kernel, points, point_hyp)
        t_obs = GaussianProcessTransform(obs.dim_in, obs.dim_out, kern_par_obs, kernel, points, point_hyp)
        super(GaussianProcessKalman, self).__init__(dyn, obs, t_dyn, t_obs)


class BayesSardKalman(GaussianInference):
    """Bayes-Sard quadrature Kalman filter and smoother (ssinf.py:454-500)."""

    def __init__(self, dyn, obs, kern_par_dyn, kern_par_obs, mulind_dyn=2, mulind_obs=2, points='ut', point_hyp=None):
        t_dyn = BayesSardTransform(dyn.dim_in, dyn.dim_state, kern_par_dyn, mulind_dyn, points, point_hyp)
        t_obs = BayesSardTransform(obs.dim_in, obs.dim_out, kern_par_obs, mulind_obs, points, point_hyp)
        super(BayesSardKalman, self).__init__(dyn, obs, t_dyn, t_obs)


class StudentProcessKalman(GaussianInference):
    """Student's t-process quadrature Kalman filter and smoother (ssinf.py:503-552).  The transforms are
    built with dim_out = 1 like the reference, which makes the TPQ variance term a full matrix (Q6)."""

    def __init__(self, dyn, obs, kern_par_dyn, kern_par_obs, kernel='rbf', points='ut', point_hyp=None, nu=3.0):
        t_dyn = StudentTProcessTransform(dyn.dim_in, 1, kern_par_dyn, kernel, points, point_hyp, nu=nu)
        t_obs = StudentTProcessTransform(obs.dim_in, 1, kern_par_obs, kernel, points, point_hyp, nu=nu)
        super(StudentProcessKalman, self).__init__(dyn, obs, t_dyn, t_obs)


class StudentianInference(StateSpaceInference):
    """Filters assuming jointly Student-t state and measurement (ssinf.py:555-740)."""

    def __init__(self, mod_dyn, mod_obs, tf_dyn, tf_obs, dof=4.0, fixed_dof=True):
        if dof <= 2.0:
            dof = 4.0
            warnings.warn("You supplied invalid DoF (must be > 2). Setting to dof=4.")
        self.x0_mean, self.x0_cov, self.x0_dof = mod_dyn.init_rv.get_stats()
        self.x_mean_fi, self.x_cov_fi, self.dof_fi = self.x0_mean, self.x0_cov, self.x0_dof
        self.q_mean, self.q_cov, self.q_dof = mod_dyn.noise_rv.get_stats()
        self.q_gain = mod_dyn.noise_gain
        self.r_mean, self.r_cov, self.r_dof = mod_obs.noise_rv.get_stats()
        scale = (dof - 2) / dof
        self.x_smat_fi = scale * self.x_cov_fi
        self.q_smat = scale * self.q_cov
        self.r_smat = scale * self.r_cov
        self.x_smat_pr, self.y_smat_pr, self.xy_smat = None, None, None
        self.dof = dof
        self.fixed_dof = fixed_dof
        super(StudentianInference, self).__init__(mod_dyn, mod_obs, tf_dyn, tf_obs)

    def _describe(self):
        d = super(StudentianInference, self)._describe()
        d.update({'dof': float(self.dof), 'fixed_dof': int(self.fixed_dof), 'x0_dof': float(self.dof_fi),
                  'q_dof': float(self.q_dof), 'r_dof': float(self.r_dof)})
        return d

    def forward_pass(self, data):
        out = super(StudentianInference, self).forward_pass(data)
        # the state carried between calls is the filtered SCALE matrix and the grown dof (ssinf.py:733-736)
        self.dof_fi = self.dof_fi + self.N * self.mod_obs.dim_out
        lc = self._fwd['last_cov']
        self.x_smat_fi = lc if self._is_torch else lc.cpu().numpy()
        if self._single:
            self.x_smat_fi = self.x_smat_fi[..., 0]
        return out

    def backward_pass(self):
        """Student smoother has not been developed in the reference (ssinf.py:738-740): the smoothed
        arrays are the filtered ones."""
        assert self.get_flag('filtered')
        self._sm = {'sm_mean': self._fwd['fi_mean'], 'sm_cov': self._fwd['fi_cov']}
        self.set_flag('smoothed', True)
        return self._out(self._fwd['fi_mean']), self._out(self._fwd['fi_cov'])

    def reset(self):
        self.x_mean_fi, self.x_cov_fi, self.dof_fi = self.x0_mean, self.x0_cov, self.x0_dof
        scale = (self.dof - 2) / self.dof
        self.x_smat_fi = scale * self.x_cov_fi
        self.x_smat_pr, self.y_smat_pr, self.xy_smat = None, None, None
        super(StudentianInference, self).reset()


class StudentProcessStudent(StudentianInference):
    """Student's t-process quadrature Student filter (TPQSF) on fully-symmetric points (ssinf.py:778-833).  The
    kernel expectations under the Student density come from the device Monte-Carlo kernel (bq.bqkern.RBFStudent)."""

    def __init__(self, dyn, obs, kern_par_dyn, kern_par_obs, point_par=None, dof=4.0, fixed_dof=True, dof_tp=4.0):
        assert isinstance(dyn.init_rv, StudentRV) and isinstance(dyn.noise_rv, StudentRV)
        q_dof, r_dof = dyn.noise_rv.dof, obs.noise_rv.dof
        if point_par is None:
            point_par = dict()
        point_par_dyn = point_par.copy()
        point_par_obs = point_par.copy()
        point_par_dyn.update({'dof': q_dof})
        point_par_obs.update({'dof': r_dof})
        t_dyn = StudentTProcessTransform(dyn.dim_in, 1, kern_par_dyn, 'rbf-student', 'fs', point_par_dyn, nu=dof_tp)
        t_obs = StudentTProcessTransform(obs.dim_in, 1, kern_par_obs, 'rbf-student', 'fs', point_par_obs, nu=dof_tp)
        super(StudentProcessStudent, self).__init__(dyn, obs, t_dyn, t_obs, dof, fixed_dof)


class FullySymmetricStudent(StudentianInference):
    """Student filter with fully-symmetric rules ("Student-t UKF", ssinf.py:743-775)."""

    def __init__(self, dyn, obs, degree=3, kappa=None, dof=4.0, fixed_dof=True):
        dyn_dof = np.min((dyn.init_rv.dof, dyn.noise_rv.dof))
        obs_dof = np.min((dyn_dof, obs.noise_rv.dof))
        t_dyn = FullySymmetricStudentTransform(dyn.dim_in, degree, kappa, dyn_dof)
        t_obs = FullySymmetricStudentTransform(obs.dim_in, degree, kappa, obs_dof)
        super(FullySymmetricStudent, self).__init__(dyn, obs, t_dyn, t_obs, dof, fixed_dof)


# ------------------------------------------------------------------------------------------------------------------
# marginalised moment-transform parameters (ssinf.py:1034-1273; "purely for experimental purposes" in the reference)
# ------------------------------------------------------------------------------------------------------------------
class _Rendezvous(object):
    """Batches the objective evaluations of many scipy optimisers running in parallel threads: every optimiser blocks in
    request(theta) until ALL optimisers that are still running have asked for a point; the main thread then evaluates
    the whole batch in one device pass and releases them.  The optimiser itself is scipy's own BFGS -- the algorithm,
    line search and finite-difference gradient of the reference (ssinf.py:1271) -- so the search path is the
    reference's; only the objective values come from the device."""

    def __init__(self, n):
        import threading
        self.cv = threading.Condition()
        self.n_running = n
        self.pending = {}
        self.results = {}

    def request(self, tid, theta):
        with self.cv:
            self.pending[tid] = np.array(theta, dtype=np.float64)
            self.cv.notify_all()
            while tid not in self.results:
                self.cv.wait()
            return self.results.pop(tid)

    def finish(self, tid):
        with self.cv:
            self.n_running -= 1
            self.cv.notify_all()

    def serve(self, evaluate):
        """main thread: evaluate(ids, thetas (n, P)) -> values (n,) until every optimiser has finished"""
        while True:
            with self.cv:
                while self.n_running > 0 and len(self.pending) < self.n_running:
                    self.cv.wait()
                if self.n_running == 0:
                    return
                ids = sorted(self.pending)
                thetas = np.stack([self.pending.pop(i) for i in ids])
            vals = evaluate(ids, thetas)
            with self.cv:
                for i, v in zip(ids, vals):
                    self.results[i] = float(v)
                self.cv.notify_all()


class MarginalInference(GaussianInference):
    """Gaussian filter with the moment-transform (kernel) parameters marginalised out (ssinf.py:1034-1273): per time
    step a Laplace approximation of the posterior over the log-parameters (BFGS on the un-normalised negative log
    posterior, its inverse-Hessian estimate as covariance, :1247-1273), then a spherical-radial rule over that posterior
    mixes the conditional state posteriors (:1089-1116).

    Device mapping: every objective evaluation of every trajectory, and the 2 P conditional state posteriors per
    trajectory, are (trajectory, parameter vector) PAIRS evaluated in batches -- quadrature weights per pair
    (ssm_bq_weights, one CTA per parameter vector) and both moment transforms with one weight set per column
    (ssm_transform_apply_batched).  The optimiser is scipy's BFGS, one instance per trajectory, run in lock step
    (_Rendezvous).  Batched extension: forward_pass(data (dy, N, M)).  Additive models; no smoother in the reference.

    Reference quirks kept: the measurement update evaluates both transforms with time = k where the plain time update
    uses k - 1 (:66-118 vs :1089-1116); pr_mean / pr_cov / pr_xx_cov come from a time update with whatever weights the
    previous step's last parameter point left in the transform objects (:104-107)."""

    def __init__(self, dyn, obs, tf_dyn, tf_obs, par_mean=None, par_cov=None):
        super(MarginalInference, self).__init__(dyn, obs, tf_dyn, tf_obs)
        if not (dyn.noise_additive and obs.noise_additive):
            raise NotImplementedError('MarginalInference on the device: additive noise models')
        self.param_dyn_dim = self.mod_dyn.dim_in + 1
        self.param_obs_dim = self.mod_obs.dim_state + 1
        self.param_dim = self.param_dyn_dim + self.param_obs_dim
        self.param_prior_mean = np.zeros(self.param_dim, ) if par_mean is None else np.asarray(par_mean, dtype=np.float64)
        self.param_prior_cov = np.eye(self.param_dim) if par_cov is None else np.asarray(par_cov, dtype=np.float64)
        self.param_mean = self.param_prior_mean
        self.param_cov = self.param_prior_cov
        self.param_jitter = 1e-8 * np.eye(self.param_dim)
        self.param_upts = SphericalRadialTransform.unit_sigma_points(self.param_dim)
        self.param_wts = SphericalRadialTransform.weights(self.param_dim)
        self.param_pts_num = self.param_upts.shape[1]
        self.max_threads = 256          # optimisers in flight (one thread each); larger batches run in groups

    def reset(self):
        super(MarginalInference, self).reset()
        self.param_mean = self.param_prior_mean
        self.param_cov = self.param_prior_cov

    # -- batched pair evaluation --------------------------------------------------------------------------------
    def _pairs(self, theta, mean, cov, time):
        """Time update with per-column parameters (ssinf.py:1137-1168): theta (n, P) log-parameters, mean (dx, n),
        cov (dx, dx, n) device tensors -> x_mean_pr, x_cov_pr, xx_cov, y_mean_pr, y_cov_pr, xy_cov (device), ok (n,)."""
        from .bq import bqmod
        prec = bqmod.get_weight_precision()
        dyn, obs = self.mod_dyn, self.mod_obs
        td, to = np.exp(theta[:, :self.param_dyn_dim]), np.exp(theta[:, self.param_dyn_dim:])
        wd = dv.bq_weights(td, self.tf_dyn.model.points, precision=prec, to_host=False)
        wo = dv.bq_weights(to, self.tf_obs.model.points, precision=prec, to_host=False)
        wo['_dim_out'] = obs.dim_out
        dev = mean.device
        mp, Pp, Pxx, s1 = dv.transform_apply_batched(0, dyn._device_id, dyn.dim_state, (0, 0), dyn._par(), self.tf_dyn.model.points,
                                                     wd, time, mean, cov)
        GQG = torch.as_tensor(np.asarray(self.G.dot(self.q_cov).dot(self.G.T), dtype=np.float64), device=dev)
        Pp = Pp + GQG[:, :, None]                                                          # ssinf.py:278-279
        my, Py, Pxy, s2 = dv.transform_apply_batched(1, obs._device_id, obs.dim_state, obs._si(), obs._par(), self.tf_obs.model.points,
                                                     wo, time, mp, Pp)
        Py = Py + torch.as_tensor(np.asarray(self.r_cov, dtype=np.float64), device=dev)[:, :, None]      # ssinf.py:290-291
        ok = (wd['info'] == 0) & (wo['info'] == 0) & (s1 == 0) & (s2 == 0)
        return mp, Pp, Pxx, my, Py, Pxy, ok, (wd, wo)

    @staticmethod
    def _logpdf(y, mean, cov):
        """log N(y | mean, cov) per column (scipy.stats.multivariate_normal.logpdf, ssinf.py:1185): y, mean (dy, n),
        cov (dy, dy, n) numpy."""
        dy, n = mean.shape
        d = (y - mean).T[:, :, None]                       # (n, dy, 1)
        S = np.moveaxis(cov, -1, 0)                        # (n, dy, dy)
        with np.errstate(all='ignore'):
            try:
                L = np.linalg.cholesky(S)
            except np.linalg.LinAlgError:
                L = np.full_like(S, np.nan)
                for i in range(n):
                    try:
                        L[i] = np.linalg.cholesky(S[i])
                    except np.linalg.LinAlgError:
                        pass
            z = np.linalg.solve(L, d)[:, :, 0] if np.isfinite(L).all() else np.stack(
                [np.linalg.solve(L[i], d[i])[:, 0] if np.isfinite(L[i]).all() else np.full(dy, np.nan) for i in range(n)])
            logdet = 2.0 * np.log(np.diagonal(L, axis1=1, axis2=2)).sum(axis=1)
            return -0.5 * (dy * np.log(2.0 * np.pi) + logdet + (z * z).sum(axis=1))

    def forward_pass(self, data):
        """data (dy, N) or (dy, N, M) -> (dx, N[, M]), (dx, dx, N[, M]) (ssinf.py:66-118 with the measurement update of
        :1089-1116).  Also fills param_mean (P[, M]) / param_cov (P, P[, M]) with the last parameter posterior."""
        import threading
        from scipy.optimize import minimize
        is_torch = isinstance(data, torch.Tensor)
        y_all = data.detach().cpu().numpy() if is_torch else np.asarray(data, dtype=np.float64)
        single = y_all.ndim == 2
        if single:
            y_all = y_all[:, :, None]
        dy, N, M = y_all.shape
        dx, P = self.mod_dyn.dim_state, self.param_dim
        self.D, self.N = dy, N
        dev = torch.device('cuda', torch.cuda.current_device())
        kw = dict(dtype=torch.float64, device=dev)
        carry = self._carry
        m = torch.as_tensor(np.repeat(np.asarray(self.x_mean_fi, dtype=np.float64).reshape(dx, 1), M, axis=1), **kw) \
            if carry is None else carry[0].clone()
        Pc = torch.as_tensor(np.repeat(np.asarray(self.x_cov_fi, dtype=np.float64).reshape(dx, dx, 1), M, axis=2), **kw) \
            if carry is None else carry[1].clone()
        mu = np.repeat(np.asarray(self.param_mean, dtype=np.float64).reshape(P, -1)[:, :1], M, axis=1) if np.ndim(self.param_mean) == 1 \
            else np.array(self.param_mean, dtype=np.float64)
        Pi = np.repeat(np.asarray(self.param_cov, dtype=np.float64).reshape(P, P, -1)[:, :, :1], M, axis=2) if np.ndim(self.param_cov) == 2 \
            else np.array(self.param_cov, dtype=np.float64)
        fi_mean, fi_cov = torch.full((dx, N, M), float('nan'), **kw), torch.full((dx, dx, N, M), float('nan'), **kw)
        pr_mean, pr_cov, pr_xx = torch.full((dx, N, M), float('nan'), **kw), torch.full((dx, dx, N, M), float('nan'), **kw), \
            torch.full((dx, dx, N, M), float('nan'), **kw)
        status = np.zeros(M, dtype=np.int32)
        # weights left in the transform objects (dummy unit parameters before the first step), as log-parameters
        last_theta = np.repeat(np.log(np.concatenate([np.asarray(self.tf_dyn.model.kernel.par, dtype=np.float64).reshape(-1),
                                                      np.asarray(self.tf_obs.model.kernel.par, dtype=np.float64).reshape(-1)]))[None], M, axis=0)
        for k in range(1, N + 1):
            alive = np.nonzero(status == 0)[0]
            if alive.size == 0:
                break
            yk = y_all[:, k - 1, :]
            # plain time update of forward_pass (time k - 1) with the weights of the previous step's last parameter point
            mp, Pp, Pxx, _, _, _, _, _ = self._pairs(last_theta[alive], m[:, alive], Pc[:, :, alive], k - 1)
            pr_mean[:, k - 1, alive], pr_cov[:, :, k - 1, alive], pr_xx[:, :, k - 1, alive] = mp, Pp, Pxx
            # ---- Laplace approximation of the parameter posterior (ssinf.py:1247-1273): one BFGS per trajectory --------
            for g0 in range(0, alive.size, self.max_threads):
                grp = alive[g0:g0 + self.max_threads]
                rv = _Rendezvous(len(grp))
                res = {}
                Pi_inv = {int(i): np.linalg.inv(Pi[:, :, i]) for i in grp}
                ldet = {int(i): np.linalg.slogdet(Pi[:, :, i])[1] for i in grp}

                def evaluate(ids, thetas):
                    idx = np.asarray(ids)
                    _, _, _, my, Py, _, ok, _ = self._pairs(thetas, m[:, idx], Pc[:, :, idx], k)
                    ll = self._logpdf(yk[:, idx], my.cpu().numpy(), Py.cpu().numpy())
                    out = np.empty(len(ids))
                    okh = ok.cpu().numpy()
                    for j, i in enumerate(ids):
                        d = thetas[j] - mu[:, i]
                        lp = -0.5 * (P * np.log(2.0 * np.pi) + ldet[i] + d.dot(Pi_inv[i]).dot(d))   # log N(theta | mu, Pi), :1204-1223
                        out[j] = -ll[j] - lp if okh[j] and np.isfinite(ll[j]) else np.inf
                    return out

                def worker(i):
                    try:
                        res[i] = minimize(lambda th: rv.request(i, th), mu[:, i].copy(), method='BFGS')
                    except Exception as e:      # noqa: BLE001
                        res[i] = e
                    finally:
                        rv.finish(i)
                threads = [threading.Thread(target=worker, args=(int(i),), daemon=True) for i in grp]
                for t in threads:
                    t.start()
                rv.serve(evaluate)
                for t in threads:
                    t.join()
                for i in grp:
                    r = res[int(i)]
                    if isinstance(r, Exception) or not np.all(np.isfinite(r.x)) or not np.all(np.isfinite(r.hess_inv)):
                        status[i] = (k << 8) | _lib.FAIL_CHOL_GAIN
                        continue
                    mu[:, i], Pi[:, :, i] = r.x, r.hess_inv + self.param_jitter
            alive = np.nonzero(status == 0)[0]
            if alive.size == 0:
                break
            # ---- marginalisation over the parameter posterior (ssinf.py:1096-1116) ---------------------------------
            pts = np.empty((alive.size, self.param_pts_num, P))
            for a, i in enumerate(alive):
                try:
                    Lp = np.linalg.cholesky(Pi[:, :, i])
                except np.linalg.LinAlgError:
                    status[i] = (k << 8) | _lib.FAIL_CHOL_GAIN
                    Lp = np.eye(P)
                pts[a] = (mu[:, i][:, None] + Lp.dot(self.param_upts)).T
            J = self.param_pts_num
            idx = np.repeat(alive, J)
            mp, Pp, _, my, Py, Pxy, ok, _ = self._pairs(pts.reshape(-1, P), m[:, idx], Pc[:, :, idx], k)
            # conditional posteriors N(x_k | y_1:k, theta) (:1137-1143): tiny dy x dy solves, on the host
            mpn, Ppn = mp.cpu().numpy(), Pp.cpu().numpy()
            myn, Pyn, Pxyn = my.cpu().numpy(), Py.cpu().numpy(), Pxy.cpu().numpy()
            okn = ok.cpu().numpy().reshape(alive.size, J)
            yy = yk[:, idx]
            S = np.moveaxis(Pyn, -1, 0)
            C_ = np.moveaxis(Pxyn, -1, 0)                          # (n, dy, dx) = Cov(y, x)
            with np.errstate(all='ignore'):
                try:
                    gain = np.swapaxes(np.linalg.solve(S, C_), 1, 2)   # (n, dx, dy) = (S^-1 Pyx)^T
                except np.linalg.LinAlgError:
                    gain = np.full((S.shape[0], dx, dy), np.nan)
                mean_ij = mpn.T + np.einsum('nij,nj->ni', gain, (yy - myn).T)
                cov_ij = np.moveaxis(Ppn, -1, 0) - np.einsum('nij,njk,nlk->nil', gain, S, gain)
            mean_i = np.einsum('ajd,j->ad', mean_ij.reshape(alive.size, J, dx), self.param_wts)
            cov_i = np.einsum('ajde,j->ade', cov_ij.reshape(alive.size, J, dx, dx), self.param_wts)
            bad = ~okn.all(axis=1) | ~np.isfinite(mean_i).all(axis=1) | (status[alive] != 0)
            status[alive[bad]] = np.where(status[alive[bad]] != 0, status[alive[bad]], (k << 8) | _lib.FAIL_CHOL_OBS)
            good = alive[~bad]
            m[:, good] = torch.as_tensor(np.ascontiguousarray(mean_i[~bad].T), **kw)
            Pc[:, :, good] = torch.as_tensor(np.ascontiguousarray(np.moveaxis(cov_i[~bad], 0, -1)), **kw)
            fi_mean[:, k - 1, good], fi_cov[:, :, k - 1, good] = m[:, good], Pc[:, :, good]
            last_theta[alive] = pts[:, -1, :]
        st = torch.as_tensor(status, device=dev)
        fwd = {'fi_mean': fi_mean, 'fi_cov': fi_cov, 'pr_mean': pr_mean, 'pr_cov': pr_cov, 'pr_xx_cov': pr_xx, 'status': st,
               'last_mean': m, 'last_cov': Pc,
               'init_mean': torch.as_tensor(np.asarray(self.x0_mean, dtype=np.float64), **kw)[:, None].expand(-1, M),
               'init_cov': torch.as_tensor(np.asarray(self.x0_cov, dtype=np.float64), **kw)[:, :, None].expand(-1, -1, M)}
        self._fwd, self._sm, self._carry, self._single, self.status = fwd, None, (m, Pc), single, st
        if single and int(status[0]) != 0:
            self._raise_for_status(int(status[0]))
        self._publish_forward(fwd, single, is_torch)
        self.param_mean, self.param_cov = (mu[:, 0], Pi[:, :, 0]) if single else (mu, Pi)
        self.set_flag('filtered', True)
        return self._out(fwd['fi_mean']), self._out(fwd['fi_cov'])

    def backward_pass(self):
        raise NotImplementedError('MarginalInference has no smoother (the reference defines none that marginalises the parameters)')


class MarginalizedGaussianProcessKalman(MarginalInference):
    """GPQ Kalman filter with marginalised kernel parameters (ssinf.py:1276-1296; "for experimental purposes only")."""

    def __init__(self, dyn, obs, kernel='rbf', points='ut', point_hyp=None, par_mean=None, par_cov=None):
        # arbitrary dummy kernel parameters, because transforms wouldn't initialize (ssinf.py:1287-1289)
        kpar_dyn = np.ones((1, dyn.dim_in + 1))
        kpar_obs = np.ones((1, obs.dim_state + 1))
        t_dyn = GaussianProcessTransform(dyn.dim_in, 1, kpar_dyn, kernel, points, point_hyp)
        t_obs = GaussianProcessTransform(obs.dim_state, 1, kpar_obs, kernel, points, point_hyp)
        super(MarginalizedGaussianProcessKalman, self).__init__(dyn, obs, t_dyn, t_obs, par_mean, par_cov)
