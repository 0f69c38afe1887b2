"""State-space models (mirror of ssmtoybox/ssmod.py for the models with a device implementation).

Class names, constructor signatures, attributes and method signatures follow the reference
(TransitionModel ssmod.py:9-244, MeasurementModel :858-1039 and the concrete models cited on each
class).  Function evaluation and simulation run on the GPU (ssm_model_eval, ssm_simulate); random
draws come from Philox streams keyed by (utils.seed, call counter, global trajectory index) instead
of numpy's global MT19937, so simulated data are statistically, not bit-wise,
equivalent to the reference's.
"""
import ctypes as C
from abc import ABCMeta

import numpy as np
import torch

from . import _lib, device as dv
from ._lib import lib
from .utils import GaussRV, StudentRV, next_stream_seed


def _eval(which, model_id, dim_state, si, par, time, x, noise, dim_out):
    """Evaluate a device model function at the columns of x (D, n) (or a single (D,) point)."""
    x = np.asarray(x, dtype=np.float64)
    single = x.ndim == 1
    xs = np.ascontiguousarray(x.reshape(x.shape[0], -1))
    n = xs.shape[1]
    xt = torch.as_tensor(xs, device='cuda')
    nt = None
    if noise is not None:
        nz = np.asarray(noise, dtype=np.float64)
        nz = np.array(np.broadcast_to(nz.reshape(nz.shape[0], -1), (nz.shape[0], n)), dtype=np.float64, order='C')
        nt = torch.as_tensor(nz, device='cuda')
    out = torch.empty((dim_out, n), dtype=torch.float64, device='cuda')
    p = (C.c_double * 8)(*(list(par) + [0.0] * (8 - len(par))))
    t = float(np.asarray(time).reshape(-1)[0]) if time is not None else 0.0
    rc = lib.ssm_model_eval(which, model_id, dim_state, si[0], si[1], p, t, dv._p(xt), dv._p(nt), dv._p(out), n, n,
                            dv._stream())
    _lib.check(rc, 'ssm_model_eval')
    o = out.cpu().numpy()
    return o[:, 0] if single else o


def _replayed(rv, noise=False):
    """True for random variables the simulation kernel cannot draw itself (anything but GaussRV / StudentRV, e.g.
    GaussianMixtureRV): their own sample() draws the noise and the kernel replays it (injected-noise mode).
    The kernel draws process / measurement noise as F z without a mean (ssm_rng carries the mean of x0 only), so a
    NOISE variable with a non-zero mean is replayed too: its sample() includes the mean like the reference's
    noise_rv.sample() (ssmod.py:193, 1033)."""
    if not isinstance(rv, (GaussRV, StudentRV)):
        return True
    return bool(noise and np.any(np.asarray(rv.mean) != 0.0))


def _draw(rv, size):
    """rv.sample(size) as a contiguous float64 CUDA tensor (dim,) + size."""
    try:
        s = rv.sample(size, device_out=True)
    except TypeError:                      # user-defined RandomVariable with the reference's signature
        s = rv.sample(size)
    if not isinstance(s, torch.Tensor):
        s = torch.as_tensor(np.ascontiguousarray(np.asarray(s, dtype=np.float64)), device='cuda')
    return s.to(dtype=torch.float64).contiguous()


def _moments(rv, dim):
    """(mean, cov) for the descriptor; placeholders for replayed random variables (the descriptor of a simulation
    never reads them)."""
    if _replayed(rv):
        return np.zeros(dim), np.eye(dim)
    return rv.get_stats()[:2]


class TransitionModel(metaclass=ABCMeta):
    """State transition model (ssmod.py:9-244)."""
    dim_in = None
    dim_state = None
    dim_noise = None
    noise_additive = None
    _device_id = None  # SSM_DYN_*

    def __init__(self, init_rv, noise_rv, noise_gain=None):
        self.dim_in = self.dim_state if self.noise_additive else self.dim_state + self.dim_noise
        self.init_rv = init_rv
        self.noise_rv = noise_rv
        self.zero_q = np.zeros(self.dim_noise)
        if noise_gain is None:
            noise_gain = np.eye(self.dim_state, self.dim_noise)
        self.noise_gain = noise_gain

    # -- description used by the lowering ---------------------------------------------------------
    def _par(self):
        return [float(getattr(self, 'dt', 0.0)), 0.0, 0.0, 0.0]

    def _desc(self):
        d = {'dyn_name': type(self).__name__, 'dyn_dt': float(getattr(self, 'dt', 0.0)), 'G': self.noise_gain}
        d['m0'], d['P0'] = _moments(self.init_rv, self.dim_state)
        d['q_mean'], d['q_cov'] = _moments(self.noise_rv, self.dim_noise)
        # degrees of freedom per random variable (0 = Gaussian): the simulators draw multivariate-t samples for
        # StudentRV like init_rv.sample() / noise_rv.sample() of the reference (utils.py:349-382, 668-671)
        if isinstance(self.init_rv, StudentRV):
            d['x0_dof'] = float(self.init_rv.dof)
        if isinstance(self.noise_rv, StudentRV):
            d['q_dof'] = float(self.noise_rv.dof)
        return d

    # -- function evaluation ------------------------------------------------------------------------
    def dyn_fcn(self, x, q, time):
        """Discrete-time dynamics f(x, q, time), evaluated on the device."""
        return _eval(0, self._device_id, self.dim_state, (0, 0), self._par(), time, x, q, self.dim_state)

    def dyn_eval(self, xq, time, dx=False):
        """Dynamics according to noise additivity (ssmod.py:129-166).  Jacobians (dx=True) are used by
        the out-of-scope linearisation transforms only."""
        if dx:
            raise NotImplementedError('Jacobians are not part of the device hot path')
        xq = np.asarray(xq, dtype=np.float64)
        if self.noise_additive:
            assert xq.shape[0] == self.dim_state
            return self.dyn_fcn(xq, self.zero_q, time)
        assert xq.shape[0] == self.dim_state + self.dim_noise      # ssmod.py:158-160
        return self.dyn_fcn(xq[:self.dim_state], xq[-self.dim_noise:], time)

    # -- simulation ---------------------------------------------------------------------------------
    def _sim_low(self, obs=None):
        from .ssinf import lower_models  # local import: ssinf imports this module
        return lower_models(self, obs)

    def simulate_discrete(self, steps, mc_sims=1, device_out=False):
        """x (dim_state, steps, mc_sims) with x[:, 0] ~ init_rv, x[:, k] = f(x[:, k-1], q[:, k-1], k-1)
        (ssmod.py:168-199), one GPU thread per trajectory."""
        low, d = self._sim_low()
        if _replayed(self.init_rv) or _replayed(self.noise_rv, noise=True):
            x0, q = _draw(self.init_rv, mc_sims), _draw(self.noise_rv, (steps, mc_sims))
            x, _ = dv.simulate(low, mc_sims, steps, mode='discrete', x0=x0, q=q, want_y=False)
        else:
            rng = dv.make_rng(d, next_stream_seed())
            x, _ = dv.simulate(low, mc_sims, steps, rng=rng, mode='discrete', want_y=False)
        return x if device_out else x.cpu().numpy()

    def simulate_continuous(self, duration, dt=0.1, mc_sims=1, device_out=False):
        """Euler-Maruyama SDE simulation, returns x[:, 1:] (ssmod.py:201-244)."""
        low, d = self._sim_low()
        steps = int(np.floor(duration / dt))
        if _replayed(self.init_rv) or _replayed(self.noise_rv, noise=True):
            x0, q = _draw(self.init_rv, mc_sims), _draw(self.noise_rv, (steps, mc_sims))
            x, _ = dv.simulate(low, mc_sims, steps, mode='continuous', dt=dt, sub=1, x0=x0, q=q, want_y=False)
        else:
            rng = dv.make_rng(d, next_stream_seed())
            x, _ = dv.simulate(low, mc_sims, steps, rng=rng, mode='continuous', dt=dt, sub=1, want_y=False)
        return x if device_out else x.cpu().numpy()


class UNGMTransition(TransitionModel):
    """Univariate non-stationary growth model, additive noise (ssmod.py:247-275)."""
    dim_state = 1
    dim_noise = 1
    noise_additive = True
    _device_id = 1

    def __init__(self, init_rv, noise_rv):
        super(UNGMTransition, self).__init__(init_rv, noise_rv)


class UNGMNATransition(TransitionModel):
    """UNGM with non-additive process noise, x' = 0.5 x + 25 x / (1 + x^2) + 8 q cos(1.2 k) (ssmod.py:278-306)."""
    dim_state = 1
    dim_noise = 1
    noise_additive = False
    _device_id = 6

    def __init__(self, init_rv, noise_rv):
        super(UNGMNATransition, self).__init__(init_rv, noise_rv)


class Pendulum2DTransition(TransitionModel):
    """Pendulum with unit length and mass (ssmod.py:309-365)."""
    dim_state = 2
    dim_noise = 2
    noise_additive = True
    g = 9.81
    _device_id = 2

    def __init__(self, init_rv, noise_rv, dt=0.01):
        super(Pendulum2DTransition, self).__init__(init_rv, noise_rv)
        self.dt = dt


class ReentryVehicle1DTransition(TransitionModel):
    """Vertically falling reentry body: altitude, velocity, ballistic coefficient (ssmod.py:368-429)."""
    dim_state = 3
    dim_noise = 3
    noise_additive = True
    _device_id = 5

    def __init__(self, init_rv, noise_rv, dt=0.1):
        super(ReentryVehicle1DTransition, self).__init__(init_rv, noise_rv)
        self.dt = dt
        self.Gamma = 1 / 6.096


class ReentryVehicle2DTransition(TransitionModel):
    """Reentry vehicle, 5-D state, 3-D noise entering the last three components (ssmod.py:436-584)."""
    dim_state = 5
    dim_noise = 3
    noise_additive = True
    _device_id = 3

    def __init__(self, init_rv, noise_rv, dt=0.1):
        self.dt = dt
        self.R0 = 6374
        self.H0 = 13.406
        self.Gm0 = 3.9860e5
        self.b0 = -0.59783
        noise_gain = np.vstack((np.zeros((2, self.dim_noise)), np.eye(self.dim_noise)))
        super(ReentryVehicle2DTransition, self).__init__(init_rv, noise_rv, noise_gain)


class CoordinatedTurnTransition(TransitionModel):
    """Coordinated turn with time-varying turn rate (ssmod.py:587-696)."""
    dim_state = 5
    dim_noise = 5
    noise_additive = True
    _device_id = 4

    def __init__(self, init_rv, noise_rv, dt=0.1):
        super(CoordinatedTurnTransition, self).__init__(init_rv, noise_rv)
        self.dt = dt


class ConstantTurnRateSpeed(TransitionModel):
    """Constant turn-rate and speed model, state [x, y, speed, heading, yaw rate], NON-additive 2-D noise
    (ssmod.py:699-781); the filters integrate it over the augmented vector [x; q]."""
    dim_state = 5
    dim_noise = 2
    noise_additive = False
    _device_id = 8

    def __init__(self, init_rv, noise_rv, dt=0.05):
        super(ConstantTurnRateSpeed, self).__init__(init_rv, noise_rv)
        self.dt = dt


class ConstantVelocity(TransitionModel):
    """Constant velocity model, state [x, vx, y, vy], 2-D acceleration noise through the gain
    [[dt^2/2, 0], [dt, 0], [0, dt^2/2], [0, dt]] (ssmod.py:783-855)."""
    dim_state = 4
    dim_noise = 2
    noise_additive = True
    _device_id = 7

    def __init__(self, init_rv, noise_rv, dt=0.1):
        self.dt = dt
        noise_gain = np.array([[self.dt ** 2 / 2, 0], [self.dt, 0], [0, self.dt ** 2 / 2], [0, self.dt]])
        super(ConstantVelocity, self).__init__(init_rv, noise_rv, noise_gain)


class MeasurementModel(metaclass=ABCMeta):
    """Measurement model (ssmod.py:858-1039)."""
    dim_substate = None
    dim_out = None
    dim_noise = None
    noise_additive = None
    _device_id = None  # SSM_OBS_*

    def __init__(self, noise_rv, dim_state, state_index):
        self.noise_rv = noise_rv
        self.zero_r = np.zeros(self.dim_noise)
        self.state_index = state_index
        self.dim_in = dim_state if self.noise_additive else dim_state + self.dim_noise
        self.dim_state = dim_state

    def _par(self):
        rl = np.asarray(getattr(self, 'radar_loc', [0.0, 0.0]), dtype=np.float64).reshape(-1)
        return [float(v) for v in rl] + [0.0] * (8 - rl.size)

    def _si(self):
        si = list(self.state_index) if self.state_index is not None else list(range(self.dim_substate))
        return (int(si[0]), int(si[1]) if len(si) > 1 else 0)

    def _desc(self):
        d = {'obs_name': type(self).__name__, 'r_mean': _moments(self.noise_rv, self.dim_noise)[0], 'r_cov': _moments(self.noise_rv, self.dim_noise)[1],
             'state_index': [] if self.state_index is None else list(self.state_index),
             'radar_loc': np.asarray(getattr(self, 'radar_loc', [0.0, 0.0]), dtype=np.float64)}
        if isinstance(self.noise_rv, StudentRV):
            d['r_dof'] = float(self.noise_rv.dof)
        return d

    def meas_fcn(self, x, r, time):
        """Measurement function h(x, r, time) of the (already index-selected) sub-state x."""
        x = np.asarray(x, dtype=np.float64)
        # the device function reads the sub-state from the full state through state_index
        full = np.zeros((self.dim_state,) + x.shape[1:])
        si = self._si()
        for j in range(self.dim_substate):
            full[si[j] if j < 2 else j] = x[j]
        return _eval(1, self._device_id, self.dim_state, si, self._par(), time, full, r, self.dim_out)

    def meas_eval(self, xr, time, dx=False):
        """Measurement model according to noise additivity (ssmod.py:960-1009)."""
        if dx:
            raise NotImplementedError('Jacobians are not part of the device hot path')
        xr = np.asarray(xr, dtype=np.float64)
        if self.noise_additive:
            return _eval(1, self._device_id, self.dim_state, self._si(), self._par(), time, xr, self.zero_r, self.dim_out)
        assert xr.shape[0] == self.dim_state + self.dim_noise      # ssmod.py:999-1001 (state_index = None)
        return _eval(1, self._device_id, self.dim_state, self._si(), self._par(), time, xr[:self.dim_state],
                     xr[-self.dim_noise:], self.dim_out)

    def simulate_measurements(self, x, device_out=False):
        """y[:, k, i] = h(x[state_index, k, i], r[:, k, i], k+1) (ssmod.py:1011-1039)."""
        from .ssinf import lower_models
        xt = x if isinstance(x, torch.Tensor) else torch.as_tensor(
            np.ascontiguousarray(np.asarray(x, dtype=np.float64)), device='cuda')
        if xt.ndim == 2:
            xt = xt[:, :, None]
        if xt.shape[0] != self.dim_state:
            raise ValueError('state array must have dim_state = {} rows'.format(self.dim_state))
        low, d = lower_models(None, self)
        if _replayed(self.noise_rv, noise=True):
            y = dv.simulate_measurements(low, xt.contiguous(), r=_draw(self.noise_rv, tuple(xt.shape[1:])))
        else:
            rng = dv.make_rng(d, next_stream_seed())
            y = dv.simulate_measurements(low, xt.contiguous(), rng=rng)
        return y if device_out else y.cpu().numpy()


class UNGMMeasurement(MeasurementModel):
    """z = 0.05 x^2 + r (ssmod.py:1042-1064)."""
    dim_substate = 1
    dim_out = 1
    dim_noise = 1
    noise_additive = True
    _device_id = 1

    def __init__(self, noise_rv, dim_state, state_index=None):
        super(UNGMMeasurement, self).__init__(noise_rv, dim_state, state_index)


class UNGMNAMeasurement(MeasurementModel):
    """z = 0.05 r x^2, non-additive measurement noise (ssmod.py:1067-1089)."""
    dim_substate = 1
    dim_out = 1
    dim_noise = 1
    noise_additive = False
    _device_id = 5

    def __init__(self, noise_rv, dim_state, state_index=None):
        super(UNGMNAMeasurement, self).__init__(noise_rv, dim_state, state_index)


class Pendulum2DMeasurement(MeasurementModel):
    """z = sin(alpha) + r (ssmod.py:1090-1118)."""
    dim_substate = 1
    dim_out = 1
    dim_noise = 1
    noise_additive = True
    _device_id = 2

    def __init__(self, noise_rv, dim_state, state_index=None):
        super(Pendulum2DMeasurement, self).__init__(noise_rv, dim_state, state_index)


class RangeMeasurement(MeasurementModel):
    """Range of a vertically moving object from a sensor at (sx, sy) = (30, 30) (ssmod.py:1121-1151)."""
    dim_substate = 1
    dim_out = 1
    dim_noise = 1
    noise_additive = True
    _device_id = 4

    def __init__(self, noise_rv, dim_state, state_index=None):
        super(RangeMeasurement, self).__init__(noise_rv, dim_state, state_index)
        self.sx = 30
        self.sy = 30

    # the sensor position travels in the radar_loc slot of the descriptor (obs_par[0..1])
    radar_loc = property(lambda self: np.array([float(self.sx), float(self.sy)]))


class Radar2DMeasurement(MeasurementModel):
    """Range and bearing from a radar at radar_loc (ssmod.py:1199-1255)."""
    dim_substate = 2
    dim_out = 2
    dim_noise = 2
    noise_additive = True
    _device_id = 3

    def __init__(self, noise_rv, dim_state, state_index=None, radar_loc=None):
        super(Radar2DMeasurement, self).__init__(noise_rv, dim_state, state_index)
        if radar_loc is None:
            radar_loc = np.array([0, 0])
        self.radar_loc = radar_loc


class BearingMeasurement(MeasurementModel):
    """Bearings from the sensors at sensor_pos (n_sensors, 2) to the object (ssmod.py:1155-1198).  The device
    implementation covers 4 sensors (the default and the reference's test set-up, tests/test_ssinf.py:77-79); the
    sensor positions travel in the radar_loc slot of the descriptor (obs_par[0..7])."""
    dim_substate = 2
    dim_out = 4
    dim_noise = 4
    noise_additive = True
    _device_id = 6

    def __init__(self, noise_rv, dim_state, state_index=None, sensor_pos=None):
        super(BearingMeasurement, self).__init__(noise_rv, dim_state, state_index)
        if sensor_pos is None:
            sensor_pos = np.vstack((np.eye(2), -np.eye(2)))
        self.sensor_pos = sensor_pos
        self.dim_out = len(self.sensor_pos)
        self.dim_noise = self.dim_out
        self.zero_r = np.zeros(self.dim_noise)
        if self.dim_out != 4:
            raise NotImplementedError('BearingMeasurement runs on the device with 4 sensors (got {})'.format(self.dim_out))

    radar_loc = property(lambda self: np.asarray(self.sensor_pos, dtype=np.float64).reshape(-1))

