"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Run in the build container, where
/root/reference exists:

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

The reference's own tests hold no filter-level golden vectors, so parity is
pinned by outputs of the reference itself on seeded inputs.  Every file records the inputs
(measurements, model parameters, the reference's quadrature points and weights) next to the
outputs (filtered / predictive / smoothed moments, simulated trajectories, BQ weights, scores),
so that both the numpy oracle (oracle/ssm_oracle.py) and the CUDA path can be checked against
them without the reference being present.

Reference entry points exercised (file:line relative to /root/reference/ssmtoybox):
  ssinf.py:66-147   forward_pass / backward_pass
  ssinf.py:254-344  Gaussian time / measurement / smoothing updates
  ssinf.py:634-736  Studentian updates
  mtran.py:105-149  SigmaPointTransform.apply
  bq/bqmtran.py:60-109, 394-415   BQTransform.apply, TPQ covariance
  bq/bqmod.py:495-523, 893-992    GP / BS quadrature weights
  bq/bqkern.py:329-424            RBFGauss kernel and its expectations
  ssmod.py:168-244, 1011-1039     simulators
  utils.py:18-148                 scores
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

warnings.simplefilter('ignore')
ref_shim.install()

import scipy  # noqa: E402
from ssmtoybox import ssinf, ssmod, mtran, utils  # noqa: E402
from ssmtoybox.bq import bqmtran, bqmod, bqkern  # noqa: E402
from ssmtoybox.utils import GaussRV, StudentRV  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
VERSIONS = {'numpy': np.__version__, 'scipy': scipy.__version__,
            'reference': 'jacobnzw/SSMToybox v0.1.1a0 (/root/reference)'}


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
class InjectedRV(utils.RandomVariable):
    """Random variable that replays given samples (pattern of research/tpq/tpq_base.py:13-31)."""

    def __init__(self, base, samples):
        self.base, self.samples = base, list(samples)
        self.dim = base.dim
        for a in ('mean', 'cov', 'scale', 'dof'):
            if hasattr(base, a):
                setattr(self, a, getattr(base, a))

    def sample(self, size):
        return self.samples.pop(0)

    def get_stats(self):
        return self.base.get_stats()


def transform_dict(tf, prefix):
    """Flatten a reference moment transform into arrays."""
    d = {}
    if isinstance(tf, bqmtran.BQTransform):
        kind = {bqmod.GaussianProcessModel: 'gp', bqmod.BayesSardModel: 'bs',
                bqmod.StudentTProcessModel: 'tp'}[type(tf.model)]
        d['kind'] = kind
        d['points'] = tf.model.points
        d['wm'], d['Wc'], d['Wcc'] = tf.wm, tf.Wc, tf.Wcc
        d['model_var'] = np.asarray(tf.model.model_var, dtype=float)
        d['integral_var'] = np.asarray(tf.model.integral_var, dtype=float)
        d['kern_par'] = tf.model.kernel.par
        d['dim_out'] = np.asarray(tf.I_out.shape[0])
        if kind == 'tp':
            d['iK'] = tf.model.iK
            d['nu'] = np.asarray(float(tf.model.nu))
        if kind == 'bs':
            d['mulind'] = np.asarray(tf.model.mulind)
    else:
        d['kind'] = 'sp'
        d['points'] = tf.unit_sp
        d['wm'], d['Wc'] = tf.wm, tf.Wc
    return {prefix + k: np.asarray(v) for k, v in d.items()}


def model_dict(dyn, obs):
    d = {'dyn_name': type(dyn).__name__, 'obs_name': type(obs).__name__}
    d['dyn_dt'] = float(getattr(dyn, 'dt', 0.0))
    d['G'] = dyn.noise_gain
    d['state_index'] = np.asarray(obs.state_index if obs.state_index is not None else [], dtype=np.int64)
    d['radar_loc'] = np.asarray(getattr(obs, 'radar_loc', [0.0, 0.0]), dtype=float)
    if hasattr(obs, 'sensor_pos'):  # BearingMeasurement: sensor positions (4, 2), flattened, in the same slot
        d['radar_loc'] = np.asarray(obs.sensor_pos, dtype=float).reshape(-1)
    if hasattr(obs, 'sx'):  # RangeMeasurement: the sensor position (ssmod.py:1143-1144) travels in the same slot
        d['radar_loc'] = np.array([float(obs.sx), float(obs.sy)])
    st = dyn.init_rv.get_stats()
    d['m0'], d['P0'] = st[0], st[1]
    d['q_mean'], d['q_cov'] = dyn.noise_rv.get_stats()[:2]
    d['r_mean'], d['r_cov'] = obs.noise_rv.get_stats()[:2]
    if len(st) == 3:
        d['x0_dof'] = float(st[2])
        d['q_dof'] = float(dyn.noise_rv.dof)
        d['r_dof'] = float(obs.noise_rv.dof)
    return {k: np.asarray(v) for k, v in d.items()}


def run_filter(alg, y, smooth=True):
    """Run forward (and backward) pass per trajectory exactly like research/gpq/icinco_demo.py:120-124.
    Exceptions are recorded per trajectory (status = failing step k, 1-based; 0 = ok)."""
    dy, N, M = y.shape
    dx = alg.mod_dyn.dim_state
    out = {
        'fi_mean': np.full((dx, N, M), np.nan), 'fi_cov': np.full((dx, dx, N, M), np.nan),
        'pr_mean': np.full((dx, N + 1, M), np.nan), 'pr_cov': np.full((dx, dx, N + 1, M), np.nan),
        'pr_xx_cov': np.full((dx, dx, N + 1, M), np.nan),
        'sm_mean': np.full((dx, N, M), np.nan), 'sm_cov': np.full((dx, dx, N, M), np.nan),
        'status': np.zeros(M, dtype=np.int64),
    }
    exc = []
    for i in range(M):
        try:
            m, P = alg.forward_pass(y[..., i])
            out['fi_mean'][..., i], out['fi_cov'][..., i] = m, P
            out['pr_mean'][..., i], out['pr_cov'][..., i] = alg.pr_mean, alg.pr_cov
            out['pr_xx_cov'][..., i] = alg.pr_xx_cov
            if smooth and not isinstance(alg, ssinf.StudentianInference):
                ms, Ps = alg.backward_pass()
                out['sm_mean'][..., i], out['sm_cov'][..., i] = ms, Ps
            exc.append('')
        except (np.linalg.LinAlgError, ValueError) as e:
            # find the step: first all-zero column of fi_mean after slot 0 (arrays are zero-initialised)
            fi = alg.fi_mean
            k = 1
            while k <= N and np.any(fi[:, k] != 0):
                k += 1
            out['status'][i] = k
            out['fi_mean'][:, :k - 1, i] = fi[:, 1:k]
            out['fi_cov'][:, :, :k - 1, i] = alg.fi_cov[:, :, 1:k]
            exc.append(type(e).__name__ + ': ' + str(e)[:60])
        alg.reset()
    out['exceptions'] = np.asarray(json.dumps(exc))
    return out


def save(name, **arrays):
    arrays['versions'] = np.asarray(json.dumps(VERSIONS))
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('{:32s} {:8.1f} kB'.format(name, os.path.getsize(path) / 1e3))


def filter_case(name, alg, x, y, smooth=True, extra=None):
    d = {'x': x, 'y': y}
    d.update(model_dict(alg.mod_dyn, alg.mod_obs))
    d.update(transform_dict(alg.tf_dyn, 'dyn_'))
    d.update(transform_dict(alg.tf_obs, 'obs_'))
    d['alg_name'] = np.asarray(type(alg).__name__)
    if isinstance(alg, ssinf.StudentianInference):
        d['dof'] = np.asarray(float(alg.dof))
        d['fixed_dof'] = np.asarray(int(alg.fixed_dof))
    d.update(run_filter(alg, y, smooth))
    if extra:
        d.update(extra)
    save(name, **d)
    return d


# ----------------------------------------------------------------------------------------------
# state-space model fixtures
# ----------------------------------------------------------------------------------------------
def ungm(steps, mc):
    x0 = GaussRV(1, cov=np.atleast_2d(5.0))
    q = GaussRV(1, cov=np.atleast_2d(10.0))
    dyn = ssmod.UNGMTransition(x0, q)
    obs = ssmod.UNGMMeasurement(GaussRV(1), 1)
    x = dyn.simulate_discrete(steps, mc_sims=mc)
    y = obs.simulate_measurements(x)
    return dyn, obs, x, y


def pendulum(steps, mc):
    x0 = GaussRV(2, mean=np.array([1.5, 0]), cov=0.01 * np.eye(2))
    dt = 0.01
    q = GaussRV(2, cov=0.01 * np.array([[(dt ** 3) / 3, (dt ** 2) / 2], [(dt ** 2) / 2, dt]]))
    r = GaussRV(1, cov=np.array([[0.1]]))
    dyn = ssmod.Pendulum2DTransition(x0, q, dt=dt)
    obs = ssmod.Pendulum2DMeasurement(r, dyn.dim_state)
    x = dyn.simulate_discrete(steps, mc_sims=mc)
    y = obs.simulate_measurements(x)
    return dyn, obs, x, y


def reentry(steps, mc):
    """research/bsq/bsq_tracking.py:230-261 (truth by Euler-Maruyama at dt=0.05, sub-sampled ::2)."""
    tau, disc_tau = 0.05, 0.1
    m0 = np.array([6500, 350, -1.8, -6.8, 0.7])
    sysm = ssmod.ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0])),
                                            GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 0])))
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([6374, 0.0]))
    x = sysm.simulate_continuous(duration=steps * disc_tau, dt=tau, mc_sims=mc)
    y = obs.simulate_measurements(x)
    x, y = x[:, ::2, ...], y[:, ::2, ...]
    m0 = np.array([6500, 350, -1.1, -6.1, 0.7])
    dyn = ssmod.ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 1])),
                                           GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 1e-6])), dt=disc_tau)
    return dyn, obs, x[:, :steps], y[:, :steps]


def coordinated_turn(steps, mc, student=False, dt=0.1):
    """Dynamics of tests/test_ssinf.py:65-78 with the radar of research/tpq/synthetic.py:2126."""
    m0 = np.array([1000, 300, 1000, 0, np.deg2rad(-3.0)])
    P0 = np.diag([100, 10, 100, 10, 0.1])
    rho_1, rho_2 = 0.1, 1.75e-4
    A = np.array([[dt ** 3 / 3, dt ** 2 / 2], [dt ** 2 / 2, dt]])
    Q = np.zeros((5, 5))
    Q[:2, :2], Q[2:4, 2:4], Q[4, 4] = rho_1 * A, rho_1 * A, rho_2 * dt
    R = np.diag([100, 10e-6])
    # data always come from the Gaussian model (simulate_discrete cannot sample StudentRV with a
    # tuple size, utils.py:380-382), heavy tails are injected as outliers below
    dyn_g = ssmod.CoordinatedTurnTransition(GaussRV(5, m0, P0), GaussRV(5, cov=Q), dt=dt)
    obs_g = ssmod.Radar2DMeasurement(GaussRV(2, cov=R), 5, state_index=[0, 2])
    x = dyn_g.simulate_discrete(steps, mc_sims=mc)
    y = obs_g.simulate_measurements(x)
    # 5 % outliers with 50x covariance (research/tpq/synthetic.py:2126-2137)
    out = np.random.rand(steps, mc) < 0.05
    y = y + out[None] * (np.sqrt(49.0) * np.sqrt(np.diag(R))[:, None, None] * np.random.randn(2, steps, mc))
    if not student:
        return dyn_g, obs_g, x, y
    nu = 6.0
    sc = (nu - 2) / nu
    dyn_s = ssmod.CoordinatedTurnTransition(StudentRV(5, m0, sc * P0, nu), StudentRV(5, scale=sc * Q, dof=nu), dt=dt)
    obs_s = ssmod.Radar2DMeasurement(StudentRV(2, scale=sc * R, dof=nu), 5, state_index=[0, 2])
    return dyn_s, obs_s, x, y


def reentry1d(steps, mc):
    """research/gpq/gpq_tracking.py:135-166: vertically falling body + range sensor, truth by Euler-Maruyama at the
    filter's own step, mis-specified ballistic coefficient in the filter model, zero process noise."""
    tau = 0.1
    P0 = np.diag([0.0929, 1.4865, 1e-4])
    sysm = ssmod.ReentryVehicle1DTransition(GaussRV(3, np.array([90, 6, 1.5]), P0), GaussRV(3, cov=np.zeros((3, 3))), dt=tau)
    x = sysm.simulate_continuous(steps * tau, mc_sims=mc)
    obs = ssmod.RangeMeasurement(GaussRV(1, cov=np.array([[0.03048 ** 2]])), 3)
    y = obs.simulate_measurements(x)
    dyn = ssmod.ReentryVehicle1DTransition(GaussRV(3, np.array([90, 6, 1.7]), P0), GaussRV(3, cov=np.zeros((3, 3))), dt=tau)
    return dyn, obs, x[:, :steps], y[:, :steps]


MUL_UT = lambda d: np.hstack((np.zeros((d, 1)), np.eye(d), 2 * np.eye(d))).astype(int)  # noqa: E731


# ----------------------------------------------------------------------------------------------
# golden sets
# ----------------------------------------------------------------------------------------------
def gen_filters():
    # C1: UNGM, UKF, 500 steps (research/gpq/icinco_demo.py:81-125)
    np.random.seed(42)
    dyn, obs, x, y = ungm(500, 6)
    filter_case('c1_ungm_ukf', ssinf.UnscentedKalman(dyn, obs), x, y)
    filter_case('c1_ungm_ckf', ssinf.CubatureKalman(dyn, obs), x[..., :2], y[..., :2])
    filter_case('c1_ungm_ghkf5', ssinf.GaussHermiteKalman(dyn, obs, deg=5), x[..., :2], y[..., :2])
    kp = np.array([[1.0, 3.0]])
    filter_case('c1_ungm_gpq_ut', ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='ut'), x[..., :3], y[..., :3])
    kp = np.array([[1.0, 0.1]])
    filter_case('c1_ungm_gpq_gh10', ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='gh',
                                                                point_hyp={'degree': 10}), x[..., :2], y[..., :2])
    kp = np.array([[1.0, 3.0]])
    filter_case('c1_ungm_tpq_ut', ssinf.StudentProcessKalman(dyn, obs, kp, kp), x[..., :3], y[..., :3])
    filter_case('c1_ungm_bsq_ut', ssinf.BayesSardKalman(dyn, obs, kp, kp, MUL_UT(1), MUL_UT(1)), x[..., :3], y[..., :3])

    # C2: UNGM GPQ length-scale sweep (research/gpq/icinco_demo.py:166-196), with reset() between
    # trajectories (the script forgets it; the batched API mirrors the reset behaviour)
    for iel, el in enumerate([1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1, 3, 1e1, 3e1, 1e2]):
        kp = np.array([[1.0, el]])
        alg = ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='ut', point_hyp={'kappa': 0.0})
        filter_case('c2_ungm_gpq_el{:02d}'.format(iel), alg, x[..., :2], y[..., :2], smooth=False)

    # C5a: pendulum (tests/test_ssinf.py:42-51)
    np.random.seed(7)
    dyn, obs, x, y = pendulum(300, 3)
    filter_case('c5_pend_ukf', ssinf.UnscentedKalman(dyn, obs), x, y)
    kp = np.array([[1.0, 1.0, 1.0]])
    filter_case('c5_pend_gpq', ssinf.GaussianProcessKalman(dyn, obs, kp, kp), x, y)
    filter_case('c5_pend_tpq', ssinf.StudentProcessKalman(dyn, obs, kp, kp), x, y)
    filter_case('c5_pend_bsq', ssinf.BayesSardKalman(dyn, obs, kp, kp, MUL_UT(2), MUL_UT(2)), x, y)
    filter_case('c5_pend_ghkf3', ssinf.GaussHermiteKalman(dyn, obs, deg=3), x[..., :2], y[..., :2])

    # C3: reentry GPQ + RTS (research/gpq/gpq_tracking.py:41-44, research/bsq/bsq_tracking.py:230-261)
    np.random.seed(0)
    dyn, obs, x, y = reentry(500, 3)
    hdyn = np.array([[1.0, 25, 25, 25, 25, 25]])
    hobs = np.array([[1.0, 25, 25, 1e4, 1e4, 1e4]])
    filter_case('c3_reentry_gpq', ssinf.GaussianProcessKalman(dyn, obs, hdyn, hobs, kernel='rbf', points='ut'), x, y)
    filter_case('c3_reentry_ukf', ssinf.UnscentedKalman(dyn, obs), x[..., :2], y[..., :2])
    # 8 shorter trajectories for the score tests (the per-step MSE matrix needs M > dx to be PD)
    np.random.seed(1)
    dyn8, obs8, x8, y8 = reentry(120, 8)
    filter_case('c3s_reentry_gpq', ssinf.GaussianProcessKalman(dyn8, obs8, hdyn, hobs, kernel='rbf', points='ut'), x8, y8)
    filter_case('c3_reentry_ukf_b0', ssinf.UnscentedKalman(dyn, obs, beta=0.0), x[..., :1], y[..., :1])
    filter_case('c3_reentry_ckf', ssinf.CubatureKalman(dyn, obs), x[..., :1], y[..., :1])
    # BSQ with model variance assigned from outside (research/bsq/bsq_tracking.py:266-281)
    par_dyn = np.array([[1.0, 1, 1, 1, 1, 1]])
    par_obs = np.array([[1.0, 0.9, 0.9, 1e4, 1e4, 1e4]])
    alg = ssinf.BayesSardKalman(dyn, obs, par_dyn, par_obs, MUL_UT(5), MUL_UT(5), points='ut')
    alg.tf_dyn.model.model_var = 2e-6 * np.eye(5)
    alg.tf_obs.model.model_var = 0 * np.eye(2)
    filter_case('c3_reentry_bsq', alg, x[..., :2], y[..., :2])
    # failure semantics: GPQ with unit length-scales is not PD on this model
    one = np.array([[1.0, 1, 1, 1, 1, 1]])
    filter_case('c3_reentry_gpq_fail', ssinf.GaussianProcessKalman(dyn, obs, one, one), x[:, :60, :2], y[:, :60, :2])

    # C4: coordinated turn + radar, heavy tails: TPQKF (Gaussian SSM) vs Student-t UKF (Student SSM)
    np.random.seed(3)
    dyn, obs, x, y = coordinated_turn(200, 4)
    par_dyn = np.array([[1.0, 1, 1, 1, 1, 1]])
    par_obs = np.array([[1.0, 1, 1e2, 1, 1e2, 1e2]])
    filter_case('c4_ct_tpq', ssinf.StudentProcessKalman(dyn, obs, par_dyn, par_obs), x, y)
    filter_case('c4_ct_gpq', ssinf.GaussianProcessKalman(dyn, obs, par_dyn, par_obs), x[..., :2], y[..., :2])
    filter_case('c4_ct_ukf', ssinf.UnscentedKalman(dyn, obs), x[..., :2], y[..., :2])
    filter_case('c4_ct_bsq', ssinf.BayesSardKalman(dyn, obs, np.ones((1, 6)), np.ones((1, 6)), MUL_UT(5), MUL_UT(5)),
                x[..., :2], y[..., :2])
    np.random.seed(3)
    dyn_s, obs_s, x, y = coordinated_turn(200, 4, student=True)
    filter_case('c4_ct_fsstudent', ssinf.FullySymmetricStudent(dyn_s, obs_s, kappa=None, dof=6.0), x, y, smooth=False)
    filter_case('c4_ct_fsstudent_incdof', ssinf.FullySymmetricStudent(dyn_s, obs_s, dof=6.0, fixed_dof=False),
                x[..., :2], y[..., :2], smooth=False)
    filter_case('c4_ct_fsstudent_deg5', ssinf.FullySymmetricStudent(dyn_s, obs_s, degree=5, dof=6.0),
                x[..., :1], y[..., :1], smooth=False)


def gen_reentry1d():
    """C6: ReentryVehicle1DTransition + RangeMeasurement, the second tracking experiment of the GPQ paper
    (research/gpq/gpq_tracking.py:114-176): GPQKF with its kernel parameters and the UKF, + RTS smoother."""
    np.random.seed(2)
    dyn, obs, x, y = reentry1d(300, 4)
    kd, ko = np.array([[0.5, 10, 10, 10]]), np.array([[0.5, 15, 20, 20]])
    filter_case('c6_reentry1d_gpq', ssinf.GaussianProcessKalman(dyn, obs, kd, ko, kernel='rbf', points='ut'), x, y)
    filter_case('c6_reentry1d_ukf', ssinf.UnscentedKalman(dyn, obs), x[..., :2], y[..., :2])
    # simulators with injected noise (ssmod.py:168-244, 1011-1039), process noise switched on to exercise its path
    rng = np.random.RandomState(12)
    steps, mc, dt = 40, 3, 0.1
    dyn = ssmod.ReentryVehicle1DTransition(GaussRV(3, np.array([90, 6, 1.5]), np.diag([0.0929, 1.4865, 1e-4])),
                                           GaussRV(3, cov=np.diag([1e-4, 1e-4, 1e-6])), dt=dt)
    x0 = dyn.init_rv.mean[:, None] + rng.randn(3, mc) * 0.1
    q = rng.randn(3, steps, mc) * np.sqrt(np.diag(dyn.noise_rv.cov))[:, None, None]
    qc = rng.randn(3, steps + 1, mc) * np.sqrt(np.diag(dyn.noise_rv.cov))[:, None, None]
    r = rng.randn(1, steps, mc) * 0.03048
    dyn.init_rv = InjectedRV(dyn.init_rv, [x0, x0])
    dyn.noise_rv = InjectedRV(dyn.noise_rv, [q, qc])
    obs.noise_rv = InjectedRV(obs.noise_rv, [r])
    xd = dyn.simulate_discrete(steps, mc_sims=mc)
    xc = dyn.simulate_continuous(duration=steps * dt, dt=dt, mc_sims=mc)
    yd = obs.simulate_measurements(xd)
    d = {'x0': x0, 'q': q, 'qc': qc, 'r': r, 'x': xd, 'xc': xc, 'y': yd, 'dtc': np.asarray(dt)}
    d.update(model_dict(dyn, obs))
    save('simulation_reentry1d', **d)


def gen_ungmna():
    """C7: UNGM with NON-additive process and measurement noise (models of tests/test_ssinf.py:32-40): the filters
    integrate over the augmented vector [x; noise] (ssinf.py:271-272, 282-283).  UKF, CKF, GH, GPQ + RTS smoother.
    The initial mean is 1 instead of the test fixture's 0: with a zero mean every symmetric rule gives a measurement
    covariance of EXACTLY 0 in exact arithmetic (z = 0.05 r x^2 at x = 0), and the reference only gets past its
    Cholesky on the 1e-16 residue that OpenBLAS' dot leaves in the predicted mean -- an artefact no other summation
    order reproduces (that fixture is kept as ungmna_zero_mean_ukf and checked for what it is)."""
    np.random.seed(9)
    dyn = ssmod.UNGMNATransition(GaussRV(1, mean=np.array([1.0])), GaussRV(1, cov=np.array([[10.0]])))
    obs = ssmod.UNGMNAMeasurement(GaussRV(1), 1)
    x = dyn.simulate_discrete(60, mc_sims=6)
    y = obs.simulate_measurements(x)
    dyn0 = ssmod.UNGMNATransition(GaussRV(1), GaussRV(1, cov=np.array([[10.0]])))
    filter_case('ungmna_zero_mean_ukf', ssinf.UnscentedKalman(dyn0, obs), x[:, :20, :2], y[:, :20, :2])

    def case(name, alg, n):
        # the noise means enter the augmented mean (ssinf.py:271, 282)
        filter_case(name, alg, x[..., :n], y[..., :n])
    case('c7_ungmna_ukf', ssinf.UnscentedKalman(dyn, obs), 6)
    case('c7_ungmna_ckf', ssinf.CubatureKalman(dyn, obs), 2)
    case('c7_ungmna_ghkf', ssinf.GaussHermiteKalman(dyn, obs, deg=4), 2)
    kp = np.array([[1.0, 3.0, 3.0]])
    case('c7_ungmna_gpq', ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='ut'), 3)
    # simulators with injected noise
    rng = np.random.RandomState(13)
    steps, mc = 40, 3
    x0, q, r = rng.randn(1, mc), rng.randn(1, steps, mc) * np.sqrt(10.0), rng.randn(1, steps, mc)
    dyn.init_rv, dyn.noise_rv, obs.noise_rv = InjectedRV(dyn.init_rv, [x0]), InjectedRV(dyn.noise_rv, [q]), InjectedRV(obs.noise_rv, [r])
    xs = dyn.simulate_discrete(steps, mc_sims=mc)
    d = {'x0': x0, 'q': q, 'r': r, 'x': xs, 'y': obs.simulate_measurements(xs)}
    d.update(model_dict(dyn, obs))
    save('simulation_ungmna', **d)


def gen_student_bq():
    """Student filters with BQ transforms whose kernel expectations are Monte-Carlo integrals under a Student density
    (RBFStudent, bq/bqkern.py:457-536): StudentProcessStudent (TPQSF, ssinf.py:778-833) and the GPQ Student filter of
    research/tpq/tpq_base.py:41-91, on the heavy-tailed coordinated-turn model.  The Monte-Carlo weights (numpy MT19937,
    2e6 samples) are stored: the device filter is checked with them assigned, its own Monte-Carlo weights statistically."""
    np.random.seed(3)
    dyn_s, obs_s, x, y = coordinated_turn(60, 3, student=True)
    par_dyn = np.array([[1.0, 1, 1, 1, 1, 1]])
    par_obs = np.array([[1.0, 1, 1e2, 1, 1e2, 1e2]])
    np.random.seed(21)
    alg = ssinf.StudentProcessStudent(dyn_s, obs_s, par_dyn, par_obs, dof=6.0)
    extra = {'kern_par_dyn': par_dyn, 'kern_par_obs': par_obs,
             'dyn_q': alg.tf_dyn.model.q, 'dyn_Q': alg.tf_dyn.model.Q, 'obs_q': alg.tf_obs.model.q, 'obs_Q': alg.tf_obs.model.Q}
    filter_case('c4_ct_fsstudent_tpq', alg, x, y, smooth=False, extra=extra)
    # GPQStudent of tpq_base.py: GaussianProcessTransform(dim_in, kern_par, 'rbf-student', 'fs', {'dof': q_dof})
    t_dyn = bqmtran.GaussianProcessTransform(5, 5, par_dyn, 'rbf-student', 'fs', {'dof': dyn_s.noise_rv.dof})
    t_obs = bqmtran.GaussianProcessTransform(5, 2, par_obs, 'rbf-student', 'fs', {'dof': obs_s.noise_rv.dof})
    alg = ssinf.StudentianInference(dyn_s, obs_s, t_dyn, t_obs, 6.0, True)
    filter_case('c4_ct_fsstudent_gpq', alg, x[..., :2], y[..., :2], smooth=False)


def gen_more_models():
    """C8-C10: the remaining models of ssmod.py on the set-ups of the reference's own tests (tests/test_ssinf.py:66-93,
    227-244): ConstantVelocity + radar (Gaussian and Student noise), CoordinatedTurn + 4 bearing sensors,
    ConstantTurnRateSpeed (non-additive noise) + radar.  Filters, smoothers and the simulators with injected noise."""
    rng = np.random.RandomState(31)

    def sims(tag, dyn, obs, steps=40, mc=3, x0_sd=0.1):
        dx, dq, dr = dyn.dim_state, dyn.dim_noise, obs.dim_noise
        qc, rc = dyn.noise_rv.get_stats()[1], obs.noise_rv.get_stats()[1]
        x0 = dyn.init_rv.get_stats()[0][:, None] + rng.randn(dx, mc) * x0_sd
        q = rng.randn(dq, steps, mc) * np.sqrt(np.diag(qc))[:, None, None]
        r = rng.randn(dr, steps, mc) * np.sqrt(np.diag(rc))[:, None, None]
        irv, qrv, rrv = dyn.init_rv, dyn.noise_rv, obs.noise_rv
        dyn.init_rv, dyn.noise_rv, obs.noise_rv = InjectedRV(irv, [x0]), InjectedRV(qrv, [q]), InjectedRV(rrv, [r])
        x = dyn.simulate_discrete(steps, mc_sims=mc)
        d = {'x0': x0, 'q': q, 'r': r, 'x': x, 'y': obs.simulate_measurements(x)}
        d.update(model_dict(dyn, obs))
        dyn.init_rv, dyn.noise_rv, obs.noise_rv = irv, qrv, rrv
        save('simulation_' + tag, **d)

    # C8: constant velocity + radar on the leading two state components (state_index None), tests/test_ssinf.py:227-244
    m0, P0 = np.array([10175, 295, 980, -35.0]), np.diag([10000, 100, 10000, 100.0])
    Q, R = np.diag([50, 5.0]), np.diag([50, 0.4e-6])
    np.random.seed(41)
    dyn = ssmod.ConstantVelocity(GaussRV(4, m0, P0), GaussRV(2, cov=Q), dt=0.5)
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=R), 4)
    x = dyn.simulate_discrete(100, mc_sims=3)
    y = obs.simulate_measurements(x)
    filter_case('c8_cv_ukf', ssinf.UnscentedKalman(dyn, obs), x, y)
    filter_case('c8_cv_ckf', ssinf.CubatureKalman(dyn, obs), x[..., :1], y[..., :1])
    kp = np.array([[1.0, 3, 3, 3, 3]])
    filter_case('c8_cv_gpq', ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='ut'), x[..., :2], y[..., :2])
    obs02 = ssmod.Radar2DMeasurement(GaussRV(2, cov=R), 4, state_index=[0, 2])
    y02 = obs02.simulate_measurements(x)
    filter_case('c8_cv02_ukf', ssinf.UnscentedKalman(dyn, obs02), x[..., :2], y02[..., :2])
    sims('cv', dyn, obs)
    dyn_s = ssmod.ConstantVelocity(StudentRV(4, m0, P0, 1000.0), StudentRV(2, scale=Q, dof=1000.0), dt=0.5)
    obs_s = ssmod.Radar2DMeasurement(StudentRV(2, scale=R, dof=4.0), 4)
    filter_case('c8_cv_fsstudent', ssinf.FullySymmetricStudent(dyn_s, obs_s), x[..., :2], y[..., :2], smooth=False)

    # C9: coordinated turn + bearings from 4 sensors, tests/test_ssinf.py:66-82
    np.random.seed(42)
    dyn, _, x, _ = coordinated_turn(100, 3)
    sen = np.vstack((1000 * np.eye(2), -1000 * np.eye(2))).astype(float)
    obs = ssmod.BearingMeasurement(GaussRV(4, cov=10e-3 * np.eye(4)), 5, state_index=[0, 2], sensor_pos=sen)
    y = obs.simulate_measurements(x)
    filter_case('c9_ctb_ukf', ssinf.UnscentedKalman(dyn, obs), x, y)
    filter_case('c9_ctb_ckf', ssinf.CubatureKalman(dyn, obs), x[..., :1], y[..., :1])
    kp = np.array([[1.0, 3, 3, 3, 3, 3]])
    filter_case('c9_ctb_gpq', ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='ut'), x[..., :2], y[..., :2])
    sims('ctb', dyn, obs)

    # C10: constant turn rate and speed (non-additive noise) + radar, tests/test_ssinf.py:84-93.  The fixture's zero
    # initial mean puts the central sigma point on the x[4] == 0 branch and the object on top of the radar; a moving
    # object away from the origin is the main case, the fixture is kept as c10_ctrs_fixture_ukf.
    np.random.seed(43)
    q, r = GaussRV(2, cov=np.diag([0.1, 0.1 * np.pi])), GaussRV(2, cov=np.diag([0.3, 0.03]))
    dyn = ssmod.ConstantTurnRateSpeed(GaussRV(5, np.array([10.0, 20, 5, 0.3, 0.1]), 0.1 * np.eye(5)), q)
    obs = ssmod.Radar2DMeasurement(r, 5)
    x = dyn.simulate_discrete(100, mc_sims=3)
    y = obs.simulate_measurements(x)
    filter_case('c10_ctrs_ukf', ssinf.UnscentedKalman(dyn, obs), x, y)
    filter_case('c10_ctrs_ckf', ssinf.CubatureKalman(dyn, obs), x[..., :1], y[..., :1])
    kpd, kpo = np.array([[1.0, 3, 3, 3, 3, 3, 3, 3]]), np.array([[1.0, 3, 3, 3, 3, 3]])
    filter_case('c10_ctrs_gpq', ssinf.GaussianProcessKalman(dyn, obs, kpd, kpo, points='ut'), x[..., :2], y[..., :2])
    sims('ctrs', dyn, obs)
    np.random.seed(44)
    dyn0 = ssmod.ConstantTurnRateSpeed(GaussRV(5, cov=0.1 * np.eye(5)), q)
    x = dyn0.simulate_discrete(100, mc_sims=2)
    y = obs.simulate_measurements(x)
    filter_case('c10_ctrs_fixture_ukf', ssinf.UnscentedKalman(dyn0, obs), x, y)


def gen_nlml():
    """Hyper-parameter fitting: neg_log_marginal_likelihood of the GP and the Student-t process model
    (bq/bqmod.py:537-596, 1191-1245) on the set-ups of tests/test_bqmod.py:51-56, 86-96, 160-176, for batches of
    log-parameter vectors, and the optimum Model.optimize (bq/bqmod.py:250-285) finds with BFGS."""
    rs = np.random.RandomState(17)
    d = {}
    cases = []

    def add(tag, model, y, lps):
        x = model.points
        jit = 1e-8 * np.eye(model.num_pts)
        vals, grads = [], []
        for lp in lps:
            f, df = model.neg_log_marginal_likelihood(lp, y, x, jit)
            vals.append(f), grads.append(df)
        d.update({tag + '_x': x, tag + '_y': y, tag + '_log_par': np.array(lps), tag + '_nlml': np.array(vals),
                  tag + '_grad': np.array(grads), tag + '_nu': np.asarray(float(getattr(model, 'nu', 0.0)) if isinstance(model, bqmod.StudentTProcessModel) else 0.0)})
        cases.append(tag)

    f1 = lambda x: 0.05 * x ** 2                                             # tests/test_bqmod.py:18
    for cls, nm in ((bqmod.GaussianProcessModel, 'gp'), (bqmod.StudentTProcessModel, 'tp')):
        m = cls(1, np.array([[1.0, 3.0]]), 'rbf', 'ut', {'alpha': 1.0})
        add(nm + '_1d_ut', m, f1(m.points).T, [np.log([1.0, 3.0])] + [rs.uniform(-1, 1.5, 2) for _ in range(7)])
        m = cls(1, np.array([[1.0, 3.0]]), 'rbf', 'gh', {'degree': 15})
        add(nm + '_1d_gh15', m, f1(m.points).T, [np.log([1.0, 0.5])] + [rs.uniform(-0.5, 1.0, 2) for _ in range(7)])
        # 5-D, 5 outputs: coordinated-turn dynamics at the sigma points (tests/test_bqmod.py:139-157)
        m0 = np.array([1000, 300, 1000, 0, np.deg2rad(-3)])
        dyn = ssmod.CoordinatedTurnTransition(GaussRV(5, m0, np.diag([100, 10, 100, 10, 0.1])), GaussRV(5))
        m = cls(5, np.array([[1.0, 3, 3, 3, 3, 3]]), 'rbf', 'ut', {'alpha': 1.0})
        x = m0[:, None] + m.points
        y = np.apply_along_axis(dyn.dyn_eval, 0, x, None)
        y = y - y.mean(axis=1, keepdims=True)                                # keep the fit well scaled
        add(nm + '_5d_ut', m, y.T, [np.log([1.0] + 5 * [3.0])] + [rs.uniform(0.0, 2.0, 6) for _ in range(7)])
    # optimum of the 1-D GH-15 case (tests/test_bqmod.py:160-176 without the constraint, which BFGS ignores)
    # (the Student-t process objective sends unbounded BFGS into parameters where the kernel matrix is not PD -- the
    # reference raises LinAlgError there --, so that model is fitted with bounds, as tests/test_bqmtran.py:196-207 does)
    bounds = ((np.log(0.5), np.log(2.0)), (np.log(0.2), np.log(5.0)))
    for cls, nm, kw in ((bqmod.GaussianProcessModel, 'gp', dict(method='BFGS')),
                        (bqmod.StudentTProcessModel, 'tp', dict(method='L-BFGS-B', bounds=bounds))):
        m = cls(1, np.array([[1.0, 3.0]]), 'rbf', 'gh', {'degree': 15})
        # 1-D x0: scipy >= 1.11 rejects the (1, 2) array of tests/test_bqmod.py:167
        res = m.optimize(np.log([1.0, 0.5]), f1(m.points).T, m.points, **kw)
        d.update({nm + '_opt_x': res.x, nm + '_opt_fun': np.asarray(res.fun), nm + '_opt_nit': np.asarray(res.nit)})
    d['opt_bounds'] = np.array(bounds)
    d['cases'] = np.array(cases)
    save('nlml', **d)


def gen_large_pointsets():
    """Gauss-Hermite Kalman filters on the 5-D models: 3^5 = 243 points at the default degree (mtran.py:309-360), the
    configuration the reference's own test runs on every model (tests/test_ssinf.py:135-149).  4-D constant velocity:
    81 points; degree 5 on the 2-D pendulum: 25 points (below the streaming threshold, for comparison)."""
    np.random.seed(2)
    dyn, obs, x, y = reentry(60, 2)
    filter_case('c3_reentry_ghkf3', ssinf.GaussHermiteKalman(dyn, obs), x, y)
    np.random.seed(3)
    dyn, obs, x, y = coordinated_turn(60, 2)
    filter_case('c4_ct_ghkf3', ssinf.GaussHermiteKalman(dyn, obs), x, y)
    np.random.seed(41)
    m0, P0 = np.array([10175, 295, 980, -35.0]), np.diag([10000, 100, 10000, 100.0])
    dyn = ssmod.ConstantVelocity(GaussRV(4, m0, P0), GaussRV(2, cov=np.diag([50, 5.0])), dt=0.5)
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=np.diag([50, 0.4e-6])), 4)
    x = dyn.simulate_discrete(60, mc_sims=2)
    y = obs.simulate_measurements(x)
    filter_case('c8_cv_ghkf3', ssinf.GaussHermiteKalman(dyn, obs), x, y)


def gen_weights():
    """BQ weights and kernel expectations (bqmod.py:495-523, 893-992; bqkern.py:329-424)."""
    cases = []
    for dim, pts, php, par in [
        (1, 'ut', None, [1.0, 3.0]), (1, 'ut', {'kappa': 0.0}, [1.0, 1e-3]), (1, 'ut', {'kappa': 0.0}, [1.0, 1e2]),
        (1, 'sr', None, [1.0, 0.3]), (1, 'gh', {'degree': 5}, [1.0, 0.3]), (1, 'gh', {'degree': 20}, [1.0, 0.1]),
        (2, 'ut', None, [1.0, 1.0, 1.0]), (2, 'ut', None, [2.5, 3.0, 0.7]), (2, 'gh', {'degree': 3}, [1.0, 2.0, 2.0]),
        (5, 'ut', None, [1.0, 25, 25, 25, 25, 25]), (5, 'ut', None, [1.0, 25, 25, 1e4, 1e4, 1e4]),
        (5, 'ut', None, [1.0, 1, 1, 1, 1, 1]), (5, 'sr', None, [1.0, 3, 3, 3, 3, 3]),
        (5, 'fs', {'degree': 3, 'kappa': None, 'dof': 6.0}, [1.0, 2, 2, 2, 2, 2]),
    ]:
        cases.append((dim, pts, php, np.array([par])))
    d = {'n': np.asarray(len(cases))}
    for i, (dim, pts, php, par) in enumerate(cases):
        p = 'w{:02d}_'.format(i)
        gp = bqmod.GaussianProcessModel(dim, par, 'rbf', pts, php)
        wm, Wc, Wcc, emv, ivar = gp.bq_weights(par)
        x = gp.points
        d.update({p + 'dim': dim, p + 'pts': pts, p + 'php': json.dumps(php), p + 'par': par, p + 'points': x,
                  p + 'K': gp.kernel.eval(par, x), p + 'iK': gp.iK, p + 'q': gp.q, p + 'Q': gp.Q,
                  p + 'R': gp.kernel.exp_x_xkx(par, x), p + 'kbar': gp.kernel.exp_xy_kxy(par),
                  p + 'gp_wm': wm, p + 'gp_Wc': Wc, p + 'gp_Wcc': Wcc, p + 'gp_emv': emv, p + 'gp_ivar': ivar})
        # Bayes-Sard with UT multi-index (pi-unisolvent special case only when N == 2D+1)
        if pts in ('ut', 'fs'):
            mi = MUL_UT(dim)
            bs = bqmod.BayesSardModel(dim, par, mi, pts, php)
            wm, Wc, Wcc, emv, ivar = bs.bq_weights(par, mi)
            d.update({p + 'bs_mulind': mi, p + 'bs_wm': wm, p + 'bs_Wc': Wc, p + 'bs_Wcc': Wcc,
                      p + 'bs_emv': emv, p + 'bs_ivar': ivar})
        # Bayes-Sard general case (fewer basis functions than points): total degree <= 1
        mi = np.hstack((np.zeros((dim, 1)), np.eye(dim))).astype(int)
        if mi.shape[1] < x.shape[1]:
            bs = bqmod.BayesSardModel(dim, par, mi, pts, php)
            wm, Wc, Wcc, emv, ivar = bs.bq_weights(par, mi)
            d.update({p + 'bsg_mulind': mi, p + 'bsg_wm': wm, p + 'bsg_Wc': Wc, p + 'bsg_Wcc': Wcc,
                      p + 'bsg_emv': emv, p + 'bsg_ivar': ivar})
    save('weights', **{k: np.asarray(v) for k, v in d.items()})

    # classical point sets and weights (mtran.py:166-520)
    d = {}
    for dim in (1, 2, 5):
        d['ut{}_pts'.format(dim)] = mtran.UnscentedTransform.unit_sigma_points(dim)
        d['ut{}_wm'.format(dim)], d['ut{}_wc'.format(dim)] = mtran.UnscentedTransform.weights(dim)
        d['ut{}k0_pts'.format(dim)] = mtran.UnscentedTransform.unit_sigma_points(dim, kappa=0.0)
        d['ut{}k2a_wm'.format(dim)], d['ut{}k2a_wc'.format(dim)] = mtran.UnscentedTransform.weights(dim, 2.0, 0.5, 1.0)
        d['ut{}k2a_pts'.format(dim)] = mtran.UnscentedTransform.unit_sigma_points(dim, 2.0, 0.5)
        d['sr{}_pts'.format(dim)] = mtran.SphericalRadialTransform.unit_sigma_points(dim)
        d['sr{}_wm'.format(dim)] = mtran.SphericalRadialTransform.weights(dim)
        for deg in (3, 5):
            d['fs{}d{}_pts'.format(dim, deg)] = mtran.FullySymmetricStudentTransform.unit_sigma_points(dim, deg, None, 6.0)
            d['fs{}d{}_wm'.format(dim, deg)] = mtran.FullySymmetricStudentTransform.weights(dim, deg, None, 6.0)
    for dim, deg in ((1, 3), (1, 5), (1, 20), (2, 3), (2, 5), (5, 3)):
        d['gh{}d{}_pts'.format(dim, deg)] = mtran.GaussHermiteTransform.unit_sigma_points(dim, deg)
        d['gh{}d{}_wm'.format(dim, deg)] = mtran.GaussHermiteTransform.weights(dim, deg)
    save('pointsets', **d)


def gen_simulation():
    """Simulators with injected noise (ssmod.py:168-244, 1011-1039)."""
    rng = np.random.RandomState(11)
    d = {}
    steps, mc = 40, 3
    for name, mk in (('ungm', ungm), ('pend', pendulum), ('reentry', reentry), ('ct', coordinated_turn)):
        np.random.seed(5)
        dyn, obs, _, _ = mk(4, 1)
        dx, dq, dr = dyn.dim_state, dyn.dim_noise, obs.dim_noise
        x0 = dyn.init_rv.mean[:, None] + rng.randn(dx, mc) * 0.1
        q = rng.randn(dq, steps, mc) * np.sqrt(np.diag(dyn.noise_rv.cov))[:, None, None]
        r = rng.randn(dr, steps, mc) * np.sqrt(np.diag(obs.noise_rv.cov))[:, None, None]
        dyn.init_rv = InjectedRV(dyn.init_rv, [x0])
        dyn.noise_rv = InjectedRV(dyn.noise_rv, [q])
        obs.noise_rv = InjectedRV(obs.noise_rv, [r])
        x = dyn.simulate_discrete(steps, mc_sims=mc)
        y = obs.simulate_measurements(x)
        d.update({name + '_x0': x0, name + '_q': q, name + '_r': r, name + '_x': x, name + '_y': y})
        d.update({name + '_' + k: v for k, v in model_dict(dyn, obs).items()})
        if name == 'reentry':
            # Euler-Maruyama (ssmod.py:201-244): q has steps+1 slices, result drops x0
            dt = 0.05
            qc = rng.randn(dq, steps + 1, mc) * np.sqrt(np.diag(dyn.noise_rv.cov))[:, None, None]
            dyn.init_rv.samples, dyn.noise_rv.samples = [x0], [qc]
            xc = dyn.simulate_continuous(duration=steps * dt, dt=dt, mc_sims=mc)
            d.update({'reentry_qc': qc, 'reentry_xc': xc, 'reentry_dtc': np.asarray(dt)})
    save('simulation', **d)


def gen_scores():
    """Scores (utils.py:18-148) aggregated like research/gpq/icinco_demo.py:17-52 (no bootstrap)."""
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_PATH, 'research', 'gpq'))
    for m in ('tqdm',):
        try:
            __import__(m)
        except ImportError:
            import types
            sys.modules[m] = types.ModuleType(m)
            sys.modules[m].trange = range
    import icinco_demo
    d = {}
    for name in ('c1_ungm_ukf', 'c5_pend_gpq', 'c3s_reentry_gpq'):
        g = np.load(os.path.join(OUT, name + '.npz'))
        x = g['x']
        mf, Pf, ms, Ps = g['fi_mean'][..., None], g['fi_cov'][..., None], g['sm_mean'][..., None], g['sm_cov'][..., None]
        sc = icinco_demo.evaluate_performance(x, mf, Pf, ms, Ps, bootstrap_variance=False)
        for k, v in zip(('rmse_f', 'nci_f', 'nll_f', 'rmse_s', 'nci_s', 'nll_s'), sc):
            d[name + '_' + k] = np.asarray(v)
        # per-step building blocks
        dx, N, M = x.shape
        mse = np.stack([utils.mse_matrix(x[:, k, :], mf[:, k, :, 0]) for k in range(N)], axis=-1)
        nll = np.array([[utils.neg_log_likelihood(x[:, k, s], mf[:, k, s, 0], Pf[:, :, k, s, 0]) for s in range(M)]
                        for k in range(N)])
        lcr = np.array([[utils.log_cred_ratio(x[:, k, s], mf[:, k, s, 0], Pf[:, :, k, s, 0], mse[..., k])
                         for s in range(M)] for k in range(N)])
        d.update({name + '_mse': mse, name + '_nll': nll, name + '_lcr': lcr})
        # tracking-style RMSE (research/bsq/bsq_tracking.py:330-337)
        se = utils.squared_error(x, mf[..., 0])
        d[name + '_rmse_vs_time'] = np.sqrt(se.sum(axis=0)).mean(axis=1)
    save('scores', **d)


def gen_c5_sweep():
    """Configuration C5: Bayes-Sard filters with the expected model variance assigned from outside
    (research/bsq/bsq_tracking.py:276-281) on the pendulum (tests/test_ssinf.py:42-51) and the coordinated turn +
    radar, scored like research/gpq/icinco_demo.py:17-52.  The files carry x, y, the reference's weights and its
    scores (RMSE / NCI / NLL, per-step MSE and credibility-ratio sums), and the filtered moments of the first 4
    trajectories; the moments of all trajectories would be 10x larger and are not needed: the score path is pinned
    by the aggregates."""
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_PATH, 'research', 'gpq'))
    import types
    if 'tqdm' not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except ImportError:
            sys.modules['tqdm'] = types.ModuleType('tqdm')
            sys.modules['tqdm'].trange = range
    import icinco_demo
    for name, fixture, M, mv, seed in (('sweep_c5_pend_bsq_mv', pendulum, 200, 1e-2, 11), ('sweep_c5_ct_bsq_mv', coordinated_turn, 100, 1e-2, 12)):
        np.random.seed(seed)
        dyn, obs, x, y = fixture(100, M)
        d = dyn.dim_in
        kp = np.ones((1, d + 1))
        alg = ssinf.BayesSardKalman(dyn, obs, kp, kp, MUL_UT(d), MUL_UT(d))
        alg.tf_dyn.model.model_var = mv * np.eye(dyn.dim_state)
        alg.tf_obs.model.model_var = 0.0 * np.eye(obs.dim_out)
        out = run_filter(alg, y, smooth=False)
        mf, Pf = out['fi_mean'], out['fi_cov']
        assert (out['status'] == 0).all(), out['status']
        sc = icinco_demo.evaluate_performance(x, mf[..., None], Pf[..., None], mf[..., None], Pf[..., None], bootstrap_variance=False)
        N = x.shape[1]
        mse = np.stack([utils.mse_matrix(x[:, k, :], mf[:, k, :]) for k in range(N)], axis=-1)
        lcr = np.array([[utils.log_cred_ratio(x[:, k, i], mf[:, k, i], Pf[:, :, k, i], mse[..., k]) for i in range(M)] for k in range(N)])
        e = {'x': x, 'y': y, 'fi_mean4': mf[..., :4], 'fi_cov4': Pf[..., :4], 'status': out['status'],
             'rmse': np.asarray(sc[0]), 'nci': np.asarray(sc[1]), 'nll': np.asarray(sc[2]), 'mse': mse, 'lcr_sum': lcr.sum(axis=1),
             'alg_name': np.asarray('BayesSardKalman'), 'assigned_model_var': np.asarray(mv)}
        e.update(model_dict(dyn, obs))
        e.update(transform_dict(alg.tf_dyn, 'dyn_'))
        e.update(transform_dict(alg.tf_obs, 'obs_'))
        save(name, **e)


def gen_weight_envelope(n_var=16, mc=32, steps=100):
    """Row a9: how far can float64 BQ weights of the C3 kernels move under perturbations the reference itself cannot
    exclude?  The obs-transform kernel matrix of research/gpq/gpq_tracking.py:41-44 has cond ~ 1e9, so
    Wc = iK Q iK carries rounding errors of order eps * cond^2.  The ensemble: the reference's own bq_weights
    (bq/bqmod.py:495-523) with Kernel._cho_inv (bq/bqkern.py:38-64) fed a kernel matrix perturbed by +-1 ulp per entry
    (symmetric), half of the members also solving with numpy.linalg.inv instead of cho_solve -- everything else is the
    reference's code.  Each member's filter (the reference's forward_pass) runs on the same seeded data.  The file holds
    the unperturbed weights, the members' weights and, per member and trajectory, the failure step and the time-mean
    squared error: the envelope a float64 re-implementation is measured against."""
    np.random.seed(21)
    dyn, obs, x, y = reentry(steps, mc)
    hdyn = np.array([[1.0, 25, 25, 25, 25, 25]])
    hobs = np.array([[1.0, 25, 25, 1e4, 1e4, 1e4]])
    orig = bqkern.Kernel._cho_inv
    eps = np.finfo(float).eps

    def member(alg):
        out = run_filter(alg, y, smooth=False)
        se = ((out['fi_mean'] - x) ** 2).mean(axis=1)          # (dx, mc), NaN for failed trajectories
        return {'wm_dyn': alg.tf_dyn.wm, 'Wc_dyn': alg.tf_dyn.Wc, 'Wcc_dyn': alg.tf_dyn.Wcc, 'mv_dyn': np.asarray(alg.tf_dyn.model.model_var),
                'wm_obs': alg.tf_obs.wm, 'Wc_obs': alg.tf_obs.Wc, 'Wcc_obs': alg.tf_obs.Wcc, 'mv_obs': np.asarray(alg.tf_obs.model.model_var),
                'status': out['status'], 'mse_time': se}

    base_alg = ssinf.GaussianProcessKalman(dyn, obs, hdyn, hobs, kernel='rbf', points='ut')
    base = member(base_alg)
    members = []
    try:
        for p in range(n_var):
            rs = np.random.RandomState(1000 + p)

            def pert(A, b=None, rs=rs, use_inv=(p % 2 == 1)):
                b = np.eye(A.shape[0]) if b is None else b
                E = np.triu(rs.randint(-1, 2, size=A.shape).astype(float))
                E = E + np.triu(E, 1).T
                A2 = A * (1.0 + eps * E)
                iA = np.linalg.inv(A2).dot(b) if use_inv else scipy.linalg.cho_solve(scipy.linalg.cho_factor(A2), b)
                return 0.5 * (iA + iA.T)
            bqkern.Kernel._cho_inv = staticmethod(pert)
            members.append(member(ssinf.GaussianProcessKalman(dyn, obs, hdyn, hobs, kernel='rbf', points='ut')))
    finally:
        bqkern.Kernel._cho_inv = orig
    d = {'x': x, 'y': y}
    d.update(model_dict(dyn, obs))
    d.update(transform_dict(base_alg.tf_dyn, 'dyn_'))
    d.update(transform_dict(base_alg.tf_obs, 'obs_'))
    for k, v in base.items():
        d['base_' + k] = np.asarray(v)
    for k in base:
        d['ens_' + k] = np.stack([np.asarray(m[k]) for m in members])
    save('weight_envelope_c3', **d)
    fails = [int((m['status'] != 0).sum()) for m in members]
    print('  base failures', int((base['status'] != 0).sum()), ' ensemble failures', fails)
    print('  max |Wc_obs - base| over members', max(np.abs(m['Wc_obs'] - base['Wc_obs']).max() for m in members),
          ' |base Wc_obs|max', np.abs(base['Wc_obs']).max())


def gen_public_attrs():
    """Public attributes the research code reads after forward_pass (research/bsq/bsq_tracking.py:1004-1013):
    x_mean_pr, x_cov_pr, xx_cov and the predictive MEASUREMENT moments y_mean_pr, y_cov_pr, xy_cov of the last step
    (ssinf.py:281-294), for an additive 5-D model (UKF and GPQ) and a model with non-additive noise."""
    d = {}
    np.random.seed(0)
    dyn, obs, x, y = reentry(60, 2)
    hdyn = np.array([[1.0, 25, 25, 25, 25, 25]])
    hobs = np.array([[1.0, 25, 25, 1e4, 1e4, 1e4]])
    cases = [('reentry_ukf', ssinf.UnscentedKalman(dyn, obs), y), ('reentry_gpq', ssinf.GaussianProcessKalman(dyn, obs, hdyn, hobs), y)]
    np.random.seed(5)
    x0, q, r = GaussRV(1, mean=np.array([1.0]), cov=np.atleast_2d(5.0)), GaussRV(1, cov=np.atleast_2d(10.0)), GaussRV(1)
    dna, ona = ssmod.UNGMNATransition(x0, q), ssmod.UNGMNAMeasurement(r, 1)
    xna = dna.simulate_discrete(40, mc_sims=2)
    yna = ona.simulate_measurements(xna)
    cases.append(('ungmna_ukf', ssinf.UnscentedKalman(dna, ona), yna))
    for name, alg, yy in cases:
        alg.forward_pass(yy[..., 0])
        d[name + '_y'] = yy
        for a in ('x_mean_pr', 'x_cov_pr', 'xx_cov', 'y_mean_pr', 'y_cov_pr', 'xy_cov', 'x_mean_fi', 'x_cov_fi'):
            d[name + '_' + a] = np.asarray(getattr(alg, a))
        if name == 'reentry_gpq':
            d.update(transform_dict(alg.tf_dyn, 'reentry_gpq_dyn_'))
            d.update(transform_dict(alg.tf_obs, 'reentry_gpq_obs_'))
    save('public_attrs', **d)


def gen_marginal(steps=15, mc=2):
    """MarginalizedGaussianProcessKalman (ssinf.py:1034-1296) on the UNGM of tests/test_ssinf.py:267-316 (spherical-radial
    points): per step a BFGS Laplace approximation over the kernel log-parameters and a spherical-radial mixture of the
    conditional state posteriors.  Filtered moments and the parameter posterior after every step of every trajectory."""
    np.random.seed(31)
    x0 = GaussRV(1, cov=np.atleast_2d(1.0))
    q = GaussRV(1, cov=np.atleast_2d(10.0))
    dyn = ssmod.UNGMTransition(x0, q)
    obs = ssmod.UNGMMeasurement(GaussRV(1, cov=np.atleast_2d(1.0)), 1)
    x = dyn.simulate_discrete(steps, mc_sims=mc)
    y = obs.simulate_measurements(x)
    fm, fc = np.zeros((1, steps, mc)), np.zeros((1, 1, steps, mc))
    pm, pc = np.zeros((4, steps, mc)), np.zeros((4, 4, steps, mc))
    for i in range(mc):
        alg = ssinf.MarginalizedGaussianProcessKalman(dyn, obs, 'rbf', 'sr')
        # step by step, to record the parameter posterior of every step (forward_pass: ssinf.py:101-111)
        for k in range(1, steps + 1):
            alg._time_update(k - 1)
            alg._measurement_update(y[:, k - 1, i], k)
            fm[:, k - 1, i], fc[:, :, k - 1, i] = alg.x_mean_fi, alg.x_cov_fi
            pm[:, k - 1, i], pc[:, :, k - 1, i] = alg.param_mean, alg.param_cov
    # the two building blocks at fixed parameter vectors (first step of trajectory 0): un-normalised negative log
    # posterior (ssinf.py:1225-1245) and conditional state posterior moments (:1118-1143)
    alg = ssinf.MarginalizedGaussianProcessKalman(dyn, obs, 'rbf', 'sr')
    rs = np.random.RandomState(3)
    thetas = 0.7 * rs.randn(6, 4)
    obj = np.array([alg._param_neg_log_posterior(th, y[:, 0, 0], 1) for th in thetas])
    cm = np.array([alg._state_posterior_moments(th, y[:, 0, 0], 1)[0] for th in thetas])
    cc = np.array([alg._state_posterior_moments(th, y[:, 0, 0], 1)[1] for th in thetas])
    save('marginal_ungm', x=x, y=y, fi_mean=fm, fi_cov=fc, param_mean=pm, param_cov=pc, thetas=thetas, obj=obj, cond_mean=cm, cond_cov=cc)


def gen_structured():
    """The REFERENCE's GPQ filters run with structured weights assigned (the pattern of research/tpq/tpq_ungm.py:114-124):
    the exact values of its own formulas (oracle/exact_weights.py: mpmath, 60 digits, rounded once) projected onto the
    reflection structure they have in exact arithmetic (ssmtoybox_b200.bq.bqmod.symmetrize_reflective, pure numpy) --
    i.e. what the package's double-double weights kernel + projection produce, up to 1e-12.  These weight sets pass
    ssm_weights_reflective, so the golden parity tests run the COMPACT sums of the forward pass against the reference's
    dense numpy sums on identical inputs."""
    from exact_weights import exact_gp_weights
    sys.path.insert(0, os.path.dirname(HERE))
    from ssmtoybox_b200.bq.bqmod import symmetrize_reflective

    def assign(alg):
        for tf in (alg.tf_dyn, alg.tf_obs):
            par, pts = tf.model.kernel.par, tf.model.points
            w = exact_gp_weights(par, pts)
            s = symmetrize_reflective(pts, {k: w[k] for k in ('wm', 'Wc', 'Wcc', 'iK')})
            assert s is not w and all(np.abs(s[k] - w[k]).max() <= 1e-12 * np.abs(w[k]).max() for k in ('wm', 'Wc', 'Wcc'))
            tf.wm, tf.Wc, tf.Wcc = s['wm'], s['Wc'], s['Wcc']
            tf.model.model_var = w['model_var']
            tf.model.integral_var = w['integral_var']
        return alg

    np.random.seed(0)
    dyn, obs, x, y = reentry(500, 3)      # the data of c3_reentry_gpq
    hdyn = np.array([[1.0, 25, 25, 25, 25, 25]])
    hobs = np.array([[1.0, 25, 25, 1e4, 1e4, 1e4]])
    filter_case('c3_reentry_gpq_structured', assign(ssinf.GaussianProcessKalman(dyn, obs, hdyn, hobs, kernel='rbf', points='ut')), x, y)
    np.random.seed(3)
    dyn, obs, x, y = coordinated_turn(200, 4)      # the data of c4_ct_gpq
    par_dyn = np.array([[1.0, 1, 1, 1, 1, 1]])
    par_obs = np.array([[1.0, 1, 1e2, 1, 1e2, 1e2]])
    filter_case('c4_ct_gpq_structured', assign(ssinf.GaussianProcessKalman(dyn, obs, par_dyn, par_obs)), x[..., :2], y[..., :2])
    np.random.seed(7)
    dyn, obs, x, y = pendulum(300, 3)      # the data of c5_pend_gpq
    kp = np.array([[1.0, 1.0, 1.0]])
    filter_case('c5_pend_gpq_structured', assign(ssinf.GaussianProcessKalman(dyn, obs, kp, kp)), x, y)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    sets = {'filters': gen_filters, 'reentry1d': gen_reentry1d, 'ungmna': gen_ungmna, 'student_bq': gen_student_bq, 'more_models': gen_more_models, 'nlml': gen_nlml, 'large_pointsets': gen_large_pointsets, 'weights': gen_weights, 'simulation': gen_simulation,
            'scores': gen_scores, 'c5_sweep': gen_c5_sweep, 'weight_envelope': gen_weight_envelope, 'public_attrs': gen_public_attrs, 'marginal': gen_marginal, 'structured': gen_structured}
    for name in (sys.argv[1:] or list(sets)):   # optional: only the named sets
        sets[name]()
