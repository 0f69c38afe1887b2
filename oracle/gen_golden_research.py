"""Golden vectors for the research drivers, produced by RUNNING THE REFERENCE.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Run in the build container, where /root/reference exists:

    python oracle/gen_golden_research.py        # rewrites tests/golden/research_*.npz

What runs (unmodified reference code, imported from /root/reference under oracle/ref_shim.py):
  research/gpq/icinco_demo.py:17-71     evaluate_performance          (called as is, bootstrap off: numpy's MT19937
                                                                        resampling cannot be reproduced)
  research/gpq/icinco_demo.py:81-125    tables: algorithm list and filter/smoother loop, at a reduced size
  research/gpq/icinco_demo.py:172-213   hypers_demo: length-scale sweep WITHOUT reset() between trajectories (Q4)
  research/bsq/bsq_ungm.py:91-142       tables: algorithm list, at a reduced size (its evaluate_performance is the same
                                        function as icinco_demo's, bsq_ungm.py:27-84)
  research/bsq/bsq_tracking.py:223-349  reentry_demo(dur, mc_sims): the driver itself is called; journal_figure,
                                        joblib and the plotting routine are stubbed, the result dict is captured
  research/tpq/tpq_base.py:154-192      run_filters, eval_perf_scores   (called as is)
The driver loops that are re-stated here (because the reference hard-codes 500 x 100 inside the functions) use the
reference's own classes and call order; only the sizes differ.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

warnings.simplefilter('ignore')
ref_shim.install()

from gen_golden import save, coordinated_turn  # noqa: E402  (also installs the shim and imports the reference)
from ssmtoybox import ssinf, ssmod  # noqa: E402
from ssmtoybox.utils import GaussRV, squared_error, mse_matrix, log_cred_ratio, neg_log_likelihood  # noqa: E402

RESEARCH = os.path.join(ref_shim.REFERENCE_PATH, 'research')


def load_research(rel, name):
    """Import a research script by path.  journal_figure (LaTeX/pgf plotting set-up) and sklearn.externals.joblib
    (removed from scikit-learn) are stubbed; nothing numerical lives there."""
    if 'journal_figure' not in sys.modules:
        jf = types.ModuleType('journal_figure')
        jf.FigurePrint = type('FigurePrint', (), {'__init__': lambda self, *a, **k: None})
        sys.modules['journal_figure'] = jf
    if 'sklearn.externals' not in sys.modules or not hasattr(sys.modules['sklearn.externals'], 'joblib'):
        import sklearn
        ext = sys.modules.get('sklearn.externals') or types.ModuleType('sklearn.externals')
        jb = types.ModuleType('sklearn.externals.joblib')
        jb.dump = lambda *a, **k: None
        jb.load = lambda *a, **k: None
        ext.joblib = jb
        sys.modules['sklearn.externals'] = ext
        sys.modules['sklearn.externals.joblib'] = jb
        sklearn.externals = ext
    for modname, attr in (('matplotlib.gridspec', 'GridSpec'), ('matplotlib.lines', 'Line2D')):
        if not hasattr(sys.modules[modname], attr):
            setattr(sys.modules[modname], attr, type(attr, (), {}))
    spec = importlib.util.spec_from_file_location(name, os.path.join(RESEARCH, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_all(algorithms, z, smooth=True):
    """The filter/smoother loop of icinco_demo.py:115-125 / bsq_ungm.py:131-137."""
    dx = algorithms[0].mod_dyn.dim_state
    _, steps, sims = z.shape
    A = len(algorithms)
    mean_f, cov_f = np.zeros((dx, steps, sims, A)), np.zeros((dx, dx, steps, sims, A))
    mean_s, cov_s = np.zeros((dx, steps, sims, A)), np.zeros((dx, dx, steps, sims, A))
    for a, alg in enumerate(algorithms):
        for sim in range(sims):
            mean_f[..., sim, a], cov_f[..., sim, a] = alg.forward_pass(z[..., sim])
            if smooth:
                mean_s[..., sim, a], cov_s[..., sim, a] = alg.backward_pass()
            alg.reset()
    return mean_f, cov_f, mean_s, cov_s


SCORE_KEYS = ('rmse_f', 'nci_f', 'nll_f', 'rmse_s', 'nci_s', 'nll_s')


def gen_icinco_tables(steps=40, sims=10):
    ic = load_research('gpq/icinco_demo.py', 'ref_icinco_demo')
    np.random.seed(42)
    dyn = ssmod.UNGMTransition(GaussRV(1, cov=np.atleast_2d(5.0)), GaussRV(1, cov=np.atleast_2d(10.0)))
    obs = ssmod.UNGMMeasurement(GaussRV(1), 1)
    x = dyn.simulate_discrete(steps, mc_sims=sims)
    z = obs.simulate_measurements(x)
    kp_sr, kp_ut, kp_gh = np.array([[1.0, 0.3]]), np.array([[1.0, 3.0]]), np.array([[1.0, 0.1]])
    GPK = ssinf.GaussianProcessKalman
    algorithms = (
        ssinf.CubatureKalman(dyn, obs), ssinf.UnscentedKalman(dyn, obs),
        ssinf.GaussHermiteKalman(dyn, obs), ssinf.GaussHermiteKalman(dyn, obs), ssinf.GaussHermiteKalman(dyn, obs),
        ssinf.GaussHermiteKalman(dyn, obs), ssinf.GaussHermiteKalman(dyn, obs),
        GPK(dyn, obs, kp_sr, kp_sr, points='sr'), GPK(dyn, obs, kp_ut, kp_ut, points='ut'),
        GPK(dyn, obs, kp_sr, kp_sr, points='gh', point_hyp={'degree': 5}),
        GPK(dyn, obs, kp_gh, kp_gh, points='gh', point_hyp={'degree': 7}),
        GPK(dyn, obs, kp_gh, kp_gh, points='gh', point_hyp={'degree': 10}),
        GPK(dyn, obs, kp_gh, kp_gh, points='gh', point_hyp={'degree': 15}),
        GPK(dyn, obs, kp_gh, kp_gh, points='gh', point_hyp={'degree': 20}),
    )
    mf, Pf, ms, Ps = run_all(algorithms, z)
    sc = ic.evaluate_performance(x, mf, Pf, ms, Ps, bootstrap_variance=False)
    d = {'x': x, 'z': z, 'mean_f': mf, 'cov_f': Pf, 'mean_s': ms, 'cov_s': Ps}
    d.update({k: v for k, v in zip(SCORE_KEYS, sc)})
    # per-simulation data that bootstrap_var resamples (icinco_demo.py:21-22, 44-45), for the statistical check
    d['rmse_data_f'] = np.sqrt(np.mean(squared_error(x[..., None], mf), axis=1))
    save('research_icinco_tables', **d)


def gen_icinco_hypers(steps=40, mc=8):
    lscale = [1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1, 3, 1e1, 3e1, 1e2]
    np.random.seed(42)
    dyn = ssmod.UNGMTransition(GaussRV(1, cov=np.atleast_2d(5.0)), GaussRV(1, cov=np.atleast_2d(10.0)))
    obs = ssmod.UNGMMeasurement(GaussRV(1), 1)
    x = dyn.simulate_discrete(steps, mc_sims=mc)
    z = obs.simulate_measurements(x)
    def sweep(ls):
        mean_f, cov_f = np.zeros((1, steps, mc, len(ls))), np.zeros((1, 1, steps, mc, len(ls)))
        ok = []
        for i, el in enumerate(ls):
            kp = np.array([[1.0, el * dyn.dim_in]])
            f = ssinf.GaussianProcessKalman(dyn, obs, kp, kp, points='ut', point_hyp={'kappa': 0.0})
            try:
                for s in range(mc):  # no reset(): icinco_demo.py:195-196
                    mean_f[..., s, i], cov_f[..., s, i] = f.forward_pass(z[..., s])
                ok.append(True)
            except (np.linalg.LinAlgError, ValueError):
                ok.append(False)
        return mean_f, cov_f, ok

    # the reference driver stops with an exception when a length-scale makes the filter fail on this data: keep the
    # length-scales of its list that run through (the driver's own result for that shorter list)
    _, _, ok = sweep(lscale)
    lscale = [el for el, o in zip(lscale, ok) if o]
    print('hypers_demo: length-scales that complete:', lscale)
    mean_f, cov_f, ok = sweep(lscale)
    assert all(ok)
    L = len(lscale)
    se = squared_error(x[..., None], mean_f)
    nci, nll = se.copy(), se.copy()
    for k in range(steps):
        for i in range(L):
            M = mse_matrix(x[:, k, :], mean_f[:, k, :, i])
            for s in range(mc):
                nci[:, k, s, i] = log_cred_ratio(x[:, k, s], mean_f[:, k, s, i], cov_f[:, :, k, s, i], M)
                nll[:, k, s, i] = neg_log_likelihood(x[:, k, s], mean_f[:, k, s, i], cov_f[:, :, k, s, i])
    save('research_icinco_hypers', x=x, z=z, lscale=np.asarray(lscale), mean_f=mean_f, cov_f=cov_f,
         rmse=np.sqrt(np.mean(se, axis=1)).mean(axis=1), nci=nci.mean(axis=(1, 2)), nll=nll.mean(axis=(1, 2)))


def gen_bsq_ungm_tables(steps=40, mc=10):
    ic = load_research('gpq/icinco_demo.py', 'ref_icinco_demo')  # same evaluate_performance as bsq_ungm.py:27-84
    dyn = ssmod.UNGMTransition(GaussRV(1, cov=5.0), GaussRV(1, cov=10.0))
    obs = ssmod.UNGMMeasurement(GaussRV(1, cov=1.0), 1)
    np.random.seed(0)
    x = dyn.simulate_discrete(steps, mc)
    z = obs.simulate_measurements(x)
    par_ut, par_gh5, par_gh7 = np.array([[3.0, 0.3]]), np.array([[5.0, 0.6]]), np.array([[3.0, 0.4]])
    mi_ut = np.array([[0, 1, 2]])
    mi_gh = lambda degree: np.atleast_2d(np.arange(degree))  # noqa: E731
    GPK, BSK = ssinf.GaussianProcessKalman, ssinf.BayesSardKalman
    algorithms = (
        ssinf.UnscentedKalman(dyn, obs, alpha=1.0, beta=0.0),
        ssinf.GaussHermiteKalman(dyn, obs, deg=5), ssinf.GaussHermiteKalman(dyn, obs, deg=7),
        GPK(dyn, obs, par_ut, par_ut, kernel='rbf', points='ut', point_hyp={'alpha': 1.0}),
        GPK(dyn, obs, par_gh5, par_gh5, kernel='rbf', points='gh', point_hyp={'degree': 5}),
        GPK(dyn, obs, par_gh7, par_gh7, kernel='rbf', points='gh', point_hyp={'degree': 7}),
        BSK(dyn, obs, par_ut, par_ut, mi_ut, mi_ut, points='ut', point_hyp={'alpha': 1.0}),
        BSK(dyn, obs, par_gh5, par_gh5, mi_gh(5), mi_gh(5), points='gh', point_hyp={'degree': 5}),
        BSK(dyn, obs, par_gh7, par_gh7, mi_gh(7), mi_gh(7), points='gh', point_hyp={'degree': 7}),
    )
    mf, Pf, ms, Ps = run_all(algorithms, z)
    sc = ic.evaluate_performance(x, mf, Pf, ms, Ps, bootstrap_variance=False)
    d = {'x': x, 'z': z, 'mean_f': mf, 'cov_f': Pf, 'mean_s': ms, 'cov_s': Ps}
    d.update({k: v for k, v in zip(SCORE_KEYS, sc)})
    save('research_bsq_ungm_tables', **d)


def gen_bsq_reentry_demo(dur=2, mc_sims=5):
    bt = load_research('bsq/bsq_tracking.py', 'ref_bsq_tracking')
    captured = {}
    bt.joblib.dump = lambda data, path: captured.update(data)
    bt.reentry_demo_results = lambda data: None
    bt.reentry_demo(dur=dur, mc_sims=mc_sims)
    # the driver does not keep its measurements: redo its data generation with its seed (bsq_tracking.py:246-253)
    sysm = ssmod.ReentryVehicle2DTransition(GaussRV(5, np.array([6500, 350, -1.8, -6.8, 0.7]), np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0])),
                                            GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 0])))
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([6374, 0.0]))
    np.random.seed(0)
    x = sysm.simulate_continuous(duration=dur, dt=0.05, mc_sims=mc_sims)
    y = obs.simulate_measurements(x)
    x, y = x[:, ::2, ...], y[:, ::2, ...]
    assert np.array_equal(x, captured['x'])
    d = {'x': x, 'y': y, 'mean': captured['mean'], 'cov': captured['cov'], 'dur': np.asarray(float(dur)),
         'alg_str': np.asarray(','.join(captured['alg_str']))}
    for part in ('state', 'position', 'velocity', 'parameter'):
        d[part + '_rmse'] = captured[part]['rmse']
        d[part + '_inc'] = captured[part]['inc']
    save('research_bsq_reentry_demo', **d)


def gen_tpq_base(steps=30, mc=6):
    tb = load_research('tpq/tpq_base.py', 'ref_tpq_base')
    np.random.seed(7)
    dyn, obs, x, y = coordinated_turn(steps, mc, student=False, dt=0.1)
    kd, ko = np.array([[1.0, 1, 1, 1, 1, 1]]), np.array([[1.0, 1, 1e2, 1, 1e2, 1e2]])
    filters = [ssinf.UnscentedKalman(dyn, obs), ssinf.StudentProcessKalman(dyn, obs, kd, ko)]
    mf, Pf = tb.run_filters(filters, y)
    rmse_avg, lcr_avg = tb.eval_perf_scores(x, mf, Pf)
    save('research_tpq_base', x=x, y=y, mean_f=mf, cov_f=Pf, rmse_avg=rmse_avg, lcr_avg=lcr_avg,
         kern_par_dyn=kd, kern_par_obs=ko)


def gen_gpq_tracking(mc=6):
    """Scores behind the plots of gpq_tracking.py:72-111 and :178-313.  The two demo functions only plot and print,
    so their filter loops and score loops run here on the reference's classes / utils, at a reduced size."""
    from gen_golden import reentry1d
    # --- reentry_simple_gpq_demo
    np.random.seed(4)
    dyn, obs, x, y = reentry1d(100, mc)
    kd, ko = np.array([[0.5, 10, 10, 10]]), np.array([[0.5, 15, 20, 20]])
    alg = (ssinf.GaussianProcessKalman(dyn, obs, kd, ko, kernel='rbf', points='ut'), ssinf.UnscentedKalman(dyn, obs))
    mean, cov, _, _ = run_all(alg, y, smooth=False)
    d, steps, _ = x.shape
    err2 = np.zeros((d, steps, mc, 2))
    lcr = np.zeros((d, steps, mc, 2))
    for a in range(2):
        for k in range(steps):
            for i in range(d):
                M = mse_matrix(x[i, None, k, :], mean[i, None, k, :, a])
                for s in range(mc):
                    lcr[i, k, s, a] = log_cred_ratio(x[i, k, s], mean[i, k, s, a], cov[i, i, k, s, a], M)
            for s in range(mc):
                err2[:, k, s, a] = squared_error(x[:, k, s], mean[:, k, s, a])
    out = {'simple_x': x, 'simple_y': y, 'simple_mean': mean, 'simple_cov': cov,
           'simple_avg_rmse': np.sqrt(err2.sum(axis=0)).mean(axis=(0, 1))}
    for i, nm in enumerate(('pos', 'vel', 'theta')):
        out['simple_' + nm + '_rmse_vs_time'] = np.sqrt(err2[i]).mean(axis=1)
        out['simple_' + nm + '_inc_vs_time'] = lcr[i].mean(axis=1)
    # --- reentry_gpq_demo (5-D): its own models (gpq_tracking.py:14-44), 60 steps
    np.random.seed(5)
    disc_tau = 0.1
    m0 = np.array([6500.4, 349.14, -1.8093, -6.7967, 0.6932])
    Q = np.diag([2.4064e-5, 2.4064e-5, 0])
    sysm = ssmod.ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0])), GaussRV(3, cov=Q), dt=disc_tau)
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([sysm.R0, 0]))
    x = sysm.simulate_continuous(duration=6, dt=disc_tau, mc_sims=mc)
    y = obs.simulate_measurements(x)
    m0 = np.array([6500.4, 349.14, -1.8093, -6.7967, 0])
    dyn = ssmod.ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 1])), GaussRV(3, cov=disc_tau * Q), dt=disc_tau)
    hdyn, hobs = np.array([[1.0, 25, 25, 25, 25, 25]]), np.array([[1.0, 25, 25, 1e4, 1e4, 1e4]])
    alg = (ssinf.GaussianProcessKalman(dyn, obs, hdyn, hobs, kernel='rbf', points='ut'), ssinf.UnscentedKalman(dyn, obs))
    mean, cov, _, _ = run_all(alg, y, smooth=False)
    d, steps, _ = x.shape
    err2, lcr = np.zeros((d, steps, mc, 2)), np.zeros((steps, mc, 2))
    for a in range(2):
        for k in range(steps):
            M = mse_matrix(x[:4, k, :], mean[:4, k, :, a])
            for s in range(mc):
                err2[:, k, s, a] = squared_error(x[:, k, s], mean[:, k, s, a])
                lcr[k, s, a] = log_cred_ratio(x[:4, k, s], mean[:4, k, s, a], cov[:4, :4, k, s, a], M)
    out.update({'x': x, 'y': y, 'mean': mean, 'cov': cov, 'pos_rmse_vs_time': np.sqrt(err2[:2].sum(axis=0)).mean(axis=1),
                'inc_ind_vs_time': lcr.mean(axis=1), 'dyn_wm': alg[0].tf_dyn.wm, 'dyn_Wc': alg[0].tf_dyn.Wc, 'dyn_Wcc': alg[0].tf_dyn.Wcc,
                'obs_wm': alg[0].tf_obs.wm, 'obs_Wc': alg[0].tf_obs.Wc, 'obs_Wcc': alg[0].tf_obs.Wcc,
                'dyn_model_var': np.asarray(alg[0].tf_dyn.model.model_var), 'obs_model_var': np.asarray(alg[0].tf_obs.model.model_var)})
    save('research_gpq_tracking', **out)


if __name__ == '__main__':
    cases = {'icinco_tables': gen_icinco_tables, 'icinco_hypers': gen_icinco_hypers, 'bsq_ungm_tables': gen_bsq_ungm_tables,
             'bsq_reentry_demo': gen_bsq_reentry_demo, 'tpq_base': gen_tpq_base, 'gpq_tracking': gen_gpq_tracking}
    for name in (sys.argv[1:] or list(cases)):   # optional: only the named cases
        cases[name]()
