"""The numpy oracle (oracle/ssm_oracle.py) against the golden vectors produced by the unmodified
reference (oracle/gen_golden.py).  CPU only.  This is what pins the oracle: every other parity test
compares the CUDA path with the oracle."""
import json

import numpy as np
import pytest

import ssm_oracle as so
from conftest import golden, golden_filter_cases, relstep, rel, one_step_problems, FULL_TOL, ONE_STEP_COV_TOL, MEAN_FLOOR

CASES = golden_filter_cases()
FAST = [c for c in CASES if not c.startswith('c2_') or c in ('c2_ungm_gpq_el00', 'c2_ungm_gpq_el07', 'c2_ungm_gpq_el10')]


@pytest.mark.parametrize('name', FAST)
def test_forward_full_trajectory_lapack(name):
    """Same library calls in the same order as the reference -> agreement far below 1e-9 wherever the
    recursion does not amplify rounding differences; failure steps identical."""
    g = golden(name)
    student = 'dof' in g
    fw = (so.student_forward_pass if student else so.forward_pass)(g, g['y'], backend='lapack')
    assert np.array_equal(fw['status'] >> 8, g['status'])
    tol = FULL_TOL[name]
    if tol is not None:
        assert relstep(fw['fi_mean'], g['fi_mean'], MEAN_FLOOR.get(name, 0.0)) < tol
        assert relstep(fw['fi_cov'], g['fi_cov']) < tol
        assert relstep(fw['pr_mean'][:, 1:], g['pr_mean'][:, 1:], MEAN_FLOOR.get(name, 0.0)) < tol
        assert relstep(fw['pr_cov'][:, :, 1:], g['pr_cov'][:, :, 1:]) < tol
        assert relstep(fw['pr_xx_cov'][:, :, 1:], g['pr_xx_cov'][:, :, 1:]) < 10 * tol
    if not student and np.isfinite(g['sm_mean']).any() and tol is not None:
        bw = so.backward_pass(g, fw, backend='lapack')
        assert relstep(bw['sm_mean'], g['sm_mean'], MEAN_FLOOR.get(name, 0.0)) < 10 * tol
        assert relstep(bw['sm_cov'], g['sm_cov']) < 10 * tol


@pytest.mark.parametrize('name', [c for c in FAST if 'fsstudent' not in c and c != 'c3_reentry_gpq_fail'])
@pytest.mark.parametrize('backend', ['lapack', 'loops'])
def test_forward_one_step(name, backend):
    """Per-step parity: restart every step from the reference's own filtered moments."""
    g = golden(name)
    p = one_step_problems(g)
    sel = slice(None, None, 3 if backend == 'lapack' else 1)  # the per-trajectory back-end is slow
    fw = so.forward_pass(g, p['y'][..., sel], backend=backend, init_mean=p['init_mean'][..., sel],
                         init_cov=p['init_cov'][..., sel], t0=p['t0'][sel])
    assert (fw['status'] == 0).all()
    assert relstep(fw['fi_mean'], p['fi_mean'][..., sel], MEAN_FLOOR.get(name, 0.0)) < (1e-9 if name not in ('c3_reentry_bsq',) else 1e-5)
    assert relstep(fw['fi_cov'], p['fi_cov'][..., sel]) < ONE_STEP_COV_TOL.get(name, 1e-9)


def test_smoother_off_by_one_quirk():
    """Slots N and N-1 are never smoothed."""
    g = golden('c5_pend_ukf')
    assert np.array_equal(g['sm_mean'][:, -2:], g['fi_mean'][:, -2:])
    fw = so.forward_pass(g, g['y'], backend='loops')
    bw = so.backward_pass(g, fw, backend='loops')
    assert np.array_equal(bw['sm_mean'][:, -2:], fw['fi_mean'][:, -2:])
    assert not np.array_equal(bw['sm_mean'][:, -3], fw['fi_mean'][:, -3])
    assert relstep(bw['sm_mean'], g['sm_mean']) < 1e-9 and relstep(bw['sm_cov'], g['sm_cov']) < 1e-9


def test_failure_semantics():
    g = golden('c3_reentry_gpq_fail')
    exc = json.loads(str(g['exceptions']))
    assert all(e.startswith('ValueError') for e in exc)  # scipy check_finite in cho_factor, ssinf.py:321
    for backend in ('lapack', 'loops'):
        fw = so.forward_pass(g, g['y'], backend=backend)
        assert np.array_equal(fw['status'] >> 8, g['status'])
        assert np.array_equal(fw['status'] & 0xFF, [so.FAIL_NONFINITE_GAIN] * 2)
    g = golden('c2_ungm_gpq_el10')
    fw = so.forward_pass(g, g['y'], backend='loops')
    assert np.array_equal(fw['status'] >> 8, g['status']) and (fw['status'][1] & 0xFF) in (so.FAIL_CHOL_DYN, so.FAIL_CHOL_OBS)


def test_bq_weights_exact():
    g = golden('weights')
    for i in range(int(g['n'])):
        p = 'w{:02d}_'.format(i)
        par, x = g[p + 'par'], g[p + 'points']
        w = so.gp_weights(par, x)
        assert np.array_equal(so.rbf_eval(par, x), g[p + 'K'])
        for k, gk in (('iK', 'iK'), ('q', 'q'), ('Q', 'Q'), ('R', 'R'), ('wm', 'gp_wm'), ('Wc', 'gp_Wc'), ('Wcc', 'gp_Wcc')):
            assert np.array_equal(w[k], g[p + gk]), (i, k)
        assert w['model_var'] == g[p + 'gp_emv'] and w['integral_var'] == g[p + 'gp_ivar']
        assert so.rbf_exp_xy_kxy(par) == g[p + 'kbar']
        for b in ('bs', 'bsg'):
            if p + b + '_wm' in g:
                w = so.bs_weights(par, x, g[p + b + '_mulind'])
                for k in ('wm', 'Wc', 'Wcc'):
                    assert np.array_equal(w[k], g[p + b + '_' + k]), (i, b, k)
                assert w['model_var'] == g[p + b + '_emv'] and w['integral_var'] == g[p + b + '_ivar']


def test_bsq_reproduces_classical_rules():
    """BSQ mean weights with the UT multi-index equal the UT weights, covariance weights are positive
    definite and EMV / IVAR non-negative (reference tests/test_bqmod.py:368-459; the 5-D PD check is an
    expectedFailure there, :461-474)."""
    for dim in (1, 2, 5):
        mi = np.hstack((np.zeros((dim, 1)), np.eye(dim), 2 * np.eye(dim))).astype(int)
        w = so.bs_weights(np.ones((1, dim + 1)), so.ut_points(dim), mi)
        assert np.allclose(w['wm'], so.ut_weights(dim)[0])
        assert w['model_var'] >= 0 and w['integral_var'] >= 0
        if dim < 5:
            np.linalg.cholesky(w['Wc'])


def test_weights_invariant_to_kernel_scale():
    """reference tests/test_bqmtran.py:40-46 (array_equal)."""
    x = so.ut_points(2)
    a = so.gp_weights(np.array([[1.0, 1.5, 0.7]]), x)
    b = so.gp_weights(np.array([[7.3, 1.5, 0.7]]), x)
    for k in ('wm', 'Wc', 'Wcc'):
        assert np.array_equal(a[k], b[k])
    assert a['model_var'] >= 0 and a['integral_var'] >= 0


def test_rbf_known_answers():
    """RBFGauss.eval and exp_x_kx against naive loops (reference tests/test_bqkern.py:23-97)."""
    rng = np.random.RandomState(0)
    x = rng.randn(2, 6)
    par = np.array([[1.3, 0.8, 2.0]])
    K = so.rbf_eval(par, x)
    q = so.rbf_exp_x_kx(par, x)
    lam = np.diag(par[0, 1:] ** 2)
    for i in range(6):
        for j in range(6):
            d = x[:, i] - x[:, j]
            assert np.isclose(K[i, j], par[0, 0] ** 2 * np.exp(-0.5 * d.dot(np.linalg.inv(lam)).dot(d)), rtol=1e-13)
        qi = np.linalg.det(np.linalg.inv(lam) + np.eye(2)) ** -0.5 * np.exp(-0.5 * x[:, i].dot(np.linalg.inv(lam + np.eye(2))).dot(x[:, i]))
        assert np.isclose(q[i], qi, rtol=1e-13)
    Q = so.rbf_exp_x_kxkx(par, par, x)
    assert np.allclose(Q, Q.T) and np.all(np.linalg.eigvalsh(Q) > 0)  # tests/test_bqkern.py:131-140


def test_pointsets_exact():
    g = golden('pointsets')
    for dim in (1, 2, 5):
        assert np.array_equal(so.ut_points(dim), g['ut%d_pts' % dim])
        wm, wc = so.ut_weights(dim)
        assert np.array_equal(wm, g['ut%d_wm' % dim]) and np.array_equal(wc, g['ut%d_wc' % dim])
        assert np.array_equal(so.ut_points(dim, 0.0), g['ut%dk0_pts' % dim])
        wm, wc = so.ut_weights(dim, 2.0, 0.5, 1.0)
        assert np.array_equal(wm, g['ut%dk2a_wm' % dim]) and np.array_equal(wc, g['ut%dk2a_wc' % dim])
        assert np.array_equal(so.ut_points(dim, 2.0, 0.5), g['ut%dk2a_pts' % dim])
        assert np.array_equal(so.sr_points(dim), g['sr%d_pts' % dim]) and np.array_equal(so.sr_weights(dim), g['sr%d_wm' % dim])
        for deg in (3, 5):
            assert np.array_equal(so.fs_points(dim, deg, None, 6.0), g['fs%dd%d_pts' % (dim, deg)])
            assert np.array_equal(so.fs_weights(dim, deg, None, 6.0), g['fs%dd%d_wm' % (dim, deg)])
    for dim, deg in ((1, 3), (1, 5), (1, 20), (2, 3), (2, 5), (5, 3)):
        assert np.array_equal(so.gh_points(dim, deg), g['gh%dd%d_pts' % (dim, deg)])
        assert np.array_equal(so.gh_weights(dim, deg), g['gh%dd%d_wm' % (dim, deg)])


def test_simulation_injected_noise():
    g = golden('simulation')
    for name in ('ungm', 'pend', 'reentry', 'ct'):
        d = {k[len(name) + 1:]: v for k, v in g.items() if k.startswith(name + '_')}
        x = so.simulate_discrete(d, d['x0'], d['q'])
        y = so.simulate_measurements(d, x, d['r'])
        assert rel(x, d['x']) < 1e-14 and rel(y, d['y']) < 1e-14
        if name == 'reentry':
            assert rel(so.simulate_continuous(d, d['x0'], d['qc'], float(d['dtc'])), d['xc']) < 1e-14


def test_simulation_reentry1d_injected_noise():
    """ReentryVehicle1DTransition / RangeMeasurement (ssmod.py:418-426, 1146-1148): discrete, Euler-Maruyama, range"""
    d = golden('simulation_reentry1d')
    x = so.simulate_discrete(d, d['x0'], d['q'])
    assert rel(x, d['x']) < 1e-14 and rel(so.simulate_measurements(d, x, d['r']), d['y']) < 1e-14
    assert rel(so.simulate_continuous(d, d['x0'], d['qc'], float(d['dtc'])), d['xc']) < 1e-14


def test_simulation_ungmna_injected_noise():
    """non-additive noise passes through the model functions (ssmod.py:299-300, 1085-1086)"""
    d = golden('simulation_ungmna')
    x = so.simulate_discrete(d, d['x0'], d['q'])
    assert rel(x, d['x']) < 1e-14 and rel(so.simulate_measurements(d, x, d['r']), d['y']) < 1e-14


@pytest.mark.parametrize('name', ['c1_ungm_ukf', 'c5_pend_gpq', 'c3s_reentry_gpq'])
def test_scores(name):
    g, c = golden('scores'), golden(name)
    e = so.evaluate_performance(c['x'], c['fi_mean'], c['fi_cov'])
    assert rel(e['rmse'], g[name + '_rmse_f'].ravel()) < 1e-13
    assert rel(e['mse'], g[name + '_mse']) < 1e-13
    assert rel(e['nll_km'][1:], g[name + '_nll'][1:]) < 1e-11
    assert rel(e['lcr_km'][1:], g[name + '_lcr'][1:]) < 1e-8
    assert abs(e['nci'] - g[name + '_nci_f'].ravel()[0]) < 1e-9 * abs(e['nci']) + 1e-12
    assert abs(e['nll'] - g[name + '_nll_f'].ravel()[0]) < 1e-11 * abs(e['nll'])


def test_rbf_student_expectations_vs_reference_monte_carlo():
    """The scale-mixture quadrature of the oracle against the reference's own 2 * 10^6-sample Monte-Carlo estimates
    stored in the golden file (bq/bqkern.py:475-536): agreement at the Monte-Carlo noise level, and the value
    exp_xy_kxy converges to (full-batch sum, SURVEY.md f2 / DESIGN.md)."""
    g = golden('c4_ct_fsstudent_tpq')
    for w in ('dyn', 'obs'):
        par, x = g[w + '_kern_par'], g[w + '_points']
        q, R, Q, kbar = so.rbf_student_expectations(par, x, 4.0)
        n = 2e6
        assert np.all(np.abs(q - g[w + '_q']) < 6 * np.sqrt(q / n))                  # Var[k] <= E[k] since 0 <= k <= 1
        assert np.all(np.abs(Q - g[w + '_Q']) < 6 * np.sqrt(Q / n) + 1e-12)
        q2 = so.rbf_student_expectations(par, x, 4.0, n=800)
        assert rel(q2[0], q) < 1e-12 and rel(q2[2], Q) < 1e-12 and abs(q2[3] - kbar) < 1e-12   # quadrature converged
        W = so.student_bq_weights(par, x, 4.0)
        assert abs(W['integral_var'] / float(g[w + '_integral_var']) - 1) < 2e-3
        assert abs(W['model_var'] - float(g[w + '_model_var'])) < 2e-3
    # Gaussian limit: dof -> inf reproduces the closed forms of RBFGauss (bqkern.py:345-424)
    par, x = g['dyn_kern_par'], g['dyn_points']
    q, R, Q, kbar = so.rbf_student_expectations(par, x, 4e6, n=200)
    assert rel(q, so.rbf_exp_x_kx(par, x)) < 1e-5 and rel(R, so.rbf_exp_x_xkx(par, x)) < 1e-5
    assert rel(Q, so.rbf_exp_x_kxkx(par, par, x)) < 1e-5 and abs(kbar - so.rbf_exp_xy_kxy(par)) < 1e-6


def test_nlml_vs_reference():
    """neg_log_marginal_likelihood of the GP / Student-t process model and its gradient (bq/bqmod.py:537-596, 1191-1245)
    against the reference's values for batches of log-parameter vectors."""
    g = golden('nlml')
    for c in map(str, g['cases']):
        x, y, nu = g[c + '_x'], g[c + '_y'], float(g[c + '_nu'])
        for lp, v, gr in zip(g[c + '_log_par'], g[c + '_nlml'], g[c + '_grad']):
            f, df = so.gp_nlml(lp, y, x, 1e-8 * np.eye(x.shape[1]), nu if nu > 0 else None)
            assert abs(f - v) <= 1e-12 * max(abs(v), 1.0)
            assert np.abs(df - gr).max() <= 1e-8 * max(np.abs(gr).max(), 1.0)     # 15 GH points: cond(K) ~ 1e8
