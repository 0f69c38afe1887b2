"""The reflection-symmetric form of the BQ moment transform (ssm_filter.cuh, SSM_TF_BQR) against the dense sums of
bqmtran.py:175-223 on the SAME weights, and both against the longdouble oracle.  The weights are the package's own
(double-double, bq/bqmod.py symmetrize_reflective): in exact arithmetic the two forms are the same number."""
import ctypes as C

import numpy as np
import pytest
import torch

import ssm_oracle as so
from conftest import golden, relstep, rel, one_step_problems

pytestmark = pytest.mark.gpu


def T(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def N_(t):
    return t.cpu().numpy()


def cum_step_err(a, b):
    """cumulative maximum over time of the per-(step, trajectory) max-norm relative error: (N, M)"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    ax = tuple(range(a.ndim - 2))
    e = np.abs(a - b).max(axis=ax) / np.maximum(np.abs(b).max(axis=ax), 1e-300)
    return np.maximum.accumulate(np.nan_to_num(e, nan=np.inf), axis=0)


def own_weights(g, symmetric=True):
    """the golden case with the reference's weights replaced by ssm_bq_weights ('dd') + the structure projection"""
    from ssmtoybox_b200 import device as dv
    from ssmtoybox_b200.bq import bqmod
    g2 = dict(g)
    for pfx in ('dyn_', 'obs_'):
        mul = g[pfx + 'mulind'] if (pfx + 'mulind') in g else None
        w = dv.bq_weights(g[pfx + 'kern_par'], g[pfx + 'points'], mul)
        assert int(w['info'][0]) == 0
        w1 = {k: w[k][0] for k in ('wm', 'Wc', 'Wcc', 'iK')}
        if symmetric:
            w2 = bqmod.symmetrize_reflective(g[pfx + 'points'], w1)
            assert w2 is not w1, 'the computed weights do not have the reflection structure'
            for k in ('wm', 'Wc', 'Wcc', 'iK'):
                assert rel(w2[k], w1[k]) < 1e-9, (pfx, k, rel(w2[k], w1[k]))     # the projection removes rounding noise only
            w1 = w2
        for k in ('wm', 'Wc', 'Wcc'):
            g2[pfx + k] = w1[k]
        if (pfx + 'iK') in g:
            g2[pfx + 'iK'] = w1['iK']
        if str(g[pfx + 'kind']) != 'tp':
            g2[pfx + 'model_var'] = w['model_var'][0]
    return g2


def run(g, y, x=None, **kw):
    from ssmtoybox_b200 import device as dv
    low = dv.lower(g)
    if x is not None:
        o = dv.filter_scored(low, T(y), T(x), **kw)
    else:
        o = dv.filter_forward(low, T(y), store_pred=True, **kw)
    torch.cuda.synchronize()
    return low, o


@pytest.mark.parametrize('name', ['c3_reentry_gpq', 'c4_ct_gpq', 'c5_pend_gpq', 'c1_ungm_gpq_ut', 'c4_ct_bsq', 'c5_pend_bsq', 'c4_ct_tpq'])
def test_compact_sums_equal_dense_sums_on_structured_weights(name, monkeypatch):
    from ssmtoybox_b200 import _lib
    g = golden(name)
    g2 = own_weights(g)
    y = g['y']
    low, o = run(g2, y)
    if True:    # (for a TPQ transform the predicate looks at the folded BQ weights the dispatch will run)
        assert _lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_dyn)) == 1 and _lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_obs)) == 1
    monkeypatch.setenv('SSM_REFL', '0')
    assert _lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_dyn)) == 0
    _, d = run(g2, y)
    monkeypatch.delenv('SSM_REFL')
    st_o, st_d = N_(o['status']), N_(d['status'])
    ld = so.forward_pass(g2, y, backend='loops', dtype=np.longdouble)
    ok = (st_o == 0) & (st_d == 0) & (ld['status'] == 0)
    assert ok.any()
    # Same weights, two summation orders; arbiter = the longdouble oracle on those weights.
    # (1) Per-step arithmetic: every (trajectory, step) pair of the golden run as an independent one-step problem that
    # restarts from the reference's filtered moments -- the compact sums are at most 4x as far from the arbiter as the
    # dense sums are (the bound test_bq_noise_floor holds the dense sums to against the reference's own arithmetic).
    p = one_step_problems(g)
    sel = slice(None, None, 3)
    init = dict(init_mean=T(p['init_mean'][..., sel]), init_cov=T(p['init_cov'][..., sel]), t_offset=T(p['t0'][sel], torch.int32))
    ld1 = so.forward_pass(g2, p['y'][..., sel], backend='loops', dtype=np.longdouble, init_mean=p['init_mean'][..., sel],
                          init_cov=p['init_cov'][..., sel], t0=p['t0'][sel])
    _, o1 = run(g2, p['y'][..., sel], **init)
    monkeypatch.setenv('SSM_REFL', '0')
    _, d1 = run(g2, p['y'][..., sel], **init)
    monkeypatch.delenv('SSM_REFL')
    ok1 = (N_(o1['status']) == 0) & (N_(d1['status']) == 0) & (ld1['status'] == 0)
    assert ok1.sum() >= 0.9 * ok1.size
    for k in ('fi_mean', 'fi_cov', 'pr_mean', 'pr_cov', 'pr_xx_cov'):
        truth = np.asarray(ld1[k], dtype=np.float64)[..., ok1]
        if k.startswith('pr_'):
            truth = truth[..., 1:, :]      # the oracle keeps the reference's slot 0 (initial moments)
        e_c, e_d = relstep(N_(o1[k])[..., ok1], truth), relstep(N_(d1[k])[..., ok1], truth)
        assert e_c <= 4.0 * e_d + 1e-12, (k, e_c, e_d)
    # (2) Whole trajectories (not for the noise-dominated BSQ filters, which drift apart from ANY second evaluation:
    # tests/test_gpu_parity.py::test_noise_dominated_filters_against_the_longdouble_arbiter): direct agreement of the two
    # forms, and the cumulative error against the arbiter at most 16x the dense sums'
    if 'bsq' not in name:
        for k, tol in (('fi_mean', 1e-8), ('pr_mean', 1e-8), ('fi_cov', 2e-5), ('pr_cov', 2e-5)):
            assert relstep(N_(o[k])[..., ok], N_(d[k])[..., ok]) < tol, (k, relstep(N_(o[k])[..., ok], N_(d[k])[..., ok]))
        for k in ('fi_mean', 'fi_cov', 'pr_mean', 'pr_cov', 'pr_xx_cov'):
            truth = np.asarray(ld[k], dtype=np.float64)[..., ok]
            if k.startswith('pr_'):
                truth = truth[..., 1:, :]
            e_c, e_d = cum_step_err(N_(o[k])[..., ok], truth), cum_step_err(N_(d[k])[..., ok], truth)
            assert np.all(e_c <= 16.0 * e_d + 1e-11), (k, float((e_c / np.maximum(e_d, 1e-16)).max()))
    # trajectories fail (or not) the same way, up to the ones the arbiter itself calls marginal
    assert np.array_equal(st_o != 0, st_d != 0) or ((st_o != 0) != (st_d != 0)).sum() <= max(1, int(0.02 * st_o.size))


def test_compact_sums_in_the_scoring_forward_pass(monkeypatch):
    """ssm_filter_scores takes the same branch: statistics equal to the dense instantiation's to the floor of the sums."""
    from ssmtoybox_b200 import device as dv, utils as U
    g2 = own_weights(golden('c4_ct_bsq'))
    low = dv.lower(g2)
    truth = {'m0': g2['m0'], 'P0': g2['P0'], 'q_cov': g2['q_cov'], 'r_cov': g2['r_cov']}
    x, ys = dv.simulate(low, 2048, 60, rng=dv.make_rng(truth, seed=7))
    a = U.evaluate_scored(dv.filter_scored(low, ys, x))
    monkeypatch.setenv('SSM_REFL', '0')
    b = U.evaluate_scored(dv.filter_scored(low, ys, x))
    assert np.allclose(a['rmse'], b['rmse'], rtol=1e-6) and abs(a['nci'] - b['nci']) < 1e-4 and abs(a['nll'] - b['nll']) < 1e-4 * max(1.0, abs(b['nll']))


def test_facade_filters_take_the_compact_form():
    """GaussianProcessKalman built through the public constructor (own weights, default switches) lowers to transforms the
    predicate accepts; with set_weight_symmetry(False) or 'float64' weights it does not."""
    from ssmtoybox_b200 import _lib, device as dv
    from ssmtoybox_b200.bq import bqmod
    import bench
    alg, _ = bench.build_filter(weights='own')
    low = dv.lower(alg._describe())
    assert _lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_dyn)) == 1 and _lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_obs)) == 1
    old = bqmod.set_weight_symmetry(False)
    try:
        alg2, _ = bench.build_filter(weights='own')
        low2 = dv.lower(alg2._describe())
        assert _lib.lib.ssm_weights_reflective(C.byref(low2.desc.tf_obs)) == 0
    finally:
        bqmod.set_weight_symmetry(old)


def test_compact_sums_at_the_benchmark_size(monkeypatch):
    """125 000 trajectories (the C3 share of one GPU, ticket-scheduled persistent grid) x 100 steps: the compact and the
    dense sums on the same structured weights fail nowhere and give the same aggregate scores."""
    from ssmtoybox_b200 import device as dv, utils as U
    g2 = own_weights(golden('c3_reentry_gpq'))
    low = dv.lower(g2)
    truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
             'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g2['r_cov']}
    x, ys = dv.simulate(low, 125000, 100, rng=dv.make_rng(truth, seed=11), mode='continuous', dt=0.05, sub=2)
    out = []
    for refl in ('1', '0'):
        monkeypatch.setenv('SSM_REFL', refl)
        f = dv.filter_forward(low, ys, store_pred=True)
        assert int((f['status'] != 0).sum()) == 0
        out.append((U.evaluate_scored(dv.smooth_scores(low.dx, f, x)), f['fi_mean'][:, -1].cpu().numpy()))
        del f
    (a, ma), (b, mb) = out
    assert np.allclose(a['rmse'], b['rmse'], rtol=1e-6) and abs(a['nci'] - b['nci']) < 1e-5 and abs(a['nll'] - b['nll']) < 1e-5 * abs(b['nll'])
    assert np.abs(ma - mb).max() / np.abs(mb).max() < 1e-8
