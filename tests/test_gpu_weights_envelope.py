"""BQ weights in the reference's own (float64) arithmetic on the ill-conditioned C3 kernels,
and the reference's BSQ known-answer tests run against the DEVICE kernel.  Needs a B200.

The reference's C3 measurement kernel (research/gpq/gpq_tracking.py:41-44) has cond(K) ~ 1e9: its covariance weights
Wc = iK Q iK are dominated by rounding noise.  tests/golden/weight_envelope_c3.npz (oracle/gen_golden.py
gen_weight_envelope) holds an ensemble of the REFERENCE's own weights under +-1 ulp perturbations of K and
cho_solve <-> inv: Wc moves by up to 3.2 (entries are O(1)), 11 of 16 members' filters fail on every trajectory at
step 1, the other 5 run everywhere with RMSEs 30-50 % apart.  A float64 re-implementation can therefore only be
asked to stay inside that envelope -- which is what is asserted here; the bit pattern of one particular reference run
is reproduced by assigning its weights (every golden filter test does)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel

pytestmark = pytest.mark.gpu


def _gpu_weights(g, prefix, precision):
    from ssmtoybox_b200 import device as dv
    w = dv.bq_weights(g[prefix + 'kern_par'], g[prefix + 'points'], precision=precision)
    assert int(w['info'][0]) == 0
    return w


@pytest.mark.parametrize('prefix', ['dyn_', 'obs_'])
def test_float64_weights_inside_the_reference_envelope(prefix):
    g = golden('weight_envelope_c3')
    w = _gpu_weights(g, prefix, 'float64')
    s = prefix[:3]
    for key, gk in (('wm', 'wm_'), ('Wc', 'Wc_'), ('Wcc', 'Wcc_')):
        base, ens = g['base_' + gk + s], g['ens_' + gk + s]
        spread = np.abs(ens - base).max()
        dev = np.abs(w[key][0] - base).max()
        assert dev <= 3.0 * spread + 1e-15, (prefix, key, dev, spread)
    mv_base, mv_ens = float(g['base_mv_' + s]), g['ens_mv_' + s].reshape(-1)
    assert abs(float(w['model_var'][0]) - mv_base) <= 3.0 * np.abs(mv_ens - mv_base).max() + 1e-18


def _run(g, wd, wo, y):
    from ssmtoybox_b200 import device as dv
    d = dict(g)
    for p, w in (('dyn_', wd), ('obs_', wo)):
        d[p + 'wm'], d[p + 'Wc'], d[p + 'Wcc'], d[p + 'model_var'] = w['wm'][0], w['Wc'][0], w['Wcc'][0], np.asarray(w['model_var'][0])
    low = dv.lower(d)
    return dv.filter_forward(low, y, store_pred=False)


def test_filter_from_float64_weights_inside_the_reference_envelope():
    """The C3 filter built from the device's float64-mode weights behaves like a member of the reference's own
    ensemble: either it fails on every trajectory at the first step (11 of 16 members do) or it runs everywhere with
    per-component RMSEs inside the range the surviving members span.  The default double-double weights (the correctly
    rounded values of the same formulas) always run."""
    g = golden('weight_envelope_c3')
    x, y = g['x'], torch.as_tensor(np.ascontiguousarray(g['y']), device='cuda')
    M = x.shape[2]
    st_ens = g['ens_status']                                  # (members, M): failing step (1-based) or 0
    assert set(np.unique((st_ens != 0).sum(axis=1))) <= {0, M}   # the reference's own outcomes: all or nothing
    alive = [p for p in range(st_ens.shape[0]) if (st_ens[p] == 0).all()]
    rm = lambda mse: np.sqrt(mse).mean(axis=-1)               # noqa: E731  (dx,)
    rms = np.stack([rm(g['ens_mse_time'][p]) for p in alive] + [rm(g['base_mse_time'])])
    lo, hi = rms.min(axis=0), rms.max(axis=0)
    out = _run(g, _gpu_weights(g, 'dyn_', 'float64'), _gpu_weights(g, 'obs_', 'float64'), y)
    st = out['status'].cpu().numpy()
    n_fail = int((st != 0).sum())
    assert n_fail in (0, M), n_fail
    if n_fail == M:
        assert set(st >> 8) == set(np.unique(st_ens[st_ens != 0]))          # the same step as the failing members
    else:
        got = rm(((out['fi_mean'].cpu().numpy() - x) ** 2).mean(axis=1))
        width = hi - lo
        assert np.all(got >= lo - 0.5 * width) and np.all(got <= hi + 0.5 * width), (got, lo, hi)
    # double-double weights: no failures, and an error no larger than the best member of the float64 ensemble allows
    out = _run(g, _gpu_weights(g, 'dyn_', 'dd'), _gpu_weights(g, 'obs_', 'dd'), y)
    assert int((out['status'] != 0).sum()) == 0
    got = rm(((out['fi_mean'].cpu().numpy() - x) ** 2).mean(axis=1))
    assert np.all(got <= hi + 0.5 * (hi - lo)), (got, hi)


def test_facade_weight_precision_switch():
    """bq.bqmod.weight_precision('float64'): the drop-in constructors compute their weights in the reference's arithmetic;
    on a well-conditioned kernel both modes agree with the reference to rounding."""
    from ssmtoybox_b200.bq import bqmod
    from ssmtoybox_b200.bq.bqmtran import GaussianProcessTransform
    g = golden('weights')
    p = 'w00_'
    par = g[p + 'par']
    dim = g[p + 'points'].shape[0]
    assert bqmod.get_weight_precision() == 'dd'
    t_dd = GaussianProcessTransform(dim, 1, par, 'rbf', 'ut')
    with bqmod.weight_precision('float64'):
        assert bqmod.get_weight_precision() == 'float64'
        t_64 = GaussianProcessTransform(dim, 1, par, 'rbf', 'ut')
    assert bqmod.get_weight_precision() == 'dd'
    if np.array_equal(t_dd.model.points, g[p + 'points']):
        assert rel(t_64.Wc, g[p + 'gp_Wc']) < 1e-10 and rel(t_dd.Wc, g[p + 'gp_Wc']) < 1e-10
    assert rel(t_64.Wc, t_dd.Wc) < 1e-10
    with pytest.raises(ValueError):
        bqmod.set_weight_precision('float32')


# ---- the reference's BSQ known-answer tests against the device kernel (tests/test_bqmod.py:368-474) -------------
def _bs(dim, par, points, mulind, precision):
    from ssmtoybox_b200 import device as dv
    w = dv.bq_weights(np.atleast_2d(par), points, np.asarray(mulind), precision=precision)
    assert int(w['info'][0]) == 0
    return w['wm'][0], w['Wc'][0], float(w['model_var'][0]), float(w['integral_var'][0])


@pytest.mark.parametrize('precision', ['float64', 'dd'])
def test_device_bsq_reproduces_classical_rules(precision):
    """BSQ mean weights with as many polynomial basis functions as points are the classical rule's weights: UT
    (kappa = 0 and 2), spherical-radial, Gauss-Hermite 5 (1-D) and 3 (2-D); covariance weights positive definite,
    expected model variance and integral variance non-negative (tests/test_bqmod.py:368-459)."""
    from ssmtoybox_b200.mtran import UnscentedTransform, SphericalRadialTransform, GaussHermiteTransform
    p1, p2 = np.array([[1.0, 3.0]]), np.array([[1.0, 1.0, 1.0]])       # ker_par_1d / the 2-D parameters of the reference's tests
    cases = [
        (1, p1, UnscentedTransform.unit_sigma_points(1, alpha=1.0), [[0, 1, 2]], UnscentedTransform.weights(1)[0]),
        (1, p1, UnscentedTransform.unit_sigma_points(1, kappa=2, alpha=1), [[0, 1, 2]], UnscentedTransform.weights(1, kappa=2, alpha=1)[0]),
        (1, p1, GaussHermiteTransform.unit_sigma_points(1, degree=5), [[0, 1, 2, 3, 4]], GaussHermiteTransform.weights(1, degree=5)),
        (2, p2, UnscentedTransform.unit_sigma_points(2, alpha=1.0), [[0, 1, 0, 2, 0], [0, 0, 1, 0, 2]], UnscentedTransform.weights(2)[0]),
        (2, p2, GaussHermiteTransform.unit_sigma_points(2, degree=3),
         [[0, 1, 0, 1, 2, 0, 1, 2, 2], [0, 0, 1, 1, 0, 2, 2, 1, 2]], GaussHermiteTransform.weights(2, degree=3)),
    ]
    for dim, par, pts, mi, want in cases:
        wm, Wc, emv, ivar = _bs(dim, par, pts, mi, precision)
        assert np.allclose(wm, want), (dim, pts.shape, wm, want)
        assert emv >= -1e-7 and ivar >= -1e-7      # zero up to rounding for the GH-5 rule (the reference's own value here: -5e-9)
        np.linalg.cholesky(Wc)
    # SR weights == UT weights for kappa = 0, alpha = 1 without the centre point (tests/test_bqmod.py:402-415)
    wm, Wc, emv, ivar = _bs(1, p1, UnscentedTransform.unit_sigma_points(1, kappa=0, alpha=1), [[0, 1, 2]], precision)
    assert np.allclose(wm[1:], SphericalRadialTransform.weights(1))
    # 5-D, the reentry kernel parameters (an expectedFailure in the reference for the Cholesky of Wc, :461-474)
    mi5 = np.hstack((np.zeros((5, 1)), np.eye(5), 2 * np.eye(5))).astype(int)
    wm, Wc, emv, ivar = _bs(5, np.array([[1.0, 25, 25, 25, 25, 25]]), UnscentedTransform.unit_sigma_points(5), mi5, precision)
    assert np.allclose(wm, UnscentedTransform.weights(5)[0], atol=1e-6)
    # both arithmetic modes give an expected model variance of -6.3e-6 here: it is the value of the reference's formula
    # (the 1e-8 jitters on K and V' K^-1 V, bqmod.py:936, outweigh a variance this close to zero), not rounding noise --
    # the reference's own test of this case is an expectedFailure
    assert emv >= -1e-4 and ivar >= -1e-4
