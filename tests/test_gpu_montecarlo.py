"""No-materialisation Monte-Carlo path (BASELINE configuration C5): scoring forward pass
(ssm_filter_scores), generator-driven driver (mc.monte_carlo_scores) and the BSQ NCI sweep.  Needs a B200."""
import numpy as np
import pytest
import torch

from conftest import golden, rel

pytestmark = pytest.mark.gpu


def eq(a, b):
    return torch.equal(torch.nan_to_num(a, nan=-1.2345e300), torch.nan_to_num(b, nan=-1.2345e300)) and \
        torch.equal(torch.isnan(a), torch.isnan(b))


def _data(name, M, N, seed=9):
    from ssmtoybox_b200 import device as dv
    g = golden(name)
    low = dv.lower(g)
    if 'reentry' in name:
        rng = dv.make_rng({'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
                           'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}, seed=seed)
        x, y = dv.simulate(low, M, N, rng=rng, mode='continuous', dt=0.05, sub=2)
    else:
        x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=seed))
    return g, low, x, y


@pytest.mark.parametrize('name,M,N,wins', [
    ('c3_reentry_gpq', 300, 40, None),
    ('c3_reentry_gpq', 300, 40, [(0, 13), (13, 30), (30, 40)]),
    ('c3_reentry_gpq', 70000, 50, None),                        # more CTAs than fit: ticket scheduler
    ('c3_reentry_ukf', 500, 30, None), ('c4_ct_tpq', 500, 30, None), ('c4_ct_bsq', 500, 30, None),
    ('c5_pend_gpq', 1000, 60, None), ('c5_pend_bsq', 1000, 60, [(0, 30), (30, 60)]), ('c1_ungm_ukf', 2000, 80, None),
    ('c8_cv_gpq', 500, 30, None), ('c6_reentry1d_gpq', 500, 30, None),
])
def test_scoring_forward_pass_equals_filter_plus_score_kernels(name, M, N, wins):
    """ssm_filter_scores keeps no moment arrays; its statistics, quadratic forms, errors and the filter itself are
    bitwise those of ssm_filter + ssm_scores_phase1 on the stored moments."""
    from ssmtoybox_b200 import device as dv, utils as U
    g, low, x, y = _data(name, M, N)
    y[:, 3, 5] = float('nan')
    ref = dv.filter_forward(low, y, store_pred=False, want_last=True)
    quad = torch.empty((N, M), dtype=torch.float64, device='cuda')
    stats, acc = dv.scores_phase1(x, ref['fi_mean'], ref['fi_cov'], ref['status'], quad=quad)
    sc = {}
    for c, (a, b) in enumerate(wins or [(0, N)]):
        dv.filter_scored(low, y, x, out=sc, window=None if wins is None else (a, b),
                         init_mean=sc['last_mean'] if c else None, init_cov=sc['last_cov'] if c else None)
    assert 'fi_mean' not in sc
    assert torch.equal(sc['status'], ref['status']) and int((ref['status'] != 0).sum()) >= 1
    ok = ref['status'] == 0
    assert eq(sc['last_mean'], ref['last_mean']) and eq(sc['last_cov'], ref['last_cov'])
    assert eq(sc['dres'][:, :, ok], (x - ref['fi_mean'])[:, :, ok])
    assert eq(sc['quad'][:, ok], quad[:, ok])
    if wins is None:
        # one-pass statistics: a trajectory that fails contributes its steps before the failure in the in-kernel pass
        # only (the stand-alone kernel drops it from every row) -- compare on data without failures
        y2 = y.clone()
        y2[:, 3, 5] = y[:, 2, 5]
        ref2 = dv.filter_forward(low, y2, store_pred=False)
        if int((ref2['status'] != 0).sum()) == 0:
            q2 = torch.empty((N, M), dtype=torch.float64, device='cuda')
            st2, acc2 = dv.scores_phase1(x, ref2['fi_mean'], ref2['fi_cov'], ref2['status'], quad=q2)
            sc2 = dv.filter_scored(low, y2, x)
            assert eq(sc2['stats'], st2) and eq(sc2['rmse_acc'], acc2) and eq(sc2['quad'], q2)
            want = U.evaluate_performance(x, ref2['fi_mean'], ref2['fi_cov'], status=ref2['status'], to_host=False)
            got = U.evaluate_scored(sc2, to_host=False)
            for k in ('rmse', 'nci', 'nll', 'mse'):
                assert eq(got[k], want[k]), k


def test_scoring_forward_pass_unsupported_filters_raise():
    from ssmtoybox_b200 import device as dv
    for name in ('c4_ct_fsstudent', 'c7_ungmna_ukf', 'c1_ungm_gpq_gh10'):
        g = golden(name)
        low = dv.lower(g)
        y = torch.as_tensor(np.ascontiguousarray(g['y']), device='cuda')
        x = torch.as_tensor(np.ascontiguousarray(g['x']), device='cuda')
        with pytest.raises(NotImplementedError):
            dv.filter_scored(low, y, x)


@pytest.mark.parametrize('smooth', [False, True])
def test_generator_driver_equals_materialised_run(smooth):
    """mc.monte_carlo_scores (chunks of simulate -> filter [-> smoother] with in-kernel scoring) == the same data
    simulated in one piece, filtered, stored and scored by evaluate_performance."""
    import bench
    from ssmtoybox_b200 import device as dv, mc, utils as U
    alg, g = bench.build_filter()
    truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
             'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
    M, N = 3000, 50
    got = mc.monte_carlo_scores(alg, M, N, truth=truth, sim='continuous', dt=0.05, sub=2, seed=4, chunk=1024, smooth=smooth)
    low = dv.lower(alg._describe())
    x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=4), mode='continuous', dt=0.05, sub=2)
    fwd = dv.filter_forward(low, y, store_pred=smooth)
    if smooth:
        sm = dv.smooth_backward(low.dx, fwd)
        want = U.evaluate_performance(x, sm['sm_mean'], sm['sm_cov'], status=sm['status'])
    else:
        want = U.evaluate_performance(x, fwd['fi_mean'], fwd['fi_cov'], status=fwd['status'])
    assert got['n_failed'] == 0 and got['kept_bytes'] == 8 * 6 * M * N
    for k in ('rmse', 'nci', 'nll', 'mse'):
        assert rel(got[k], want[k]) < 1e-11, (k, rel(got[k], want[k]))     # chunk sums are added in a different order


@pytest.mark.parametrize('name', ['sweep_c5_pend_bsq_mv', 'sweep_c5_ct_bsq_mv'])
def test_c5_bsq_scores_match_reference_driver_on_injected_data(name):
    """Configuration C5 on the reference's own data: BSQ filter with the expected model variance assigned from outside,
    weights as the reference computed them, scored in-kernel -> RMSE / NCI / NLL of the reference's loop
    (research/gpq/icinco_demo.py:17-52) and its per-step MSE matrices and credibility-ratio sums."""
    from ssmtoybox_b200 import device as dv, utils as U
    g = golden(name)
    low = dv.lower(g)
    x = torch.as_tensor(np.ascontiguousarray(g['x']), device='cuda')
    y = torch.as_tensor(np.ascontiguousarray(g['y']), device='cuda')
    sc = dv.filter_scored(low, y, x, keep_moments=True)
    assert int((sc['status'] != 0).sum()) == 0
    from conftest import relstep
    assert relstep(sc['fi_mean'][..., :4].cpu().numpy(), g['fi_mean4']) < 1e-8
    assert relstep(sc['fi_cov'][..., :4].cpu().numpy(), g['fi_cov4']) < 1e-7
    got = U.evaluate_scored(sc)
    assert rel(got['rmse'], g['rmse'][0]) < 1e-8
    assert rel(got['mse'], g['mse']) < 1e-8
    assert abs(got['nll'] - float(g['nll'].item())) < 1e-7 * max(1.0, abs(float(g['nll'].item())))
    assert abs(got['nci'] - float(g['nci'].item())) < 1e-7 * max(1.0, abs(float(g['nci'].item())))


def test_bsq_nci_sweep_driver_runs_and_is_shard_invariant():
    """research.bsq_nci_sweep: the NCI of a sweep point does not depend on the chunking (Philox keyed by the global
    trajectory index), a larger assigned model variance lowers the NCI (more conservative filter)."""
    from ssmtoybox_b200.research import bsq_nci_sweep as sw
    a = sw.bsq_nci_sweep('pendulum', mc_sims=(2000,), model_var=(1e-1, 1e-3), chunk=512)
    b = sw.bsq_nci_sweep('pendulum', mc_sims=(2000,), model_var=(1e-1, 1e-3), chunk=100000)
    for ra, rb in zip(a, b):
        assert abs(ra['nci'] - rb['nci']) < 1e-9 and ra['n_failed'] == rb['n_failed'] == 0
    assert a[0]['nci'] < a[1]['nci']
    c = sw.bsq_nci_sweep('coordturn', mc_sims=(1000,), model_var=(1e-2,))
    assert np.isfinite(c[0]['nci']) and c[0]['kept_bytes'] == 8 * 6 * 1000 * sw.N_STEPS
