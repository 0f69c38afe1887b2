"""Time-window entry points (ssm_filter_window / ssm_smooth_window / ssm_scores_phase{1,2}_window) and the
time-streaming Monte-Carlo driver built on them: walking the windows must reproduce the one-pass results bit for
bit, failures included.  Needs a B200."""
import numpy as np
import pytest
import torch

from conftest import golden, rel

pytestmark = pytest.mark.gpu


def T(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def eq(a, b):
    """bitwise equality, NaNs at the same places"""
    return torch.equal(torch.nan_to_num(a, nan=-1.2345e300), torch.nan_to_num(b, nan=-1.2345e300)) and \
        torch.equal(torch.isnan(a), torch.isnan(b))


def _sim(g, M, N, seed=3):
    from ssmtoybox_b200 import device as dv
    low = dv.lower(g)
    rng = dv.make_rng({'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
                       'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}, seed=seed)
    x, y = dv.simulate(low, M, N, rng=rng, mode='continuous', dt=0.05, sub=2)
    return low, x, y


def _windowed_forward(dv, low, y, wins, store_pred=True):
    o = {}
    for c, (a, b) in enumerate(wins):
        dv.filter_forward(low, y, store_pred=store_pred, out=o, window=(a, b), want_last=True,
                          init_mean=o['last_mean'] if c else None, init_cov=o['last_cov'] if c else None)
    return o


@pytest.mark.parametrize('M,N,wins', [
    (300, 40, [(0, 13), (13, 30), (30, 40)]),                    # plain launches
    (300, 7, [(0, 1), (1, 2), (2, 6), (6, 7)]),                  # one-step windows, windows touching slots N-1, N-2
    (60000, 64, [(0, 32), (32, 64)]),                            # more CTAs than fit: ticket scheduler inside each window
])
def test_windows_equal_one_pass_bitwise(M, N, wins):
    from ssmtoybox_b200 import device as dv
    g = golden('c3_reentry_gpq')
    low, x, y = _sim(g, M, N)
    # make a few trajectories fail at different steps: a wild measurement drives the covariance indefinite
    bad = [1, 77, M - 1]
    for i, t in enumerate(bad):
        y[:, min(N - 2, 3 + 9 * i), t] = float('nan')
    ref = dv.filter_forward(low, y, store_pred=True, want_last=True)
    st = ref['status'].cpu().numpy()
    assert (st[bad] != 0).all() and (np.delete(st, bad) == 0).all()
    o = _windowed_forward(dv, low, y, wins)
    for k in ('fi_mean', 'fi_cov', 'pr_mean', 'pr_cov', 'pr_xx_cov', 'last_mean', 'last_cov'):
        assert eq(o[k], ref[k]), k
    assert torch.equal(o['status'], ref['status'])
    # smoother (+ in-kernel statistics), windows walked backwards
    sref = dv.smooth_backward(low.dx, ref, x_truth=x)
    sm = {}
    for a, b in reversed(wins):
        dv.smooth_backward(low.dx, o, out=sm, x_truth=x, window=(a, b))
    for k in ('sm_mean', 'sm_cov', 'stats', 'rmse_acc'):
        assert eq(sm[k], sref[k]), k
    assert torch.equal(sm['status'], sref['status'])
    # plain smoother windows (no statistics)
    sm2 = {}
    for a, b in reversed(wins):
        dv.smooth_backward(low.dx, o, out=sm2, window=(a, b))
    assert eq(sm2['sm_mean'], sref['sm_mean']) and eq(sm2['sm_cov'], sref['sm_cov'])
    # score phases
    s1, acc1 = dv.scores_phase1(x, ref['fi_mean'], ref['fi_cov'], ref['status'])
    W = s1.shape[1]
    s2, acc2 = torch.empty((N, W), dtype=torch.float64, device='cuda'), torch.empty((low.dx, M), dtype=torch.float64, device='cuda')
    for a, b in wins:
        dv.scores_phase1(x, ref['fi_mean'], ref['fi_cov'], ref['status'], window=(a, b), out=(s2, acc2))
    assert eq(s2, s1) and eq(acc2, acc1)
    mse = (s1[:, low.dx:low.dx + low.dx ** 2] / s1[:, -1:]).T.reshape(low.dx, low.dx, N).contiguous()
    l1 = dv.scores_phase2(x, ref['fi_mean'], ref['fi_cov'], mse, ref['status'])
    l2 = torch.empty((N, 2), dtype=torch.float64, device='cuda')
    for a, b in wins:
        dv.scores_phase2(x, ref['fi_mean'], ref['fi_cov'], mse, ref['status'], window=(a, b), out=l2)
    assert eq(l2, l1)


def test_window_argument_validation():
    from ssmtoybox_b200 import device as dv
    g = golden('c3_reentry_gpq')
    low, x, y = _sim(g, 64, 10)
    for w in ((-1, 5), (5, 4), (0, 11)):
        with pytest.raises(ValueError):
            dv.filter_forward(low, y, window=w)
    o = dv.filter_forward(low, y, window=(3, 3))      # empty window: nothing to do
    torch.cuda.synchronize()


@pytest.mark.parametrize('smooth', [True, False])
def test_time_streaming_driver_matches_device_evaluation(smooth):
    """mc.filter_scores on HOST arrays (time-window pipeline: y forward, x backward, smoother chasing the copies) ==
    forward_pass / backward_pass / evaluate_performance on device arrays == the trajectory-chunked fallback."""
    from ssmtoybox_b200 import mc, utils as U, device as dv
    import bench
    alg, g = bench.build_filter()
    M, N = 5000, 60
    low, x, y = _sim(g, M, N, seed=11)
    xh = torch.empty(x.shape, dtype=torch.float64).pin_memory().copy_(x)
    yh = torch.empty(y.shape, dtype=torch.float64).pin_memory().copy_(y)
    r1 = mc.filter_scores(alg, yh, xh, smooth=smooth, n_windows=7)
    fwd = dv.filter_forward(low, y, store_pred=True)
    if smooth:
        sm = dv.smooth_backward(low.dx, fwd)
        mean, cov, st = sm['sm_mean'], sm['sm_cov'], sm['status']
    else:
        mean, cov, st = fwd['fi_mean'], fwd['fi_cov'], fwd['status']
    r2 = U.evaluate_performance(x, mean, cov, status=st)
    r3 = mc.filter_scores(alg, yh, xh, smooth=smooth, n_chunks=3)
    for r in (r2, r3):
        assert rel(r1['rmse'], r['rmse']) < 1e-12
        assert abs(r1['nll'] - r['nll']) < 1e-11 * abs(r['nll'])
        assert abs(r1['nci'] - r['nci']) < 1e-10 * abs(r['nci']) + 1e-12
        assert rel(r1['mse'], r['mse']) < 1e-12
    assert (r1['status'] == 0).all()
    # numpy in, device in: same numbers
    r4 = mc.filter_scores(alg, yh.numpy(), xh.numpy(), smooth=smooth, n_windows=3)
    r5 = mc.filter_scores(alg, y, x, smooth=smooth, n_windows=4)
    for r in (r4, r5):
        assert rel(r1['rmse'], r['rmse']) < 1e-12 and abs(r1['nci'] - r['nci']) < 1e-10 * abs(r['nci']) + 1e-12


def test_time_streaming_driver_with_failures_falls_back_to_exact_scores():
    """Trajectories that fail after they have contributed to the rows of an earlier window are excluded from every
    row, like a one-pass evaluation (and like the reference, whose failing runs never reach the score loops)."""
    from ssmtoybox_b200 import mc, utils as U, device as dv
    import bench
    alg, g = bench.build_filter()
    M, N = 2000, 40
    low, x, y = _sim(g, M, N, seed=5)
    y[:, 25, 17] = float('nan')
    y[:, 3, 900] = float('nan')
    r1 = mc.filter_scores(alg, y.cpu().numpy(), x.cpu().numpy(), smooth=False, n_windows=4)
    fwd = dv.filter_forward(low, y)
    r2 = U.evaluate_performance(x, fwd['fi_mean'], fwd['fi_cov'], status=fwd['status'])
    assert (r1['status'] != 0).sum() == 2 and r1['n_ok'] == M - 2
    assert rel(r1['rmse'], r2['rmse']) < 1e-12 and abs(r1['nci'] - r2['nci']) < 1e-10 * abs(r2['nci']) + 1e-12
    assert abs(r1['nll'] - r2['nll']) < 1e-11 * abs(r2['nll'])


@pytest.mark.parametrize('M,N,wins', [
    (300, 40, None),
    (300, 40, [(0, 13), (13, 30), (30, 40)]),
    (300, 7, [(0, 1), (1, 2), (2, 6), (6, 7)]),        # windows touching slots N-1, N-2 (filtered values)
    (5000, 50, [(0, 25), (25, 50)]),
    (70000, 60, None),
    (70000, 110, [(0, 55), (55, 110)]),
])
@pytest.mark.parametrize('ticket,stage', [('0', None), ('0', '0'), ('0', '3'), ('0', '7'), ('1', None)])
def test_score_only_smoother_equals_stored_smoother(M, N, wins, ticket, stage, monkeypatch):
    """ssm_smooth_scores (no smoothed arrays stored, errors d = x - m_s kept for the second score phase) gives bitwise
    the statistics, quadratic forms and scores of ssm_smooth_quad + evaluate_performance on the stored arrays."""
    from ssmtoybox_b200 import device as dv, utils as U
    if ticket == '1' and M < 60000:
        pytest.skip('the ticket scheduler only engages for multi-wave launches')
    monkeypatch.setenv('SSM_SMOOTH_TICKET', ticket)   # 1: opt-in ticket scheduler over (block, time chunk) items
    if stage is not None:   # staging mode of the score-only kernel (None: the library's default; 0 plain loads, 3 inputs and
        monkeypatch.setenv('SSM_SMOOTH_STAGE', stage)   # truth staged through shared memory, 7 two components per copy)
    g = golden('c3_reentry_gpq')
    low, x, y = _sim(g, M, N)
    y[:, 3, 5] = float('nan')                                       # one failed trajectory
    fwd = dv.filter_forward(low, y, store_pred=True)
    ref = dv.smooth_backward(low.dx, fwd, x_truth=x, want_quad=True)
    want = U.evaluate_performance(x, ref['sm_mean'], ref['sm_cov'], status=ref['status'], to_host=False,
                                  phase1=(ref['stats'], ref['rmse_acc']), quad=ref['quad'])
    sc = {}
    for (a, b) in (reversed(wins) if wins else [(0, N)]):
        dv.smooth_scores(low.dx, fwd, x, out=sc, window=None if wins is None else (a, b))
    assert 'sm_mean' not in sc and 'sm_cov' not in sc
    assert torch.equal(sc['status'], ref['status'])
    ok = (ref['status'] == 0)
    assert eq(sc['stats'], ref['stats']) and eq(sc['rmse_acc'], ref['rmse_acc'])
    assert eq(sc['quad'][:, ok], ref['quad'][:, ok])          # rows of failed trajectories are never written
    assert eq(sc['dres'][:, :, ok], (x - ref['sm_mean'])[:, :, ok])
    got = U.evaluate_scored(sc, to_host=False)
    for k in ('rmse', 'nci', 'nll', 'abs_nci', 'mse', 'rmse_vs_time', 'n_ok'):
        assert eq(got[k], want[k]), k


@pytest.mark.parametrize('name,M,N', [('c1_ungm_gpq_ut', 4098, 50), ('c5_pend_gpq', 4098, 50), ('c4_ct_gpq', 2050, 40),
                                      ('c3_reentry_gpq', 130, 30)])
def test_score_only_smoother_staging_modes_agree_bitwise(name, M, N, monkeypatch):
    """Every staging mode of the score-only smoother (SSM_SMOOTH_STAGE: 0 plain loads, 1 / 2 / 3 inputs / truth / both
    staged through shared memory, 7 lane pairs copying two components per instruction) moves the same numbers into the
    same arithmetic: statistics, quadratic forms and errors are bitwise equal for state dimensions 1, 2 and 5, with a
    last CTA that is only partly filled and a failed trajectory next to its pair partner."""
    from ssmtoybox_b200 import device as dv
    g = golden(name)
    if name == 'c3_reentry_gpq':
        low, x, y = _sim(g, M, N)
    else:
        low = dv.lower(g)
        x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=11))
    y[:, 2, 7] = float('nan')
    fwd = dv.filter_forward(low, y, store_pred=True)
    assert int((fwd['status'] == 0).sum()) > M // 2
    outs = {}
    for mode in ('0', '1', '2', '3', '7'):
        monkeypatch.setenv('SSM_SMOOTH_STAGE', mode)
        outs[mode] = dv.smooth_scores(low.dx, fwd, x)
    torch.cuda.synchronize()
    ok = outs['0']['status'] == 0
    assert int((~ok).sum()) >= 1
    for mode in ('1', '2', '3', '7'):
        assert torch.equal(outs[mode]['status'], outs['0']['status']), mode
        assert eq(outs[mode]['stats'], outs['0']['stats']) and eq(outs[mode]['rmse_acc'], outs['0']['rmse_acc']), mode
        assert eq(outs[mode]['quad'][:, ok], outs['0']['quad'][:, ok]) and eq(outs[mode]['dres'][:, :, ok], outs['0']['dres'][:, :, ok]), mode


def test_streaming_driver_score_only_equals_keep():
    """mc.filter_scores(keep=False) -- score-only smoother, statistics rows of several windows all-reduced together --
    returns the scores of the keep=True pipeline."""
    from ssmtoybox_b200 import mc
    import bench
    alg, g = bench.build_filter()
    _, x, y = _sim(g, 4000, 60)
    a = mc.filter_scores(alg, y.cpu().numpy(), x.cpu().numpy(), smooth=True, n_windows=6, keep=True)
    for every in (1, 4):
        b = mc.filter_scores(alg, y.cpu().numpy(), x.cpu().numpy(), smooth=True, n_windows=6, keep=False, reduce_every=every)
        for k in ('rmse', 'nci', 'nll', 'mse'):
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
        assert np.array_equal(a['status'], b['status'])


def test_lower_triangle_only_forward_pass_feeds_the_score_only_smoother():
    """ssm_filter_window_lower need only write the entries (row, column <= row) of fi_cov / pr_cov: those are bit for bit
    the entries of the full pass, the compact-sum instantiation (structured weights) never touches the others, and the
    score-only smoother -- which reads nothing else -- gives bitwise the same statistics from either."""
    from ssmtoybox_b200 import device as dv, utils as U
    g = dv.own_weights(golden('c3_reentry_gpq'))
    low = dv.lower(g)
    assert dv.weights_reflective(low) == (True, True)
    truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
             'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
    x, y = dv.simulate(low, 1000, 60, rng=dv.make_rng(truth, seed=5), mode='continuous', dt=0.05, sub=2)
    full = dv.filter_forward(low, y, store_pred=True)
    sentinel = -12345.0
    out = {k: torch.full_like(full[k], sentinel) for k in ('fi_cov', 'pr_cov')}
    low_o = dv.filter_forward(low, y, store_pred=True, out=out, lower_only=True)
    for k in ('fi_mean', 'pr_mean', 'pr_xx_cov', 'status'):
        assert torch.equal(low_o[k], full[k]), k
    for k in ('fi_cov', 'pr_cov'):
        for r in range(5):
            for c in range(5):
                if c <= r:
                    assert torch.equal(low_o[k][r, c], full[k][r, c]), (k, r, c)
                else:
                    assert bool((low_o[k][r, c] == sentinel).all()), (k, r, c)
    a, b = dv.smooth_scores(low.dx, full, x), dv.smooth_scores(low.dx, low_o, x)
    for k in ('stats', 'rmse_acc', 'quad', 'dres', 'status'):
        assert torch.equal(a[k], b[k]), k
    # windows: the carried state travels through last_mean / last_cov, which stay full matrices
    w = {}
    for c, (k0, k1) in enumerate(((0, 25), (25, 60))):
        dv.filter_forward(low, y, store_pred=True, out=w, window=(k0, k1), want_last=True, lower_only=True,
                          init_mean=w['last_mean'] if c else None, init_cov=w['last_cov'] if c else None)
    assert torch.equal(w['fi_mean'], full['fi_mean']) and torch.equal(w['fi_cov'][2, 1], full['fi_cov'][2, 1])
    # dense-sum instantiation (the reference's weights): same contract, full matrices written
    gd = golden('c3_reentry_gpq')
    lowd = dv.lower(gd)
    fd, ld_ = dv.filter_forward(lowd, y, store_pred=True), dv.filter_forward(lowd, y, store_pred=True, lower_only=True)
    assert torch.equal(torch.tril(fd['fi_cov'].permute(2, 3, 0, 1)), torch.tril(ld_['fi_cov'].permute(2, 3, 0, 1)))
    sa, sb = dv.smooth_scores(lowd.dx, fd, x), dv.smooth_scores(lowd.dx, ld_, x)
    assert torch.equal(sa['stats'], sb['stats'])


def test_smoother_rejects_non_finite_inputs_like_the_reference():
    """scipy's cho_factor / cho_solve raise ValueError on NaN / infinite input (ssinf.py:342): status = failing step | code 4,
    in the stored-moments and the score-only smoother alike.  The kernel tests the SUM of a step's inputs and falls back to
    the element-wise test only when that sum is not finite -- so huge FINITE inputs whose sum overflows must not be flagged
    (the reference goes on with the infinities they produce)."""
    import ssm_oracle as so
    from ssmtoybox_b200 import device as dv
    g = golden('c3_reentry_gpq')
    low = dv.lower(g)
    truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
             'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
    x, y = dv.simulate(low, 256, 40, rng=dv.make_rng(truth, seed=9), mode='continuous', dt=0.05, sub=2)
    fwd = dv.filter_forward(low, y, store_pred=True)
    assert int((fwd['status'] != 0).sum()) == 0
    fwd['pr_xx_cov'][1, 2, 11, 3] = float('nan')
    fwd['pr_cov'][2, 1, 21, 5] = float('inf')
    fwd['pr_cov'][0, 0, 30, 6] = float('-inf')
    fwd['pr_xx_cov'][:, :, 16, 7] = 1e308          # finite, but the 25 of them sum to infinity
    want = np.zeros(256, dtype=np.int64)
    want[3], want[5], want[6] = (11 << 8) | so.FAIL_NONFINITE_GAIN, (21 << 8) | so.FAIL_NONFINITE_GAIN, (30 << 8) | so.FAIL_NONFINITE_GAIN
    for run in (lambda f: dv.smooth_backward(low.dx, f), lambda f: dv.smooth_scores(low.dx, f, x),
                lambda f: dv.smooth_backward(low.dx, f, x_truth=x)):
        f = dict(fwd)
        f['status'] = fwd['status'].clone()
        st = run(f)['status'].cpu().numpy()
        assert np.array_equal(st, want), (st[[3, 5, 6, 7]], want[[3, 5, 6, 7]])
    sm = dv.smooth_backward(low.dx, dict(fwd, status=fwd['status'].clone()))
    m = sm['sm_mean'].cpu().numpy()
    assert np.isfinite(m[:, 16:, 7]).all() and not np.isfinite(m[:, :15, 7]).any()     # overflow goes on as inf / NaN, unflagged
    assert np.isnan(m[:, :11, 3]).all() and np.isfinite(m[:, 11:, 3]).all()            # NaN rows from the failing step down
