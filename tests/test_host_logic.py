"""Host-side logic of the facade that needs no GPU: point sets / classical weights (host constant tables),
multi-index generation, trajectory sharding and the gloo-backed reduction of the packed statistics."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ssm_oracle as so
from conftest import golden


def test_facade_pointsets_match_reference():
    from ssmtoybox_b200 import mtran
    g = golden('pointsets')
    for dim in (1, 2, 5):
        assert np.array_equal(mtran.UnscentedTransform.unit_sigma_points(dim), g['ut%d_pts' % dim])
        wm, wc = mtran.UnscentedTransform.weights(dim)
        assert np.array_equal(wm, g['ut%d_wm' % dim]) and np.array_equal(wc, g['ut%d_wc' % dim])
        wm, wc = mtran.UnscentedTransform.weights(dim, 2.0, 0.5, 1.0)
        assert np.array_equal(wm, g['ut%dk2a_wm' % dim]) and np.array_equal(wc, g['ut%dk2a_wc' % dim])
        assert np.array_equal(mtran.SphericalRadialTransform.unit_sigma_points(dim), g['sr%d_pts' % dim])
        assert np.array_equal(mtran.SphericalRadialTransform.weights(dim), g['sr%d_wm' % dim])
        for deg in (3, 5):
            assert np.array_equal(mtran.FullySymmetricStudentTransform.unit_sigma_points(dim, deg, None, 6.0), g['fs%dd%d_pts' % (dim, deg)])
            assert np.array_equal(mtran.FullySymmetricStudentTransform.weights(dim, deg, None, 6.0), g['fs%dd%d_wm' % (dim, deg)])
    for dim, deg in ((1, 3), (1, 5), (1, 20), (2, 3), (2, 5), (5, 3)):
        assert np.array_equal(mtran.GaussHermiteTransform.unit_sigma_points(dim, deg), g['gh%dd%d_pts' % (dim, deg)])
        assert np.allclose(mtran.GaussHermiteTransform.weights(dim, deg), g['gh%dd%d_wm' % (dim, deg)], rtol=1e-14)


def test_symmetric_set_shapes():
    """reference tests/test_mtran.py:64-88"""
    from ssmtoybox_b200.mtran import FullySymmetricStudentTransform as FS
    assert FS.symmetric_set(3, []).shape == (3, 1)
    assert FS.symmetric_set(3, [1.0]).shape == (3, 6)
    assert FS.symmetric_set(3, [1.0, 1.0]).shape == (3, 12)
    assert FS.symmetric_set(5, [2.0, 2.0]).shape == (5, 40)


def test_n_sum_k():
    from ssmtoybox_b200.bq.bqmod import n_sum_k
    a = n_sum_k(3, 2)
    assert a.shape == (3, 6) and (a.sum(axis=0) == 2).all()
    assert len({tuple(c) for c in a.T}) == 6
    assert np.array_equal(n_sum_k(2, 0), np.zeros((2, 1), dtype=int)) and np.array_equal(n_sum_k(2, 1), np.eye(2, dtype=int))


def test_shard_ranges_cover_everything():
    from ssmtoybox_b200.dist import shard_range
    for n, ws in ((10 ** 6, 8), (1000, 3), (5, 8), (0, 2), (125000, 1)):
        spans = [shard_range(n, r, ws) for r in range(ws)]
        assert sum(c for _, c in spans) == n
        off = 0
        for o, c in spans:
            assert o == off or c == 0
            off += c
    assert shard_range(10 ** 6, 3, 8) == (375000, 125000)


def _packed_stats_numpy(x, m, P):
    """numpy statement of the K6 phase-1 row layout [sum SE | sum dd' | sum NLL | sum |d| | count]."""
    dx, N, M = x.shape
    d = x - m
    se = (d ** 2).sum(axis=2).T
    outer = np.einsum('ikm,jkm->kij', d, d).reshape(N, dx * dx)
    nll = so.neg_log_likelihood(x, m, P).sum(axis=1)[:, None]
    nrm = np.sqrt((d ** 2).sum(axis=0)).sum(axis=1)[:, None]
    return np.hstack([se, outer, nll, nrm, np.full((N, 1), float(M))]), np.sqrt((d ** 2).mean(axis=1)).sum(axis=1)


def _worker(rank, world_size, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    from ssmtoybox_b200.dist import Communicator, finalize_scores
    comm = Communicator.from_env(backend='gloo')
    g = golden('c5_pend_gpq')
    x, m, P = g['x'], g['fi_mean'], g['fi_cov']
    dx, N, M = x.shape
    off, cnt = comm.shard(M)
    sl = slice(off, off + cnt)
    stats, rm = _packed_stats_numpy(x[..., sl], m[..., sl], P[..., sl])
    pack = torch.tensor(np.concatenate([stats.ravel(), rm]))
    comm.allreduce_sum(pack)                                    # phase 1: one collective
    st = pack[:stats.size].reshape(stats.shape).numpy()
    sc = finalize_scores(st, pack[stats.size:].numpy(), None, dx, N)
    lcr = so.log_cred_ratio(x[:, :, sl], m[:, :, sl], P[:, :, :, sl], sc['mse'])   # needs the GLOBAL mse
    l2 = torch.tensor(np.stack([lcr.sum(axis=1), np.abs(lcr).sum(axis=1)], axis=1))
    comm.allreduce_sum(l2)                                      # phase 2: one more, N x 2 doubles
    sc = finalize_scores(st, pack[stats.size:].numpy(), l2.numpy(), dx, N)
    tmax = comm.allreduce_max(float(rank + 1))
    comm.barrier()
    if rank == 0:
        ret.put({k: np.asarray(v) for k, v in sc.items()} | {'tmax': tmax, 'ws': comm.world_size})
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_reduction_matches_single_process():
    """world_size-2 gloo run of the sharded score reduction == the oracle on the whole set."""
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out['ws'] == 2 and out['tmax'] == 2.0
    g, gs = golden('c5_pend_gpq'), golden('scores')
    e = so.evaluate_performance(g['x'], g['fi_mean'], g['fi_cov'])
    assert np.allclose(out['rmse'], e['rmse'], rtol=1e-13)
    assert np.allclose(out['mse'], e['mse'], rtol=1e-12, atol=1e-300)
    assert abs(out['nll'] - e['nll']) < 1e-12 * abs(e['nll'])
    assert abs(out['nci'] - e['nci']) < 1e-9 * abs(e['nci'])
    assert abs(out['nci'] - gs['c5_pend_gpq_nci_f'].ravel()[0]) < 1e-9 * abs(e['nci'])


def test_rendezvous_batches_scipy_optimisers_without_changing_their_path():
    """MarginalInference runs one scipy BFGS per trajectory in a thread and serves all their objective requests in
    batches (ssinf._Rendezvous).  Host logic only: with a numpy objective the batched optimisers must return exactly what
    scipy.optimize.minimize returns when called directly, every request must be served in a batch of all optimisers still
    running, and optimisers that finish early must not stall the others."""
    import threading
    from scipy.optimize import minimize
    from ssmtoybox_b200.ssinf import _Rendezvous
    rs = np.random.RandomState(0)
    n = 7
    A = [np.diag(rs.uniform(0.5, 5.0, 3)) + 0.1 * np.ones((3, 3)) for _ in range(n)]
    b = [rs.randn(3) for _ in range(n)]

    def f(i, th):
        d = th - b[i]
        return 0.5 * d.dot(A[i]).dot(d) + 0.1 * np.sum(d ** 4) * (i % 3)     # different difficulty -> different iteration counts
    direct = [minimize(lambda th, i=i: f(i, th), np.zeros(3), method='BFGS') for i in range(n)]
    rv = _Rendezvous(n)
    res, batch_sizes = {}, []

    def worker(i):
        try:
            res[i] = minimize(lambda th: rv.request(i, th), np.zeros(3), method='BFGS')
        finally:
            rv.finish(i)

    def evaluate(ids, thetas):
        batch_sizes.append(len(ids))
        return [f(i, th) for i, th in zip(ids, thetas)]
    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    rv.serve(evaluate)
    for t in threads:
        t.join(timeout=30)
        assert not t.is_alive()
    for i in range(n):
        assert np.array_equal(res[i].x, direct[i].x) and np.array_equal(res[i].hess_inv, direct[i].hess_inv)
        assert res[i].nfev == direct[i].nfev
    assert batch_sizes[0] == n and min(batch_sizes) >= 1 and sorted(batch_sizes, reverse=True) == batch_sizes
    assert len(batch_sizes) == max(r.nfev for r in direct)


def test_reflection_symmetric_weights_projection_and_predicate():
    """symmetrize_reflective projects a weight set on [0 | cI | -cI] points onto the structure the formulas have in exact
    arithmetic (bq/bqmod.py:495-523): afterwards ssm_weights_reflective (the host check that selects the compact sums of
    the forward pass) accepts it; weights that carry rounding noise, or no structure at all, are refused / left alone."""
    import ctypes as C
    from conftest import golden, rel
    from ssmtoybox_b200 import device as dv, _lib
    from ssmtoybox_b200.bq import bqmod
    g = golden('c3_reentry_gpq')
    pred = lambda low: (_lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_dyn)), _lib.lib.ssm_weights_reflective(C.byref(low.desc.tf_obs)))  # noqa: E731
    assert pred(dv.lower(g)) == (0, 0)                     # the reference's weights: 1e-12 ... O(1) noise where zeros belong
    w = dict(wm=g['dyn_wm'], Wc=g['dyn_Wc'], Wcc=g['dyn_Wcc'], iK=None)
    s = bqmod.symmetrize_reflective(g['dyn_points'], w)
    assert s is not w
    for k in ('wm', 'Wc', 'Wcc'):
        assert rel(s[k], w[k]) < 2e-7, k                   # the dynamics kernel is well conditioned: structure up to 1e-7
    s2 = bqmod.symmetrize_reflective(g['dyn_points'], s)
    assert all(np.array_equal(s2[k], s[k]) for k in ('wm', 'Wc', 'Wcc'))          # a projection
    D = 5
    assert np.count_nonzero(s['Wcc']) == 2 * D and np.array_equal(s['Wcc'][:, 1:1 + D], -s['Wcc'][:, 1 + D:])
    # the measurement kernel (cond 1e9): the reference's float64 Wc is O(1) noise, no structure to project onto
    wo = dict(wm=g['obs_wm'], Wc=g['obs_Wc'], Wcc=g['obs_Wcc'], iK=None)
    assert bqmod.symmetrize_reflective(g['obs_points'], wo) is wo
    g2 = dict(g)
    for k in ('wm', 'Wc', 'Wcc'):
        g2['dyn_' + k] = s[k]
        g2['obs_' + k] = s[k]        # any exactly structured set will do for the predicate
    assert pred(dv.lower(g2)) == (1, 1)
    g3 = dict(g2)
    g3['dyn_Wcc'] = s['Wcc'].copy()
    g3['dyn_Wcc'][0, 3] = 1e-300                           # one entry that should be zero is not
    assert pred(dv.lower(g3)) == (0, 1)
    g4 = dict(g2)
    g4['obs_wm'] = s['wm'].copy()
    g4['obs_wm'][2] = np.nextafter(s['wm'][2], 1.0)        # one ulp off its mirror image
    assert pred(dv.lower(g4)) == (1, 0)
    # other point sets are left alone
    gh = golden('c3_reentry_ghkf3')
    wg = dict(wm=gh['dyn_wm'], Wc=np.diag(gh['dyn_Wc']) if gh['dyn_Wc'].ndim == 1 else gh['dyn_Wc'], Wcc=np.zeros((5, gh['dyn_points'].shape[1])), iK=None)
    assert bqmod.symmetrize_reflective(gh['dyn_points'], wg) is wg


@pytest.mark.parametrize('dim,kind', [(1, 'gp'), (2, 'gp'), (5, 'gp'), (2, 'bs'), (5, 'bs')])
def test_weight_formulas_have_the_reflection_structure(dim, kind):
    """The oracle's float64 restatement of GaussianProcessModel.bq_weights (bq/bqmod.py:495-523) and
    BayesSardModel.bq_weights (:893-992) on UT points with well-conditioned kernels: the computed weights are within
    rounding of the reflection structure (so the projection accepts them and moves them by rounding noise only) -- for the
    polynomial-mean model too, whose basis 1, x_j, x_j^2 is closed under coordinate reflections."""
    import ssm_oracle as so
    from ssmtoybox_b200.bq import bqmod
    from ssmtoybox_b200.mtran import UnscentedTransform
    pts = UnscentedTransform.unit_sigma_points(dim)
    assert bqmod.reflective_axis_set(pts) is not None
    par = np.array([[1.0] + [1.5 + 0.5 * d for d in range(dim)]])       # distinct length-scales: no permutation symmetry
    if kind == 'gp':
        w = so.gp_weights(par, pts)
    else:
        # the basis of the reference's BSQ set-ups: 1, x_j, x_j^2 (as many functions as UT points; tests/test_ssinf.py:195-203)
        w = so.bs_weights(par, pts, np.hstack((np.zeros((dim, 1)), np.eye(dim), 2 * np.eye(dim))).astype(int))
    w = dict(wm=np.asarray(w['wm']), Wc=np.asarray(w['Wc']), Wcc=np.asarray(w['Wcc']), iK=None)
    s = bqmod.symmetrize_reflective(pts, w)
    assert s is not w, 'weights further than 1e-6 from the reflection structure'
    for k in ('wm', 'Wc', 'Wcc'):
        assert np.abs(s[k] - w[k]).max() <= 1e-10 * max(np.abs(w[k]).max(), 1e-300), (k, np.abs(s[k] - w[k]).max())
    assert np.count_nonzero(s['Wcc']) <= 2 * dim
