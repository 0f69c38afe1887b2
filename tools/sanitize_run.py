"""Developer tool: one pass over every kernel family at small sizes, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck): ragged trajectory counts (tails of the 128-thread CTAs), the ticket scheduler
(multi-wave launch), time windows, in-kernel scores, both score phases, bootstrap, simulators, BQ weights."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
G = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden')
big = len(sys.argv) > 1 and sys.argv[1] == 'ticket'
for name, M, N in (('c3_reentry_gpq', 60000 if big else 333, 50 if big else 12), ('c1_ungm_ukf', 1000, 20), ('c5_pend_gpq', 515, 15),
                   ('c6_reentry1d_gpq', 130, 15), ('c7_ungmna_ukf', 257, 15), ('c4_ct_fsstudent', 200, 10)):
    g = dict(np.load(os.path.join(G, name + '.npz')))
    low = dv.lower(g)
    x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=3))
    student = 'dof' in g
    fwd = dv.filter_forward(low, y, store_pred=not student)
    if not student:
        sm = dv.smooth_backward(low.dx, fwd, x_truth=x)
        h = N // 2
        o2 = {}
        dv.smooth_backward(low.dx, fwd, out=o2, x_truth=x, window=(h, N))
        dv.smooth_backward(low.dx, fwd, out=o2, x_truth=x, window=(0, h))
        assert torch.equal(o2['sm_mean'], sm['sm_mean'])
    st, acc = dv.scores_phase1(x, fwd['fi_mean'], fwd['fi_cov'], fwd['status'], nll_acc=torch.zeros(M, dtype=torch.float64, device='cuda'))
    dx = low.dx
    mse = (st[:, dx:dx + dx * dx] / st[:, -1:]).T.reshape(dx, dx, N).contiguous()
    dv.scores_phase2(x, fwd['fi_mean'], fwd['fi_cov'], mse, fwd['status'], lcr_acc=torch.zeros(M, dtype=torch.float64, device='cuda'))
    dv.bootstrap_var(acc[0], 200)
    torch.cuda.synchronize()
    print(name, 'ok', int((fwd['status'] != 0).sum()), 'failed of', M, flush=True)
w = dv.bq_weights(np.array([[1.0, 3.0, 2.0], [1.0, 0.5, 0.7]]), np.array([[0, 1.7, 0, -1.7, 0], [0, 0, 1.7, 0, -1.7]], dtype=float))
print('weights ok', w['wm'].shape)
