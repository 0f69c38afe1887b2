import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from ssmtoybox_b200 import device as dv
g = dict(np.load('tests/golden/c3_reentry_gpq.npz'))
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
for own in (False, True):
    gg = dv.own_weights(g) if own else g
    low = dv.lower(gg)
    x, y = dv.simulate(low, 125000, 500, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
    del x
    o = {}
    for lo in (False, True, False, True):
        for _ in range(3):
            dv.filter_forward(low, y, store_pred=True, out=o, lower_only=lo)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dv.filter_forward(low, y, store_pred=True, out=o, lower_only=lo); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('own' if own else 'reference', 'lower_only', lo, '%.2f ms' % float(np.median(ts)))
    del o, y
