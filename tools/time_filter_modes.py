"""Developer tool: forward-pass time by set of stored outputs (separates store cost from cross-covariance compute)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
M, N = 125000, 500
g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'c3_reentry_gpq.npz')))
low = dv.lower(g)
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
kw = dict(dtype=torch.float64, device='cuda')
full = dv.filter_forward(low, y, store_pred=True)
for name, keys in (('fi only', ('fi_mean', 'fi_cov')), ('fi_mean only', ('fi_mean',)), ('fi + pr_mean + pr_cov', ('fi_mean', 'fi_cov', 'pr_mean', 'pr_cov')),
                   ('fi + pr_xx', ('fi_mean', 'fi_cov', 'pr_xx_cov')), ('all', ('fi_mean', 'fi_cov', 'pr_mean', 'pr_cov', 'pr_xx_cov'))):
    o = {k: full[k] for k in keys}
    o['status'] = full['status']
    ts = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.filter_forward(low, y, store_pred=False, store_cov='fi_cov' in keys, out=o); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('%-24s %.2f ms' % (name, float(np.median(ts[1:]))))
