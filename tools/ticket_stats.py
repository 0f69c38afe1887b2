"""Developer tool: distribution of forward-pass times over fresh allocations (ticket scheduler stability)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
M, N = 125000, 500
g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'c3_reentry_gpq.npz')))
low = dv.lower(g)
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
res = []
junk = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    junk.append(torch.empty(int(np.random.randint(1, 64)) * 1024 * 1024, dtype=torch.uint8, device='cuda'))  # shift addresses
    o = {}
    ts = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.filter_forward(low, y, store_pred=True, out=o); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    res.append(ts[1:])
    del o
    torch.cuda.empty_cache()
print(os.environ.get('SSM_TICKET_CHUNK', 'default'), ' '.join('%.1f/%.1f/%.1f' % tuple(r) for r in res))
