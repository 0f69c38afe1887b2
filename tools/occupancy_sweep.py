"""Developer tool: forward-pass time vs resident warps per SM (single wave), to tell latency-bound from pipe-bound."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'c3_reentry_gpq.npz')))
low = dv.lower(g)
N = 200
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
for wpsm in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 12, 16]:
    M = 148 * 32 * wpsm
    x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
    o = {}
    for _ in range(2):
        dv.filter_forward(low, y, store_pred=False, out=o)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.filter_forward(low, y, store_pred=False, out=o); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print('warps/SM %2d  M=%6d: %.3f ms  cycles per warp-step (1.965 GHz) %.0f   %.3e traj-steps/s' % (wpsm, M, ms, ms * 1e-3 * 1.965e9 / N, M * N / ms * 1e3))
