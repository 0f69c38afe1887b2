"""Developer tool: time the fused forward pass (reentry GPQ, C3 share) on one GPU."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv

def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
    name = sys.argv[2] if len(sys.argv) > 2 else 'c3_reentry_gpq'
    own = name.endswith(':own')     # <case>:own = the package's own (structured) weights instead of the golden run's
    name = name.split(':')[0]
    g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', name + '.npz')))
    if own:
        g = dv.own_weights(g)
    low = dv.lower(g)
    print('compact sums (dyn, obs):', dv.weights_reflective(low))
    N = 500
    if 'reentry' in name:
        truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
        x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
    else:
        x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=1))
    for sp in (False, True):
        o = {}
        for _ in range(3):
            dv.filter_forward(low, y, store_pred=sp, out=o)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dv.filter_forward(low, y, store_pred=sp, out=o); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts)); mn = float(np.min(ts))
        print('%s M=%d N=%d store_pred=%s: %.2f ms (min %.2f)  %.3e traj-steps/s  %.2f TFLOP/s(alg 5756)  fails=%d' % (name, M, N, sp, ms, mn, M * N / ms * 1e3, M * N * 5756 / ms * 1e-9, int((o['status'] != 0).sum())))
        del o

if __name__ == '__main__':
    main()
