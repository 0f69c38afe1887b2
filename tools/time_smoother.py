"""Developer tool: RTS smoother (+ in-kernel scores) and score-phase timings at the bench size; compares the
smoothed moments of the TMA path with the per-thread ld/st kernel (SSM_SMOOTH_TMA=0 in a second process)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
M = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'c3_reentry_gpq.npz')))
low = dv.lower(g)
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
fwd = dv.filter_forward(low, y, store_pred=True)
def t(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), float(np.median(ts[1:]))
sm, sm2 = {}, {}
a = t(lambda: dv.smooth_backward(low.dx, fwd, out=sm))
b = t(lambda: dv.smooth_backward(low.dx, fwd, out=sm2, x_truth=x))
s1 = t(lambda: dv.scores_phase1(x, sm['sm_mean'], sm['sm_cov'], sm['status']))
st, acc = dv.scores_phase1(x, sm['sm_mean'], sm['sm_cov'], sm['status'])
mse = (st[:, 5:30] / st[:, -1:]).T.reshape(5, 5, N).contiguous()
s2 = t(lambda: dv.scores_phase2(x, sm['sm_mean'], sm['sm_cov'], mse, sm['status']))
print('TMA=%s PF=%s MINB=%s M=%d N=%d  smoother %.2f/%.2f ms  smoother+scores %.2f/%.2f ms  phase1 %.2f/%.2f  phase2 %.2f/%.2f  (min/median)' % (
    os.environ.get('SSM_SMOOTH_TMA', '0'), os.environ.get('SSM_SMOOTH_PF', '1'), os.environ.get('SSM_SMOOTH_MINB', '2'), M, N, *a, *b, *s1, *s2))
print('fails', int((sm['status'] != 0).sum()), 'checksum %.17g %.17g' % (sm['sm_mean'].double().sum().item(), sm['sm_cov'].double().sum().item()),
      'stats %.17g' % sm2['stats'].sum().item(), 'eq', torch.equal(sm['sm_mean'], sm2['sm_mean']), torch.equal(sm['sm_cov'], sm2['sm_cov']),
      'stats vs phase1 rel %.2e' % ((sm2['stats'] - st).abs().max() / st.abs().max()).item())
