"""Developer tool: a few forward-pass launches at a given trajectory count (for ncu captures)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
M = int(sys.argv[1]); N = int(sys.argv[2]); name = sys.argv[3] if len(sys.argv) > 3 else 'c3_reentry_gpq'
sp = len(sys.argv) > 4 and sys.argv[4] in ('pred', 'predlow')
low_only = len(sys.argv) > 4 and sys.argv[4] == 'predlow'     # the score pipeline's forward pass: lower triangles only
own = name.endswith(':own')     # <case>:own = the package's own (structured) weights instead of the golden run's
name = name.split(':')[0]
g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', name + '.npz')))
if own:
    g = dv.own_weights(g)
low = dv.lower(g)
print('compact sums (dyn, obs):', dv.weights_reflective(low))
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
if 'reentry' in name and 'reentry1d' not in name:
    x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
else:   # the other models: discrete simulation from the golden descriptor
    x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=1))
o = {}
scored = len(sys.argv) > 4 and sys.argv[4] == 'scored'     # scoring forward pass (ssm_filter_scores): no moment arrays
for _ in range(3):
    if scored:
        dv.filter_scored(low, y, x, out=o)
    else:
        dv.filter_forward(low, y, store_pred=sp, out=o, lower_only=low_only)
torch.cuda.synchronize()
print('ok', int((o['status'] != 0).sum()))
