import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from conftest import golden, relstep
from ssmtoybox_b200.research import bsq_tracking, icinco_demo
g = golden('research_bsq_reentry_demo')
out = bsq_tracking.reentry_demo(dur=float(g['dur']), x=g['x'], y=g['y'], keep_arrays=True)
for a in range(4):
    print(a, 'mean relstep %.2e cov relstep %.2e' % (relstep(out['mean'][a].cpu().numpy(), g['mean'][..., a]), relstep(out['cov'][a].cpu().numpy(), g['cov'][..., a])))
    for part in ('state', 'position', 'velocity', 'parameter'):
        print('   ', part, 'rmse rel %.2e  inc abs %.2e' % (np.max(np.abs(out[part]['rmse'][:, a] / g[part + '_rmse'][:, a] - 1)), np.max(np.abs(out[part]['inc'][:, a] - g[part + '_inc'][:, a]))))
g = golden('research_icinco_hypers')
o = icinco_demo.hypers_demo(lscale=list(g['lscale']), x=g['x'], z=g['z'], carry_over=True)
for k, r in (('rmse', 'rmse'), ('nci', 'nci'), ('neg_log_likelihood', 'nll')):
    print(k, np.abs(o[k] / g[r] - 1).ravel())
