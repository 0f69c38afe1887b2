"""Developer tool: distribution of (device error) / (reference error) against the longdouble oracle, cumulative maxima per
(step, trajectory), for the noise-dominated golden cases."""
import os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'oracle')); sys.path.insert(0, os.path.join(root, 'tests'))
import ssm_oracle as so
from ssmtoybox_b200 import device as dv

def cum(a, b):
    ax = tuple(range(a.ndim - 2))
    e = np.abs(a - b).max(axis=ax) / np.maximum(np.abs(b).max(axis=ax), 1e-300)
    return np.maximum.accumulate(np.nan_to_num(e, nan=np.inf), axis=0)

for name in sys.argv[1:] or ['c3_reentry_bsq', 'c4_ct_bsq', 'c5_pend_bsq']:
    g = dict(np.load(os.path.join(root, 'tests', 'golden', name + '.npz')))
    ld = so.forward_pass(g, g['y'], backend='loops', dtype=np.longdouble)
    o = dv.filter_forward(dv.lower(g), torch.as_tensor(g['y'], device='cuda'), store_pred=True)
    for key in ('fi_mean', 'fi_cov'):
        t = np.asarray(ld[key], dtype=np.float64)
        er, eg = cum(g[key], t), cum(o[key].cpu().numpy(), t)
        v = er < 1e-2
        q = eg[v] / np.maximum(er[v], 1e-16)
        print(name, key, 'valid %d/%d' % (v.sum(), v.size), 'ratio quantiles 50/90/99/max: %.2f %.2f %.2f %.2f' % tuple(np.quantile(q, [0.5, 0.9, 0.99, 1.0])),
              ' log-mean %.2f' % np.exp(np.mean(np.log(np.maximum(q, 1e-3)))))
