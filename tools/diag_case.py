"""Developer tool: failure statistics of a golden filter configuration on device-simulated data against the oracle."""
import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle')]
import numpy as np, torch
import ssm_oracle as so
from ssmtoybox_b200 import device as dv
name = sys.argv[1] if len(sys.argv) > 1 else 'c4_ct_tpq'
M, N = int(sys.argv[2]) if len(sys.argv) > 2 else 512, int(sys.argv[3]) if len(sys.argv) > 3 else 500
g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', name + '.npz')))
low = dv.lower(g)
x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=1))
o = dv.filter_forward(low, y)
st = o['status'].cpu().numpy()
ref = so.forward_pass(g, y.cpu().numpy(), backend='loops')
print('gpu failed', (st != 0).sum(), 'oracle failed', (ref['status'] != 0).sum(), 'same set', np.array_equal(st != 0, ref['status'] != 0),
      'same status', (st == ref['status']).mean())
print('gpu codes', np.unique(st & 0xff, return_counts=True), 'steps', np.percentile((st >> 8)[st != 0], [0, 25, 50, 75, 100]) if (st != 0).any() else None)
print('oracle codes', np.unique(ref['status'] & 0xff, return_counts=True))
