import sys
sys.path[:0] = ['.', 'oracle', 'tests']
import numpy as np, torch
import ssm_oracle as so
from conftest import golden
from ssmtoybox_b200 import device as dv
g = golden('c10_ctrs_fixture_ukf')
low = dv.lower(g)
o = dv.filter_forward(low, torch.as_tensor(g['y'], device='cuda'), store_pred=True)
fm = o['fi_mean'].cpu().numpy(); pm = o['pr_mean'].cpu().numpy()
np.set_printoptions(linewidth=200, precision=6)
for k in range(4):
    print('k', k, 'pr gpu', pm[:, k, 0], 'ref', g['pr_mean'][:, k + 1, 0])
    print('     fi gpu', fm[:, k, 0], 'ref', g['fi_mean'][:, k, 0])
print('pr_cov gpu', o['pr_cov'].cpu().numpy()[:, :, 0, 0]); print('ref', g['pr_cov'][:, :, 1, 0])
print('y', g['y'][:, :3, 0])
