"""Developer tool: per-component error of the device forward pass and of the reference's golden run against the
longdouble oracle on a noise-dominated golden case (run once per library variant via SSM_B200_LIB)."""
import os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'oracle')); sys.path.insert(0, os.path.join(root, 'tests'))
import ssm_oracle as so
from ssmtoybox_b200 import device as dv
name = sys.argv[1] if len(sys.argv) > 1 else 'c3_reentry_bsq'
g = dict(np.load(os.path.join(root, 'tests', 'golden', name + '.npz')))
ld = so.forward_pass(g, g['y'], backend='loops', dtype=np.longdouble)
low = dv.lower(g)
o = dv.filter_forward(low, torch.as_tensor(g['y'], device='cuda'), store_pred=True)
fm = o['fi_mean'].cpu().numpy(); fc = o['fi_cov'].cpu().numpy()
t = np.asarray(ld['fi_mean'], dtype=np.float64); tc = np.asarray(ld['fi_cov'], dtype=np.float64)
N = t.shape[1]
for K in (10, 30, 60, 100, N):
    K = min(K, N)
    eg = np.abs(fm[:, :K] - t[:, :K]).max(axis=(1, 2)); er = np.abs(g['fi_mean'][:, :K] - t[:, :K]).max(axis=(1, 2))
    cg = np.abs(fc[:, :, :K] - tc[:, :, :K]).max(axis=(2, 3)); cr = np.abs(g['fi_cov'][:, :, :K] - tc[:, :, :K]).max(axis=(2, 3))
    print('steps < %3d  mean err device %s\n             mean err refer. %s\n             cov diag err device %s\n             cov diag err refer. %s' % (
        K, np.array2string(eg, precision=2), np.array2string(er, precision=2), np.array2string(np.diag(cg), precision=2), np.array2string(np.diag(cr), precision=2)))
x = g['x']
print('rmse device', np.sqrt(((fm - x) ** 2).mean(axis=1)).T)
print('rmse refer.', np.sqrt(((g['fi_mean'] - x) ** 2).mean(axis=1)).T)
print('rmse ldbl  ', np.sqrt(((t - x) ** 2).mean(axis=1)).T)
