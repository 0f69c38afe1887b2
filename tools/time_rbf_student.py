import sys, time
sys.path[:0] = ['.', 'oracle', 'tests']
import numpy as np, torch
from ssmtoybox_b200.bq.bqkern import RBFStudent
from conftest import golden
g = golden('c4_ct_fsstudent_tpq')
par, x = g['obs_kern_par'], g['obs_points']
for n in (2000000, 20000000):
    for rep in range(3):
        k = RBFStudent(5, par, dof=4.0, num_samples=n, seed=rep)
        torch.cuda.synchronize(); t = time.perf_counter()
        q = k.exp_x_kx(par, x)
        torch.cuda.synchronize(); print(n, 'samples: %.3f ms' % ((time.perf_counter() - t) * 1e3))
