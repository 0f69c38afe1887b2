"""Developer tool: the tpq_ungm driver at a small size; failures of the fully-symmetric Student filter against the oracle."""
import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np, torch
from ssmtoybox_b200 import utils as U
from ssmtoybox_b200.research import tpq_ungm
U.seed(5)
o = tpq_ungm.ungm_demo(steps=60, mc_sims=400, mc_weight_samples=200000, num_bs_samples=2000)
print(o['n_failed']); print(o['table'])
from ssmtoybox_b200.ssinf import FullySymmetricStudent
from ssmtoybox_b200.ssmod import UNGMTransition, UNGMMeasurement
from ssmtoybox_b200.utils import StudentRV
nu=4.0
dyn = UNGMTransition(StudentRV(1, scale=(nu - 2) / nu * 1.0, dof=nu), StudentRV(1, scale=((nu - 2) / nu) * 10.0, dof=nu))
obs = UNGMMeasurement(StudentRV(1, scale=((nu - 2) / nu) * 0.01, dof=nu), 1)
f = FullySymmetricStudent(dyn, obs, kappa=0.0, dof=4.0)
m, P = f.forward_pass(o['z'])
st = np.asarray(f.status.cpu() if hasattr(f.status,'cpu') else f.status)
print('failed', (st!=0).sum(), np.unique(st & 0xff, return_counts=True), np.unique(st>>8)[:10])
import ssm_oracle as so
d = f._describe()
ref = so.student_forward_pass(d, o['z'].cpu().numpy(), backend='loops')
print('oracle failed', (ref['status']!=0).sum(), np.unique(ref['status'] & 0xff, return_counts=True), np.nonzero(st)[0], np.nonzero(ref['status'])[0], st[st != 0], ref['status'][ref['status'] != 0])
mg = m if isinstance(m, np.ndarray) else m.cpu().numpy()
Pg = P if isinstance(P, np.ndarray) else P.cpu().numpy()
ok = st == 0
num = np.abs(mg[0][:, ok] - ref['fi_mean'][0][:, ok]); den = np.maximum(np.abs(ref['fi_mean'][0][:, ok]), 1e-3)
print('max rel diff GPU vs oracle (ok trajectories): %.3e; at step-wise median %.3e' % ((num / den).max(), np.median(num / den)))
xx = o['x'].cpu().numpy()
print('oracle RMSE (time-avg of mean |err|):', np.abs(xx[0] - ref['fi_mean'][0]).mean(), ' GPU:', np.nanmean(np.abs(xx[0] - mg[0])))
i = int(np.nonzero(st)[0][0])
print('traj', i, 'gpu mean', mg[0, :11, i], '\n oracle mean', ref['fi_mean'][0, :11, i], '\n gpu cov', Pg[0, 0, :11, i], '\n oracle cov', ref['fi_cov'][0, 0, :11, i])
