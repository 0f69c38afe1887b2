"""Developer tool: A/B of the two forward-pass mappings (one thread per trajectory vs warp pair) on one GPU.

    python tools/ab_pair.py [M] [golden case]

Runs both kernels on the same simulated measurements, compares every output array and the failure status, and times
them (filter only / with predictive moments).  SSM_PAIR is read by the library at every launch.
"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv


def run(low, y, sp, pair, reps=5):
    os.environ['SSM_PAIR'] = '1' if pair else '0'
    o = {}
    for _ in range(3):
        dv.filter_forward(low, y, store_pred=sp, out=o)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.filter_forward(low, y, store_pred=sp, out=o); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return o, float(np.median(ts)), float(np.min(ts))


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
    name = sys.argv[2] if len(sys.argv) > 2 else 'c3_reentry_gpq'
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', name + '.npz')))
    low = dv.lower(g)
    if 'reentry' in name:
        truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
        x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
    else:
        x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=1))
    for sp in (False, True):
        res = {}
        for pair in (False, True):
            o, ms, mn = run(low, y, sp, pair)
            res[pair] = {k: v.clone() for k, v in o.items() if torch.is_tensor(v)}
            print('%s M=%d N=%d store_pred=%s pair=%d: %.2f ms (min %.2f)  %.3e traj-steps/s  fails=%d' % (name, M, N, sp, pair, ms, mn, M * N / ms * 1e3, int((o['status'] != 0).sum())), flush=True)
            del o
        a, b = res[False], res[True]
        same_status = bool((a['status'] == b['status']).all())
        print('  status equal:', same_status, ' mismatches:', int((a['status'] != b['status']).sum()))
        okm = (a['status'] == 0) & (b['status'] == 0)
        for k in a:
            if k == 'status' or a[k].dtype != torch.float64:
                continue
            u, v = a[k][..., okm], b[k][..., okm]
            scale = u.abs().amax(dim=tuple(range(u.dim() - 1)), keepdim=True).clamp_min(1e-300) if u.dim() > 1 else u.abs().max()
            err = ((u - v).abs() / scale)
            print('  %-10s max rel (per-trajectory max-norm) %.3e   nan mismatch %d' % (k, float(torch.nan_to_num(err).max()), int((torch.isnan(u) != torch.isnan(v)).sum())), flush=True)
        del res, a, b
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
