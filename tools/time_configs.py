"""Developer tool: forward pass (+ RTS smoother) throughput of every BASELINE.json configuration on one GPU, with both
roofline fractions (algorithmic FLOP and bytes per trajectory-step from SURVEY.md section 8(d), FP64 peak 36.5 TFLOP/s
measured by ssm_fp64_peak_kernel, HBM peak from MEASURED_PEAKS.json).  Data are simulated on the device from the
golden descriptors; the filters are the golden files' own (reference weights).

    python tools/time_configs.py [M] [--own]     --own: the package's own structured weights for the BQ / TPQ filters
                                                 (compact reflection-symmetric sums of the forward pass)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssmtoybox_b200 import device as dv  # noqa: E402

# name, golden, steps, FLOP/step filter (BQ or UT column), bytes/step filter-only, smoother
CONFIGS = [
    ('C1 UNGM UKF', 'c1_ungm_ukf', 500, 125, 24, True),
    ('C2 UNGM GPQ (one of 11 length-scales)', 'c2_ungm_gpq_el06', 500, 153, 24, True),
    ('C3 reentry GPQ', 'c3_reentry_gpq', 500, 5756, 256, True),
    ('C4 coordinated turn TPQ', 'c4_ct_tpq', 500, 5668, 256, True),
    ('C4 coordinated turn Student-t UKF', 'c4_ct_fsstudent', 500, 3823, 256, False),
    ('C5 pendulum BSQ', 'c5_pend_bsq', 500, 510, 56, True),
    ('C5 coordinated turn BSQ', 'c4_ct_bsq', 500, 5668, 256, True),
]


def ev_time(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    hbm = float(peaks.get('hbm_gbs', 6556.8))
    fp64 = 36.5
    own = '--own' in sys.argv
    args = [a for a in sys.argv[1:] if a != '--own']
    M_arg = int(args[0]) if args else 0
    print('| configuration | M x N | filter only ms | traj-steps/s | FP64 frac | HBM frac | + predictive moments + RTS smoother ms | traj-steps/s | failed |')
    print('|---|---|---:|---:|---:|---:|---:|---:|---:|')
    for label, name, N, flop, bts, smooth in CONFIGS:
        g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')))
        if own:
            if str(g['dyn_kind']) == 'sp':
                continue
            g = dv.own_weights(g)
            label += ' (own weights, compact sums: %s)' % (dv.weights_reflective(dv.lower(g)),)
        low = dv.lower(g)
        dx = low.dx
        # outputs with predictive moments: 8 (2 dx + 3 dx^2) bytes per step, + smoothed 8 (dx + dx^2): stay below ~60 GB
        per_step = 8 * (3 * dx + 4 * dx * dx + low.dy + dx)
        M = M_arg or int(min(2 ** 20, 60e9 / (per_step * N)) // 1184 * 1184)
        if 'reentry' in name:   # Euler-Maruyama truth, as bench.py (simulate_discrete diverges on this model)
            truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
                     'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
            x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
        else:
            x, y = dv.simulate(low, M, N, rng=dv.make_rng(g, seed=1))
        del x
        o = {}
        t_f = ev_time(lambda: dv.filter_forward(low, y, store_pred=False, out=o))
        nf = int((o['status'] != 0).sum())
        del o
        line = '| %s | %d x %d | %.2f | %.3e | %.1f %% | %.1f %% |' % (label, M, N, t_f, M * N / t_f * 1e3, 100 * M * N * flop / t_f * 1e-9 / fp64,
                                                                  100 * M * N * bts / t_f * 1e-6 / hbm)
        if smooth and low.family == 1:
            o = {}
            sm = {}

            def both():
                dv.filter_forward(low, y, store_pred=True, out=o)
                dv.smooth_backward(low.dx, o, out=sm)
            t_b = ev_time(both, reps=3)
            line += ' %.2f | %.3e | %d |' % (t_b, M * N / t_b * 1e3, nf)
            del o, sm
        else:
            line += ' — | — | %d |' % nf
        print(line, flush=True)
        del y
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
