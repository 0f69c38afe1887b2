"""Developer tool: wall time of the research drivers (ssmtoybox_b200/research/) at the reference's own sizes and at
10^5 Monte-Carlo simulations; data simulated on the device."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200.research import icinco_demo, bsq_ungm, bsq_tracking, gpq_tracking


def timed(label, fn, reps=2):
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%-58s %8.3f s' % (label, dt), flush=True)
    return out


timed('icinco_demo.tables(500, 100)  14 filters+smoothers, bootstrap', lambda: icinco_demo.tables(500, 100))
timed('icinco_demo.tables(500, 100000)', lambda: icinco_demo.tables(500, 100000), reps=1)
timed('icinco_demo.hypers_demo(11 length-scales, 500 x 100)', lambda: icinco_demo.hypers_demo())
timed('icinco_demo.hypers_demo(11 length-scales, 500 x 100000)   [C2]', lambda: icinco_demo.hypers_demo(mc=100000), reps=1)
timed('bsq_ungm.tables(500, 100)  9 filters+smoothers, bootstrap', lambda: bsq_ungm.tables(500, 100))
timed('bsq_tracking.reentry_demo(dur=200, mc=100)  2000 steps, 4 filters', lambda: bsq_tracking.reentry_demo(200, 100))
timed('bsq_tracking.reentry_demo(dur=50, mc=100000)  500 steps', lambda: bsq_tracking.reentry_demo(50, 100000), reps=1)
timed('gpq_tracking.reentry_simple_gpq_demo(dur=30, mc=100)', lambda: gpq_tracking.reentry_simple_gpq_demo())
o = timed('gpq_tracking.reentry_gpq_demo(mc=20000, duration=50)', lambda: gpq_tracking.reentry_gpq_demo(20000, 50), reps=1)
print('   avg position RMSE [GPQKF, UKF]', o['avg_rmse'], 'inclination', o['avg_inc'], 'failed', o['n_failed'])
