"""Developer tool: host->device copy bandwidth from pinned memory, contiguous vs strided (cudaMemcpy2DAsync) as used
by mc.filter_scores, and the e2e pipeline per-chunk timeline."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv
torch.cuda.set_device(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
M, N = 125000, 500
xh = torch.empty((5, N, M), dtype=torch.float64, pin_memory=True); xh.fill_(1.0)
xd = torch.empty((5, N, M), dtype=torch.float64, device='cuda')
s = torch.cuda.Stream()
def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); a, b = ev(), ev(); a.record(s); fn(); b.record(s); b.synchronize(); best = min(best, a.elapsed_time(b))
    return best
with torch.cuda.stream(s):
    ms = timeit(lambda: xd.copy_(xh, non_blocking=True))
    print('contiguous 2.5 GB: %.1f ms  %.1f GB/s' % (ms, xh.numel() * 8 / ms / 1e6))
    for nch in (4, 10, 20, 40):
        mc = ((-(-M // nch)) + 127) // 128 * 128
        def f():
            for a in range(0, M, mc):
                b = min(a + mc, M)
                dst = xd.view(-1)[: 5 * N * (b - a)]
                rc = dv.lib.ssm_memcpy2d(dv._p(dst), (b - a) * 8, C.c_void_p(xh.data_ptr() + a * 8), M * 8, (b - a) * 8, 5 * N, 1, C.c_void_p(s.cuda_stream))
                assert rc == 0
        ms = timeit(f)
        print('strided 2D, %d chunks: %.1f ms  %.1f GB/s' % (nch, ms, xh.numel() * 8 / ms / 1e6))
    # D2H for reference
    ms = timeit(lambda: xh.copy_(xd, non_blocking=True))
    print('D2H contiguous 2.5 GB: %.1f ms  %.1f GB/s' % (ms, xh.numel() * 8 / ms / 1e6))
# ---- e2e pipeline vs chunk count
import bench
from ssmtoybox_b200 import mc as MC
alg, g = bench.build_filter()
low = dv.lower(alg._describe())
truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]), 'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
del xd
x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=1), mode='continuous', dt=0.05, sub=2)
xh.copy_(x); yh = torch.empty(y.shape, dtype=torch.float64, pin_memory=True).copy_(y)
del x, y
torch.cuda.empty_cache()
for kw in ({'n_chunks': 5}, {'n_windows': 4}, {'n_windows': 10}, {'n_windows': 20}, {'n_windows': 50}):
    ts = []
    for i in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = MC.filter_scores(alg, yh, xh, smooth=True, **kw)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print('e2e %s: %s ms   nci=%.6f' % (kw, ' '.join('%.1f' % t for t in ts), out['nci']))
