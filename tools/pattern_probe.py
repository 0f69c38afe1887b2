"""Developer tool: HBM bandwidth of the score-only smoother's ACCESS PATTERN with nothing computed (tools/pattern_probe.cu),
next to a contiguous device copy in the same process.  usage: python tools/pattern_probe.py [M] [N]"""
import ctypes as C, json, os, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, 'pattern_probe.so'))
lib.probe_run.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
M = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
torch.cuda.set_device(0)
assert lib.probe_init() == 0
src = torch.zeros((90, N, M), dtype=torch.float64, device='cuda')
dst = torch.empty((6, N, M), dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


out = {'M': M, 'N': N, 'bytes_per_unit': 608}
for name, minb in (('pattern_2_ctas_per_sm', 2), ('pattern_4_ctas_per_sm', 0)):
    def run(minb=minb):
        rc = lib.probe_run(src.data_ptr(), dst.data_ptr(), M, N, minb, st)
        assert rc == 0, rc
    med, best = timed(run)
    out[name] = {'ms_median': med, 'ms_min': best, 'GBps_median': 608.0 * M * N / med / 1e6}
lib.probe_ring_run.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
for depth in (2, 3):   # cp.async ring: depth - 1 steps in flight per thread (depth 2: 140 KB -> 1 CTA per SM as well)
    def run(depth=depth):
        rc = lib.probe_ring_run(src.data_ptr(), dst.data_ptr(), M, N, depth, st)
        assert rc == 0, rc
    med, best = timed(run)
    out['ring_depth_%d_1_cta_per_sm' % depth] = {'ms_median': med, 'ms_min': best, 'GBps_median': 608.0 * M * N / med / 1e6}
a = src[:16].reshape(-1)      # 16 planes = 8 GB at the default size: far larger than the L2
b = src[16:32].reshape(-1)
med, best = timed(lambda: b.copy_(a))
out['contiguous_copy'] = {'ms_median': med, 'GBps_median': 2.0 * a.numel() * 8 / med / 1e6, 'bytes': 2 * a.numel() * 8}
try:
    out['measured_peak_GBps'] = json.load(open(os.path.join(HERE, '..', 'MEASURED_PEAKS.json')))
except Exception:
    pass
print(json.dumps(out))
