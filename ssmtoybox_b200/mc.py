"""Monte-Carlo experiment drivers as one call (the loops of research/gpq/icinco_demo.py:115-125 + :17-52,
research/gpq/gpq_tracking.py:52-57, research/bsq/bsq_tracking.py:300-337 on the GPU).

`filter_scores` runs filter (+ RTS smoother) and the error scores over all trajectories of host or device arrays.
Host inputs are streamed: the trajectory axis is cut into chunks, chunk c+1 is copied host -> device on a second
CUDA stream while chunk c is filtered, smoothed and reduced; only the per-chunk smoothed (or filtered) moments
and the truth stay resident for the second score phase (the log credibility ratio needs the GLOBAL per-step MSE
matrix first, utils.py:113-120).  With a Communicator, every rank handles its own trajectories and the packed
statistics are all-reduced (one NCCL call per phase).
"""
import ctypes as C

import numpy as np
import torch

from . import device as dv
from .dist import finalize_scores
from .ssinf import StudentianInference


def _as_host_or_device(a):
    if isinstance(a, torch.Tensor):
        return a
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))


def filter_scores(alg, y, x, smooth=True, n_chunks=8, comm=None, keep=False):
    """RMSE / NCI / NLL of filter `alg` on measurements y (dy, N, M) against the truth x (dx, N, M).
    y, x: numpy arrays, CPU torch tensors (pinned for full copy speed) or CUDA tensors.
    Returns the dict of ssmtoybox_b200.utils.evaluate_performance (+ 'status' (M,) int32 on the host, and the
    per-chunk device arrays under 'chunks' when keep=True)."""
    y, x = _as_host_or_device(y), _as_host_or_device(x)
    dy, N, M = y.shape
    dx = x.shape[0]
    dev = torch.device('cuda', torch.cuda.current_device())
    low = dv.lower(alg._describe())
    do_smooth = smooth and not isinstance(alg, StudentianInference)
    if low.dy != dy or low.dx != dx:
        raise ValueError('data dimensions do not match the model')
    n_chunks = max(1, min(int(n_chunks), (M + 127) // 128))
    mc = -(-M // n_chunks)
    mc = ((mc + 127) // 128) * 128
    bounds = [(a, min(a + mc, M)) for a in range(0, M, mc)]
    comp = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    kw = dict(dtype=torch.float64, device=dev)
    W = dv.lib.ssm_scores_width(dx)
    stats = torch.zeros((N, W), **kw)
    rm = torch.zeros((dx,), **kw)
    kept, status_all = [], []
    scratch = {}          # forward-pass arrays of the chunk in flight (reused by equally sized chunks)

    def h2d(src, a, b):
        """columns [a, b) of a host (c, N, M) array -> compact device (c, N, b-a): one strided DMA"""
        if src.is_cuda:
            return src[:, :, a:b].contiguous()
        if src.dtype != torch.float64 or not src.is_contiguous():
            raise ValueError('host arrays must be C-contiguous float64')
        dst = torch.empty((src.shape[0], src.shape[1], b - a), **kw)
        rc = dv.lib.ssm_memcpy2d(dv._p(dst), (b - a) * 8, C.c_void_p(src.data_ptr() + a * 8), src.shape[2] * 8, (b - a) * 8,
                                 src.shape[0] * src.shape[1], 1, C.c_void_p(copy.cuda_stream))
        dv._lib.check(rc, 'ssm_memcpy2d')
        return dst

    def stage(c):
        a, b = bounds[c]
        with torch.cuda.stream(copy):
            ys, xs = h2d(y, a, b), h2d(x, a, b)
            ev = torch.cuda.Event()
            ev.record(copy)
        return ys, xs, ev

    nxt = stage(0)
    for c in range(len(bounds)):
        ys, xs, ev = nxt
        if c + 1 < len(bounds):
            nxt = stage(c + 1)
        comp.wait_event(ev)
        ys.record_stream(comp), xs.record_stream(comp)
        m = ys.shape[-1]
        if scratch.get('m') != m:
            scratch = {'m': m, 'fwd': {}}
        fwd = dv.filter_forward(low, ys, store_pred=do_smooth, out=scratch['fwd'])
        if do_smooth:
            # fresh outputs (they stay resident for phase 2); phase-1 statistics accumulated inside the smoother
            sm = dv.smooth_backward(low.dx, fwd, x_truth=xs)
            mean, cov, st = sm['sm_mean'], sm['sm_cov'], sm['status']
            s1, acc = sm['stats'], sm['rmse_acc']
        else:
            mean, cov, st = fwd['fi_mean'].clone(), fwd['fi_cov'].clone(), fwd['status'].clone()
            s1, acc = dv.scores_phase1(xs, mean, cov, st)
        stats += s1
        ok = (st == 0)
        rm += torch.where(ok[None, :], torch.sqrt(acc / N), torch.zeros_like(acc)).sum(dim=1)
        kept.append((xs, mean, cov, st))
        status_all.append(st)
    pack = torch.cat([stats.reshape(-1), rm])
    if comm is not None:
        pack = comm.allreduce_sum(pack)
    st_g, rm_g = pack[:N * W].reshape(N, W), pack[N * W:]
    cnt = st_g[:, -1]
    mse = (st_g[:, dx:dx + dx * dx] / cnt[:, None]).T.reshape(dx, dx, N).contiguous()
    lcr = torch.zeros((N, 2), **kw)
    for xs, mean, cov, st in kept:
        lcr += dv.scores_phase2(xs, mean, cov, mse, st)
    if comm is not None:
        lcr = comm.allreduce_sum(lcr)
    out = finalize_scores(st_g, rm_g, lcr, dx, N)
    res = {k: (v.cpu().numpy() if v.ndim else float(v)) for k, v in out.items()}
    res['status'] = torch.cat(status_all).cpu().numpy()
    if keep:
        res['chunks'] = kept
    return res
