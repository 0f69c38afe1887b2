"""Monte-Carlo experiment drivers as one call (the loops of research/gpq/icinco_demo.py:115-125 + :17-52,
research/gpq/gpq_tracking.py:52-57, research/bsq/bsq_tracking.py:300-337 on the GPU).

`filter_scores` runs filter (+ RTS smoother) and the error scores over all trajectories of host or device arrays.

Host inputs are streamed along the TIME axis (default path, everything resident): a trajectory is a serial
recursion, so a launch is only efficient when it spans (at least) a full wave of trajectories -- cutting the
trajectory axis into chunks leaves the GPU mostly idle (measured on B200, 125 000 x 500: 10 trajectory chunks
123 ms, PCIe alone 64 ms).  Instead the measurements are copied window by window in time order and every forward
pass launch (ssm_filter_window) covers ALL trajectories for one window as soon as that window has landed; the
truth follows in REVERSE time order while the forward pass is still running, and the RTS smoother
(ssm_smooth_window, in-kernel phase-1 statistics) chases it backwards.  The per-step MSE matrix of step k needs
step k only, so the second score phase of a window (log credibility ratio, utils.py:113-120) follows its smoother
window immediately: after the last byte of x has arrived only one window of smoother + scores remains.
With a Communicator, every rank handles its own trajectories; the packed statistics rows of each window are
all-reduced (tiny, latency-bound) before its second phase.

When the arrays of the batch do not fit in device memory the trajectory axis is cut into chunks instead
(`_filter_scores_chunked`): chunk c+1 is copied while chunk c is processed, and only the moments needed by the
second score phase stay resident.
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


def _windows(N, n):
    n = max(1, min(int(n), N))
    w = -(-N // n)
    return [(a, min(a + w, N)) for a in range(0, N, w)]


def filter_scores(alg, y, x, smooth=True, n_windows=20, comm=None, keep=False, n_chunks=None, reduce_every=4):
    """RMSE / NCI / NLL of filter `alg` on measurements y (dy, N, M) against the truth x (dx, N, M).
    y, x: numpy arrays, CPU torch tensors (pinned for full copy speed) or CUDA tensors.
    Returns the dict of ssmtoybox_b200.utils.evaluate_performance (+ 'status' (M,) int32 on the host, and the
    device arrays under 'arrays' when keep=True).  n_windows: time windows of the streaming pipeline;
    n_chunks forces the trajectory-chunked fallback.  keep=False (default) runs the RTS smoother in its score-only mode
    (ssm_smooth_scores): the smoothed moments are scored in registers and never stored.  reduce_every: smoother
    windows whose statistics rows share one all-reduce (and one second-phase launch)."""
    y, x = _as_host_or_device(y), _as_host_or_device(x)
    dy, N, M = y.shape
    dx = x.shape[0]
    dev = torch.device('cuda', torch.cuda.current_device())
    low = dv.lower(alg._describe())
    do_smooth = smooth and not isinstance(alg, StudentianInference)
    if low.dy != dy or low.dx != dx:
        raise ValueError('data dimensions do not match the model')
    per_traj = 8 * N * ((0 if y.is_cuda else dy) + (0 if x.is_cuda else dx) + (dx + dx * dx)
                        + ((2 * dx + 3 * dx * dx) if do_smooth else 0))
    free = torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
    # cudaMemcpy2DAsync pitches are limited (cudaDevAttrMaxPitch, 2 GiB): very long rows take the chunked path
    pitch_ok = (y.is_cuda and x.is_cuda) or N * M * 8 <= 2 ** 31 - 1
    chunked = n_chunks is not None or per_traj * M > 0.85 * free or N < 4 or not pitch_ok
    if comm is not None:
        # the two paths issue different numbers of all-reduces: every rank must take the same one
        flag = comm.allreduce_sum(torch.tensor([1.0 if chunked else 0.0], dtype=torch.float64, device=dev))
        chunked = bool(flag.item() > 0.0)
    if chunked:
        return _filter_scores_chunked(alg, y, x, smooth=smooth, n_chunks=n_chunks or 8, comm=comm, keep=keep)
    for src in (y, x):
        if not src.is_cuda and (src.dtype != torch.float64 or not src.is_contiguous()):
            raise ValueError('host arrays must be C-contiguous float64')
    wins = _windows(N, n_windows)
    comp = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    kw = dict(dtype=torch.float64, device=dev)

    def h2d_window(src, dst, k0, k1):
        """time steps [k0, k1) of a host (c, N, M) array -> the same slots of the device array: one strided DMA
        (c rows of (k1 - k0) * M contiguous doubles)"""
        rc = dv.lib.ssm_memcpy2d(C.c_void_p(dst.data_ptr() + k0 * M * 8), N * M * 8, C.c_void_p(src.data_ptr() + k0 * M * 8),
                                 N * M * 8, (k1 - k0) * M * 8, src.shape[0], 1, C.c_void_p(copy.cuda_stream))
        dv._lib.check(rc, 'ssm_memcpy2d')
        e = torch.cuda.Event()
        e.record(copy)
        return e

    # ---- copies: y in time order, then x in the order its consumer walks the windows ---------------
    ev_y, ev_x = [None] * len(wins), [None] * len(wins)
    copy.wait_stream(comp)
    yd = y if y.is_cuda else torch.empty((dy, N, M), **kw)
    xd = x if x.is_cuda else torch.empty((dx, N, M), **kw)
    with torch.cuda.stream(copy):
        if not y.is_cuda:
            yd.record_stream(copy)
            for c, (k0, k1) in enumerate(wins):
                ev_y[c] = h2d_window(y, yd, k0, k1)
        if not x.is_cuda:
            xd.record_stream(copy)
            order = range(len(wins) - 1, -1, -1) if do_smooth else range(len(wins))
            for c in order:
                ev_x[c] = h2d_window(x, xd, *wins[c])
    W = dv.lib.ssm_scores_width(dx)
    stats = torch.empty((N, W), **kw)
    acc = torch.empty((dx, M), **kw)
    mse = torch.empty((dx * dx, N), **kw)
    lcr = torch.empty((N, 2), **kw)

    def second_phase(k0, k1, mean, cov, status, quad=None):
        """stats rows [k0, k1) are complete on this rank: all-reduce them, form the MSE matrices, second phase"""
        rows = stats[k0:k1]
        if comm is not None:
            comm.allreduce_sum(rows)
        mse[:, k0:k1] = (rows[:, dx:dx + dx * dx] / rows[:, -1:]).T
        dv.scores_phase2(xd, mean, cov, mse, status, window=(k0, k1), out=lcr, quad=quad)

    # ---- forward pass, window by window --------------------------------------------------------------
    fwd = {}
    for c, (k0, k1) in enumerate(wins):
        if ev_y[c] is not None:
            comp.wait_event(ev_y[c])
        # (score-only smoother behind it: it reads the lower triangles of the covariances only, so only those are written)
        dv.filter_forward(low, yd, store_pred=do_smooth, out=fwd, window=(k0, k1), want_last=True,
                          init_mean=fwd['last_mean'] if c else None, init_cov=fwd['last_cov'] if c else None,
                          lower_only=do_smooth and not keep)
        if not do_smooth:
            if ev_x[c] is not None:
                comp.wait_event(ev_x[c])
            if c == 0:
                quad_f = torch.empty((N, M), **kw)
            dv.scores_phase1(xd, fwd['fi_mean'], fwd['fi_cov'], fwd['status'], window=(k0, k1), out=(stats, acc), quad=quad_f)
            if c == 0:
                cnt0 = stats[0, -1].clone()      # trajectories of this rank alive after the first window
            second_phase(k0, k1, fwd['fi_mean'], fwd['fi_cov'], fwd['status'], quad=quad_f)
    mean, cov, st = fwd['fi_mean'], fwd['fi_cov'], fwd['status']
    n_bad = torch.zeros((), **kw)
    # ---- RTS smoother + scores, walking the windows backwards ------------------------------------------
    if do_smooth:
        sm = {'stats': stats, 'rmse_acc': acc}
        pending = []
        for c in range(len(wins) - 1, -1, -1):
            k0, k1 = wins[c]
            if ev_x[c] is not None:
                comp.wait_event(ev_x[c])
            if keep:
                dv.smooth_backward(dx, fwd, out=sm, x_truth=xd, window=(k0, k1), want_quad=True)
            else:
                dv.smooth_scores(dx, fwd, xd, out=sm, window=(k0, k1))
            pending.append((k0, k1))
            if len(pending) >= max(1, int(reduce_every)) or c == 0:
                # adjacent windows: one all-reduce of their statistics rows, one second-phase launch
                a, b = pending[-1][0], pending[0][1]
                if keep:
                    second_phase(a, b, sm['sm_mean'], sm['sm_cov'], sm['status'], quad=sm['quad'])
                else:
                    rows = stats[a:b]
                    if comm is not None:
                        comm.allreduce_sum(rows)
                    mse[:, a:b] = (rows[:, dx:dx + dx * dx] / rows[:, -1:]).T
                    dv.scores_phase2_res(sm['dres'], sm['quad'], mse, sm['status'], window=(a, b), out=lcr)
                pending = []
        mean, cov, st = (sm['sm_mean'], sm['sm_cov'], sm['status']) if keep else (None, None, sm['status'])
        # trajectories that failed INSIDE the smoother were still alive in the rows of later steps
        n_bad = ((st != 0).sum() - (fwd['status'] != 0).sum()).to(torch.float64)
    else:
        # a trajectory that fails in a later window was still alive in the rows of earlier ones
        n_bad = (st != 0).sum().to(torch.float64) - (M - cnt0)
    ok = (st == 0)
    rm = torch.where(ok[None, :], torch.sqrt(acc / N), torch.zeros_like(acc)).sum(dim=1)
    pack = torch.cat([rm, lcr.reshape(-1), n_bad.reshape(1)])
    if comm is not None:
        pack = comm.allreduce_sum(pack)
    rm_g, lcr_g, bad = pack[:dx], pack[dx:dx + 2 * N].reshape(N, 2), pack[-1]
    out = finalize_scores(stats, rm_g, lcr_g, dx, N)
    out['n_bad'] = bad
    res = {k: (v.cpu().numpy() if v.ndim else float(v)) for k, v in out.items()}
    if res.pop('n_bad') != 0.0:
        # rare: some trajectories failed after they had contributed to some rows -> exact two-pass recomputation
        if mean is None:   # score-only smoother kept no arrays: run again with them
            return filter_scores(alg, yd, xd, smooth=smooth, n_windows=n_windows, comm=comm, keep=True, reduce_every=reduce_every)
        from . import utils as U
        res = U.evaluate_performance(xd, mean, cov, status=st, comm=comm)
    res['status'] = st.cpu().numpy()
    if keep:
        res['arrays'] = dict(x=xd, y=yd, mean=mean, cov=cov, status=st, fwd=fwd)
    return res


def _filter_scores_chunked(alg, y, x, smooth=True, n_chunks=8, comm=None, keep=False):
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

    # CUDA sources (e.g. simulate(..., device_out=True) just before) are sliced on the copy stream: it must not run
    # ahead of the kernels that produce them
    copy.wait_stream(comp)
    for src in (y, x):
        if src.is_cuda:
            src.record_stream(copy)
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


def monte_carlo_scores(alg, n_traj, n_steps, truth=None, sim='discrete', dt=0.0, sub=1, seed=0, chunk=131072, smooth=False,
                       comm=None, to_host=True):
    """Generator-driven Monte-Carlo experiment with nothing materialised per trajectory but the two score inputs:
    for every chunk of trajectories  simulate (Philox, keyed by the GLOBAL trajectory index) -> filter [-> RTS smoother]
    with the phase-1 error statistics accumulated IN the filter / smoother kernel; the states, measurements and filter
    arrays of a chunk are dropped when the chunk is done -- like the reference's loops, which keep one trajectory at a
    time (research/gpq/icinco_demo.py:115-125, research/bsq/bsq_tracking.py:300-306).  Per scored unit only
    d = x - m (dx doubles) and d' P^-1 d (1 double) stay resident: the log credibility ratio needs the GLOBAL per-step
    MSE matrix first (research/gpq/icinco_demo.py:34-40), so its second phase runs after the last chunk.

    alg: a Gaussian-family filter of ssmtoybox_b200.ssinf.  truth: dict(m0, P0, q_cov, r_cov[, x0_dof, q_dof, r_dof])
    of the data-generating system (default: the filter's own model), sim / dt / sub as device.simulate.
    With a Communicator every rank runs its contiguous share of the n_traj trajectories; the packed statistics are
    all-reduced once per phase.  Returns the dict of utils.evaluate_performance (+ n_failed, kept_bytes)."""
    from . import utils as U
    dev = torch.device('cuda', torch.cuda.current_device())
    d = alg._describe()
    low = dv.lower(d)
    dx, N = low.dx, int(n_steps)
    if isinstance(alg, StudentianInference):
        raise NotImplementedError('monte_carlo_scores: Gaussian-family filters')
    if truth is None:
        truth = {'m0': d['m0'], 'P0': d['P0'], 'q_cov': d['q_cov'], 'r_cov': d['r_cov']}
    off, cnt_loc = (0, int(n_traj)) if comm is None else comm.shard(int(n_traj))
    a, b = off, off + cnt_loc
    kw = dict(dtype=torch.float64, device=dev)
    W = dv.lib.ssm_scores_width(dx)
    stats = torch.zeros((N, W), **kw)
    rm = torch.zeros((dx,), **kw)
    kept, n_failed = [], 0
    scratch = {}
    chunk = max(128, int(chunk))
    for c0 in range(a, b, chunk):
        mc = min(chunk, b - c0)
        x, y = dv.simulate(low, mc, N, rng=dv.make_rng(truth, seed=seed, traj_offset=c0), mode=sim, dt=dt, sub=sub, device=dev)
        if scratch.get('m') != mc:
            scratch = {'m': mc, 'fwd': {}}
        sc = {}
        if smooth:
            fwd = dv.filter_forward(low, y, store_pred=True, out=scratch['fwd'], lower_only=True)
            dv.smooth_scores(dx, fwd, x, out=sc)
        else:
            try:
                dv.filter_scored(low, y, x, out=sc)
            except NotImplementedError:   # no scoring instantiation of this filter: score the stored moments of the chunk
                fwd = dv.filter_forward(low, y, store_pred=False, out=scratch['fwd'])
                sc['quad'] = torch.empty((N, mc), **kw)
                sc['stats'], sc['rmse_acc'] = dv.scores_phase1(x, fwd['fi_mean'], fwd['fi_cov'], fwd['status'], quad=sc['quad'])
                sc['dres'] = x - fwd['fi_mean']
                sc['status'] = fwd['status'].clone()
        st = sc['status']
        stats += sc['stats']
        ok = (st == 0)
        rm += torch.where(ok[None, :], torch.sqrt(sc['rmse_acc'] / N), torch.zeros_like(sc['rmse_acc'])).sum(dim=1)
        kept.append((sc['dres'], sc['quad'], st))
        del x, y
    pack = torch.cat([stats.reshape(-1), rm])
    if comm is not None:
        pack = comm.allreduce_sum(pack)
    st_g, rm_g = pack[:N * W].reshape(N, W), pack[N * W:]
    cnt = st_g[:, -1]
    mse = (st_g[:, dx:dx + dx * dx] / cnt[:, None]).T.reshape(dx, dx, N).contiguous()
    lcr = torch.zeros((N, 2), **kw)
    for dres, quad, st in kept:
        lcr += dv.scores_phase2_res(dres, quad, mse, st)
    nf = torch.stack([(st != 0).sum() for _, _, st in kept]).sum().to(torch.float64).reshape(1) if kept else torch.zeros(1, **kw)
    if comm is not None:
        lcr = comm.allreduce_sum(lcr)
        nf = comm.allreduce_sum(nf)
    out = finalize_scores(st_g, rm_g, lcr, dx, N)
    out['n_failed'] = nf[0]
    if to_host:
        out = {k: (v.cpu().numpy() if v.ndim else float(v)) for k, v in out.items()}
    out['kept_bytes'] = int(sum(t[0].numel() + t[1].numel() for t in kept) * 8)
    return out
