"""Low-level batched device API: lowers a flat filter description (plain dict of numpy arrays)
to the C-ABI descriptor (ssm_desc) and launches the CUDA kernels on torch CUDA tensors.

All bulk arrays are float64 torch tensors on one CUDA device laid out (component..., step,
trajectory) with the trajectory axis contiguous -- the natural shape of the reference's batched
arrays (dim, N, M).  PyTorch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(_lib.c_double_p)


class Lowered:
    """An ssm_desc plus the host arrays it points to (kept alive with it)."""

    def __init__(self, desc, keep, dx, dy, family):
        self.desc, self.keep, self.dx, self.dy, self.family = desc, keep, dx, dy, family


def _lower_transform(d, prefix, keep):
    kind = str(d[prefix + 'kind'])
    pts = _c(d[prefix + 'points'])
    dim_in, n = pts.shape
    wm = _c(d[prefix + 'wm']).reshape(n)
    Wc = _c(d[prefix + 'Wc'])
    if Wc.ndim == 1:
        Wc = np.diag(Wc)
    Wc = _c(Wc).reshape(n, n)
    t = _lib.SsmTransform()
    t.dim_in, t.n_pts = dim_in, n
    t.points, t.wm, t.Wc = _ptr(pts), _ptr(wm), _ptr(Wc)
    keep += [pts, wm, Wc]
    dim_out = int(d[prefix + 'dim_out_fn'])
    t.dim_out = dim_out
    if kind == 'sp':
        t.kind = _lib.TF_SP
        return t
    Wcc = _c(d[prefix + 'Wcc']).reshape(dim_in, n)
    t.Wcc = _ptr(Wcc)
    keep.append(Wcc)
    mv = np.asarray(d[prefix + 'model_var'], dtype=np.float64)
    if kind == 'tp':
        t.kind = _lib.TF_TP
        mva = _c(mv.reshape(-1)[:1])
        iK = _c(d[prefix + 'iK']).reshape(n, n)
        t.iK = _ptr(iK)
        t.nu = float(d[prefix + 'nu'])
        # I_out is (dim_out x dim_out) of the transform object; 1x1 -> the full matrix is added
        t.tp_full_matrix = 1 if int(d.get(prefix + 'dim_out', 1)) == 1 else 0
        keep += [mva, iK]
    else:
        t.kind = _lib.TF_BQ
        mva = _c(mv * np.eye(dim_out))  # model_var * I_out, bqmtran.py:198 (scalar or assigned matrix)
        keep.append(mva)
    t.model_var = _ptr(mva)
    return t


def lower(d):
    """dict -> Lowered.  Keys follow tests/golden/*.npz (see oracle/gen_golden.py):
    dyn_name, obs_name, dyn_dt, G, state_index, radar_loc, m0, P0, q_cov, r_cov,
    {dyn,obs}_{kind,points,wm,Wc,Wcc,model_var,iK,nu,dim_out} and, for the Student family,
    dof, fixed_dof, x0_dof, q_dof, r_dof."""
    keep = []
    dyn_name, obs_name = str(d['dyn_name']), str(d['obs_name'])
    if dyn_name not in _lib.DYN_IDS:
        raise NotImplementedError('no device implementation of transition model ' + dyn_name)
    if obs_name not in _lib.OBS_IDS:
        raise NotImplementedError('no device implementation of measurement model ' + obs_name)
    m0, P0 = _c(d['m0']), _c(d['P0'])
    dx = m0.shape[0]
    R = _c(d['r_cov'])
    dy = R.shape[0]
    G = _c(d['G'])
    q_cov = np.atleast_2d(_c(d['q_cov']))
    GQG = _c(G.dot(q_cov).dot(G.T))  # ssinf.py:279
    keep += [m0, P0, R, GQG]
    s = _lib.SsmDesc()
    s.dyn_model, s.obs_model, s.dx, s.dy = _lib.DYN_IDS[dyn_name], _lib.OBS_IDS[obs_name], dx, dy
    s.dyn_par[0] = float(d.get('dyn_dt', 0.0))
    rl = np.asarray(d.get('radar_loc', [0.0, 0.0]), dtype=np.float64).reshape(-1)
    for i, v in enumerate(rl[:8]):      # radar_loc (2) / RangeMeasurement sensor (2) / BearingMeasurement sensor_pos (4, 2)
        s.obs_par[i] = float(v)
    si = np.asarray(d.get('state_index', []), dtype=np.int64).reshape(-1)
    s.n_state_index = len(si)
    for i, v in enumerate(si):
        s.state_index[i] = int(v)
    s.m0, s.P0, s.GQG, s.R = _ptr(m0), _ptr(P0), _ptr(GQG), _ptr(R)
    # noise moments for models with non-additive noise (augmented transforms, ssinf.py:271-272, 282-283)
    dq = q_cov.shape[0]
    q_mean = _c(np.asarray(d.get('q_mean', np.zeros(dq)), dtype=np.float64).reshape(-1))
    r_mean = _c(np.asarray(d.get('r_mean', np.zeros(dy)), dtype=np.float64).reshape(-1))
    q_cov = _c(q_cov)
    keep += [q_mean, r_mean, q_cov]
    s.q_mean, s.q_cov, s.r_mean, s.dq = _ptr(q_mean), _ptr(q_cov), _ptr(r_mean), dq
    if 'dof' in d:
        s.family = _lib.FAMILY_STUDENT
        s.dof, s.x0_dof = float(d['dof']), float(d['x0_dof'])
        s.q_dof, s.r_dof = float(d['q_dof']), float(d['r_dof'])
        s.fixed_dof = int(d['fixed_dof'])
    else:
        s.family = _lib.FAMILY_GAUSS
    dd = dict(d)
    dd['dyn_dim_out_fn'], dd['obs_dim_out_fn'] = dx, dy
    s.tf_dyn = _lower_transform(dd, 'dyn_', keep)
    s.tf_obs = _lower_transform(dd, 'obs_', keep)
    return Lowered(s, keep, dx, dy, s.family)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_bulk(t, lead, n_steps, ld, name):
    if t is None:
        return
    if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
        raise ValueError('{} must be a contiguous float64 CUDA tensor'.format(name))
    if tuple(t.shape) != tuple(lead) + (n_steps, ld):
        raise ValueError('{} has shape {}, expected {}'.format(name, tuple(t.shape), tuple(lead) + (n_steps, ld)))


def bulk_ld(t):
    """Leading dimension of a [component..., step, trajectory] tensor that may be a trajectory-range view
    t_full[..., a:b] of a contiguous array: trajectory stride 1, step stride ld, component strides N*ld multiples."""
    if t.dtype != torch.float64 or not t.is_cuda:
        raise ValueError('bulk arrays must be float64 CUDA tensors')
    st, sh = t.stride(), t.shape
    ld = st[-2] if sh[-2] > 1 else max(sh[-1], 1)
    ok = (st[-1] == 1 or sh[-1] <= 1) and ld >= sh[-1]
    exp = ld * sh[-2]
    for d in range(t.ndim - 3, -1, -1):
        ok = ok and (st[d] == exp or sh[d] <= 1)
        exp *= sh[d]
    if not ok:
        raise ValueError('bulk array must be C-contiguous or a trajectory-range view of a C-contiguous array')
    return int(ld)


def filter_forward(low, y, store_pred=False, store_cov=True, init_mean=None, init_cov=None, t_offset=None, k0=0,
                   want_last=False, out=None, window=None, lower_only=False):
    """Run the fused forward pass (ssm_filter) on y (dy, N, M) -> dict of device tensors.
    lower_only: write only the lower triangles (column <= row) of fi_cov / pr_cov (ssm_filter_window_lower) -- for
    pipelines whose only consumer is smooth_scores, which reads nothing else; the other entries stay uninitialised.
    window = (k_lo, k_hi): process only those time steps of the N slots (ssm_filter_window); successive windows
    carry the state through out['last_mean'/'last_cov'] -> init_mean / init_cov and the status through out['status']."""
    dy, N, M = y.shape
    dx = low.dx
    dev = y.device
    ld = bulk_ld(y)
    kw = dict(dtype=torch.float64, device=dev)
    o = out if out is not None else {}
    if 'fi_mean' not in o:
        o['fi_mean'] = torch.empty((dx, N, M), **kw)
    if store_cov and 'fi_cov' not in o:
        o['fi_cov'] = torch.empty((dx, dx, N, M), **kw)
    if store_pred:
        for k, shp in (('pr_mean', (dx, N, M)), ('pr_cov', (dx, dx, N, M)), ('pr_xx_cov', (dx, dx, N, M))):
            if k not in o:
                o[k] = torch.empty(shp, **kw)
    if want_last and 'last_mean' not in o:
        o['last_mean'] = torch.empty((dx, M), **kw)
        o['last_cov'] = torch.empty((dx, dx, M), **kw)
    if 'status' not in o:
        o['status'] = torch.empty((M,), dtype=torch.int32, device=dev)
    if init_mean is not None:
        init_mean = init_mean.contiguous()
        init_cov = init_cov.contiguous()
    if M == 0 or N == 0:
        o['status'].zero_()
        return o
    for k in ('fi_mean', 'fi_cov', 'pr_mean', 'pr_cov', 'pr_xx_cov'):
        if o.get(k) is not None and bulk_ld(o[k]) != ld and M > 1:
            raise ValueError('all bulk arrays of one call must share the leading dimension (y: {}, {}: {})'.format(ld, k, bulk_ld(o[k])))
    if (init_mean is not None or want_last) and ld != M:
        raise ValueError('init / last moments are not supported on trajectory-range views')
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    rc = (lib.ssm_filter_window_lower if lower_only else lib.ssm_filter_window)(
                               C.byref(low.desc), _p(y), _p(o.get('fi_mean')), _p(o.get('fi_cov')), _p(o.get('pr_mean')),
                               _p(o.get('pr_cov')), _p(o.get('pr_xx_cov')), _p(init_mean), _p(init_cov),
                               _p(o.get('last_mean') if want_last else None), _p(o.get('last_cov') if want_last else None),
                               _p(t_offset), int(k0), _p(o['status']), M, N, k_lo, k_hi, ld, _stream())
    _lib.check(rc, 'ssm_filter')
    return o


def filter_scored(low, y, x_truth, out=None, window=None, init_mean=None, init_cov=None, t_offset=None, k0=0,
                  want_res=True, keep_moments=False):
    """Scoring forward pass (ssm_filter_scores): the filter of filter_forward with the phase-1 error statistics of the
    filtered moments accumulated in-kernel; no moment arrays are stored unless keep_moments.  Returns / fills
    out['stats'] (N, W), out['rmse_acc'] (dx, M), out['quad'] (N, M), out['dres'] (dx, N, M) (want_res), out['status'],
    out['last_mean' / 'last_cov'] -- the inputs of utils.evaluate_scored.  Raises NotImplementedError for filters
    without a scoring instantiation (non-additive noise, generic point sets, Student family)."""
    dy, N, M = y.shape
    dx = low.dx
    dev = y.device
    ld = bulk_ld(y)
    if ld != M and M > 1:
        raise ValueError('filter_scored is not supported on trajectory-range views')
    if bulk_ld(x_truth) != ld and M > 1:
        raise ValueError('x_truth must share the leading dimension of y')
    kw = dict(dtype=torch.float64, device=dev)
    o = out if out is not None else {}
    if 'stats' not in o:
        o['stats'] = torch.empty((N, lib.ssm_scores_width(dx)), **kw)
        o['rmse_acc'] = torch.empty((dx, M), **kw)
    if 'quad' not in o:
        o['quad'] = torch.empty((N, M), **kw)
    if want_res and 'dres' not in o:
        o['dres'] = torch.empty((dx, N, M), **kw)
    if keep_moments and 'fi_mean' not in o:
        o['fi_mean'] = torch.empty((dx, N, M), **kw)
        o['fi_cov'] = torch.empty((dx, dx, N, M), **kw)
    if 'last_mean' not in o:
        o['last_mean'] = torch.empty((dx, M), **kw)
        o['last_cov'] = torch.empty((dx, dx, M), **kw)
    if 'status' not in o:
        o['status'] = torch.empty((M,), dtype=torch.int32, device=dev)
    if M == 0 or N == 0:
        o['status'].zero_()
        return o
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    if init_mean is not None:
        init_mean, init_cov = init_mean.contiguous(), init_cov.contiguous()
    rc = lib.ssm_filter_scores(C.byref(low.desc), _p(y), _p(x_truth), _p(o.get('fi_mean') if keep_moments else None),
                               _p(o.get('fi_cov') if keep_moments else None), _p(o['stats']), _p(o['rmse_acc']), _p(o['quad']),
                               _p(o.get('dres') if want_res else None), _p(init_mean), _p(init_cov), _p(o['last_mean']),
                               _p(o['last_cov']), _p(t_offset), int(k0), _p(o['status']), M, N, k_lo, k_hi, ld, _stream())
    _lib.check(rc, 'ssm_filter_scores')
    return o


def smooth_backward(dx, fwd, out=None, x_truth=None, window=None, want_quad=False):
    """Run the RTS smoother (ssm_smooth) over the arrays stored by filter_forward(store_pred=True).
    window = (k_lo, k_hi): smooth only those steps (ssm_smooth_window); windows must be walked from the last to the
    first with the same `out` (the first call, k_hi == N, copies the forward-pass status).
    With x_truth (dx, N, M) the kernel also accumulates the phase-1 score statistics of the smoothed moments:
    out['stats'] (N, W) and out['rmse_acc'] (dx, M), identical to scores_phase1(x_truth, sm_mean, sm_cov, status);
    want_quad: also out['quad'] (N, M) = d' P_s^-1 d per unit for scores_phase2(..., quad=...) (ssm_smooth_quad)."""
    _, N, M = fwd['fi_mean'].shape
    o = out if out is not None else {}
    if 'sm_mean' not in o:
        o['sm_mean'] = torch.empty_like(fwd['fi_mean'])
        o['sm_cov'] = torch.empty_like(fwd['fi_cov'])
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    if k_hi == N or 'status' not in o:
        if 'status' in o and o['status'].shape == fwd['status'].shape:
            o['status'].copy_(fwd['status'])
        else:
            o['status'] = fwd['status'].clone()
    ld = bulk_ld(fwd['fi_mean'])
    if any(bulk_ld(t) != ld for t in (fwd['fi_cov'], fwd['pr_mean'], fwd['pr_cov'], fwd['pr_xx_cov'], o['sm_mean'], o['sm_cov'])) and M > 1:
        raise ValueError('all bulk arrays of one ssm_smooth call must share the leading dimension')
    if x_truth is not None:
        if bulk_ld(x_truth) != ld and M > 1:
            raise ValueError('x_truth must share the leading dimension of the other bulk arrays')
        if ld != M:   # rmse_acc (dx, M) is indexed with the bulk leading dimension inside the kernel
            raise ValueError('in-kernel scoring (x_truth) is not supported on trajectory-range views')
        W = lib.ssm_scores_width(dx)
        if 'stats' not in o:
            o['stats'] = torch.empty((N, W), dtype=torch.float64, device=fwd['fi_mean'].device)
            o['rmse_acc'] = torch.empty((dx, M), dtype=torch.float64, device=fwd['fi_mean'].device)
    quad = None
    if want_quad:
        if x_truth is None:
            raise ValueError('want_quad needs x_truth')
        if ld != M:
            raise ValueError('want_quad is not supported on trajectory-range views')
        if 'quad' not in o:
            o['quad'] = torch.empty((N, M), dtype=torch.float64, device=fwd['fi_mean'].device)
        quad = o['quad']
    rc = lib.ssm_smooth_quad(dx, _p(fwd['fi_mean']), _p(fwd['fi_cov']), _p(fwd['pr_mean']), _p(fwd['pr_cov']),
                             _p(fwd['pr_xx_cov']), _p(o['sm_mean']), _p(o['sm_cov']), _p(o['status']),
                             _p(x_truth), _p(o.get('stats') if x_truth is not None else None),
                             _p(o.get('rmse_acc') if x_truth is not None else None), _p(quad), M, N, k_lo, k_hi, ld, _stream())
    _lib.check(rc, 'ssm_smooth')
    return o


def smooth_scores(dx, fwd, x_truth, out=None, window=None, want_res=True):
    """Score-only RTS smoother (ssm_smooth_scores): the recursion of smooth_backward with the phase-1 statistics
    accumulated in-kernel and NO smoothed arrays stored.  out['stats'] (N, W), out['rmse_acc'] (dx, M), out['quad']
    (N, M) = d' P_s^-1 d, out['dres'] (dx, N, M) = x - m_s (want_res) -- the inputs of scores_phase2_res --,
    out['status'].  window = (k_lo, k_hi): walk the windows from the last to the first with the same `out`."""
    _, N, M = fwd['fi_mean'].shape
    o = out if out is not None else {}
    dev = fwd['fi_mean'].device
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    if k_hi == N or 'status' not in o:
        if 'status' in o and o['status'].shape == fwd['status'].shape:
            o['status'].copy_(fwd['status'])
        else:
            o['status'] = fwd['status'].clone()
    ld = bulk_ld(fwd['fi_mean'])
    if ld != M and M > 1:
        raise ValueError('smooth_scores is not supported on trajectory-range views')
    if any(bulk_ld(t) != ld for t in (fwd['fi_cov'], fwd['pr_mean'], fwd['pr_cov'], fwd['pr_xx_cov'], x_truth)) and M > 1:
        raise ValueError('all bulk arrays of one ssm_smooth_scores call must share the leading dimension')
    kw = dict(dtype=torch.float64, device=dev)
    if 'stats' not in o:
        o['stats'] = torch.empty((N, lib.ssm_scores_width(dx)), **kw)
        o['rmse_acc'] = torch.empty((dx, M), **kw)
    if 'quad' not in o:
        o['quad'] = torch.empty((N, M), **kw)
    if want_res and 'dres' not in o:
        o['dres'] = torch.empty((dx, N, M), **kw)
    if (k_lo > 0 or k_hi < N) and 'carry' not in o:
        o['carry'] = torch.empty((dx + dx * (dx + 1) // 2, M), **kw)
    rc = lib.ssm_smooth_scores(dx, _p(fwd['fi_mean']), _p(fwd['fi_cov']), _p(fwd['pr_mean']), _p(fwd['pr_cov']),
                               _p(fwd['pr_xx_cov']), _p(o['status']), _p(x_truth), _p(o['stats']), _p(o['rmse_acc']),
                               _p(o['quad']), _p(o.get('dres') if want_res else None), _p(o.get('carry')),
                               M, N, k_lo, k_hi, ld, _stream())
    _lib.check(rc, 'ssm_smooth_scores')
    return o


def scores_phase2_res(dres, quad, mse, status=None, window=None, out=None, lcr_acc=None):
    """Second score phase from the stored errors d = x - m (dx, N, M) and quadratic forms d' P^-1 d (N, M)
    (ssm_scores_phase2_res): per-step sums of the log credibility ratio and of its absolute value, (N, 2)."""
    dx, N, M = dres.shape
    _check_score_layout(dres, None, None, quad)
    lcr = out if out is not None else torch.empty((N, 2), dtype=torch.float64, device=dres.device)
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    rc = lib.ssm_scores_phase2_res(dx, _p(dres), _p(quad), _p(status), _p(mse.contiguous()), _p(lcr), _p(lcr_acc), M, N, k_lo, k_hi, M, _stream())
    _lib.check(rc, 'ssm_scores_phase2_res')
    return lcr


def fp64_peak(n_blocks=148 * 8, n_iters=4096, reps=5, mode=0):
    """Measured FP64 FMA throughput (FLOP/s) of the device: the roofline denominator for the
    FP64-bound kernels (MEASURED_PEAKS.json carries HBM and bf16 only).  mode = number of register
    operands beyond the first: 0 acc=fma(acc,c,c) (headline peak), 1 acc=fma(x,c,acc), 2 acc=fma(x,y,acc)."""
    if mode:
        n_iters = -((mode << 24) + n_iters)
    sink = torch.zeros(1, dtype=torch.float64, device='cuda')
    flops = C.c_double(0.0)
    best = 0.0
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ssm_fp64_peak_kernel(n_blocks, n_iters, _p(sink), C.byref(flops), _stream()), 'fp64_peak')
        e1.record()
        e1.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3))
    return best


# ------------------------------------------------------------------------------------------------
# K1: simulation
# ------------------------------------------------------------------------------------------------
def _cov_factor(cov):
    """A with A A^T = cov for a (possibly singular) PSD matrix; same construction as
    numpy.random.multivariate_normal's SVD path (the reference's sampler, utils.py:619)."""
    cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    u, s, vh = np.linalg.svd(cov)
    return _c((vh.T * np.sqrt(s)))


def make_rng(d, seed, traj_offset=0):
    """dict description (m0, P0, q_cov, r_cov [, x0_dof, q_dof, r_dof]) -> (SsmRng, keepalive)."""
    keep = [_c(d['m0']), _cov_factor(d['P0']), _cov_factor(d['q_cov']), _cov_factor(d['r_cov'])]
    r = _lib.SsmRng()
    r.seed, r.traj_offset = int(seed) & 0xFFFFFFFFFFFFFFFF, int(traj_offset)
    r.x0_mean, r.x0_factor, r.q_factor, r.r_factor = (_ptr(a) for a in keep)
    if d.get('sample_student', False):   # per random variable; 0 (absent) = Gaussian draws with the factor of its covariance
        r.x0_dof, r.q_dof, r.r_dof = float(d.get('x0_dof', 0.0)), float(d.get('q_dof', 0.0)), float(d.get('r_dof', 0.0))
    r.dq = keep[2].shape[0]
    return r, keep


def simulate(low, n_traj, n_steps, rng=None, mode='discrete', dt=0.0, sub=1, x0=None, q=None, r=None,
             want_y=True, device='cuda'):
    """Simulate states (dx, N, M) and measurements (dy, N, M) with ssm_simulate.  rng = (SsmRng, keep)
    from make_rng, or injected noise tensors x0 (dx, M), q (dq, Nq, M), r (dy, N, M)."""
    kw = dict(dtype=torch.float64, device=device)
    x = torch.empty((low.dx, n_steps, n_traj), **kw)
    y = torch.empty((low.dy, n_steps, n_traj), **kw) if want_y else None
    m = {'discrete': _lib.SIM_DISCRETE, 'continuous': _lib.SIM_CONTINUOUS}[mode]
    for t in (x0, q, r):
        if t is not None and (t.dtype != torch.float64 or not t.is_contiguous() or t.shape[-1] != n_traj):
            raise ValueError('injected noise must be contiguous float64 with trajectory axis last')
    rc = lib.ssm_simulate(C.byref(low.desc), C.byref(rng[0]) if rng is not None else None, m, float(dt), int(sub),
                          _p(x0), _p(q), _p(r), _p(x), _p(y), n_traj, n_steps, n_traj, _stream())
    _lib.check(rc, 'ssm_simulate')
    return x, y


def simulate_measurements(low, x, rng=None, r=None):
    """Measurements of a given state array x (dx, N, M) (MeasurementModel.simulate_measurements)."""
    dx, N, M = x.shape
    y = torch.empty((low.dy, N, M), dtype=torch.float64, device=x.device)
    rc = lib.ssm_simulate(C.byref(low.desc), C.byref(rng[0]) if rng is not None else None, _lib.SIM_MEASURE, 0.0, 1,
                          None, None, _p(r), _p(x.contiguous()), _p(y), M, N, M, _stream())
    _lib.check(rc, 'ssm_simulate')
    return y


# ------------------------------------------------------------------------------------------------
# K5: Bayesian-quadrature weights
# ------------------------------------------------------------------------------------------------
def bq_weights(par, points, mulind=None, device='cuda', precision='dd', to_host=True):
    """Batched BQ weights (ssm_bq_weights).  par (n_par, D+1), points (D, N), mulind (D, Q) or None.
    Returns dict of numpy arrays: wm (n_par, N), Wc (n_par, N, N), Wcc (n_par, D, N), iK (n_par, N, N),
    model_var (n_par,), integral_var (n_par,), info (n_par,).
    precision 'dd' (default): double-double arithmetic, rounded once -> correctly rounded weights even for the
    cond(K) ~ 1e9 kernels of the reference's tracking scripts; 'float64': plain float64 like the reference."""
    par = _c(np.atleast_2d(par))
    points = _c(points)
    D, N = points.shape
    n_par = par.shape[0]
    if par.shape[1] != D + 1:
        raise ValueError('kernel parameters must have D+1 = {} columns'.format(D + 1))
    kw = dict(dtype=torch.float64, device=device)
    wm, Wc, Wcc = torch.empty((n_par, N), **kw), torch.empty((n_par, N, N), **kw), torch.empty((n_par, D, N), **kw)
    iK, scal = torch.empty((n_par, N, N), **kw), torch.empty((n_par, 2), **kw)
    info = torch.empty((n_par,), dtype=torch.int32, device=device)
    mi, nb = None, 0
    if mulind is not None:
        mi = np.ascontiguousarray(np.asarray(mulind, dtype=np.int32))
        if mi.shape[0] != D:
            raise ValueError('Dimension mismatch {:d} != {:d}. Dimension of monomials must be equal to the dimension'
                             ' of the sigma-points.'.format(mi.shape[0], D))
        nb = mi.shape[1]
    rc = lib.ssm_bq_weights(D, N, n_par, _ptr(par), _ptr(points),
                            mi.ctypes.data_as(_lib.c_int32_p) if mi is not None else None, nb,
                            _p(wm), _p(Wc), _p(Wcc), _p(iK), _p(scal), _p(info), 0 if precision == 'float64' else 1,
                            _stream())
    _lib.check(rc, 'ssm_bq_weights')
    if to_host is False:   # device tensors in the layout ssm_transform_apply_batched reads
        return dict(wm=wm, Wc=Wc, Wcc=Wcc, iK=iK, model_var=scal[:, 0].contiguous(), integral_var=scal[:, 1].contiguous(), info=info)
    sc = scal.cpu().numpy()
    return dict(wm=wm.cpu().numpy(), Wc=Wc.cpu().numpy(), Wcc=Wcc.cpu().numpy(), iK=iK.cpu().numpy(),
                model_var=sc[:, 0].copy(), integral_var=sc[:, 1].copy(), info=info.cpu().numpy())


def own_weights(d, symmetric=True):
    """A description dict (keys of tests/golden/*.npz, see lower()) with its BQ / TPQ weights re-derived from the kernel
    parameters by ssm_bq_weights ('dd'), and -- symmetric=True, what the facade does by default -- projected onto their
    exact reflection structure (bq/bqmod.symmetrize_reflective), so that lower() + filter_forward() take the compact sums
    of the forward pass.  Sigma-point transforms are left as they are."""
    from .bq import bqmod
    out = dict(d)
    for pfx in ('dyn_', 'obs_'):
        if str(d[pfx + 'kind']) == 'sp':
            continue
        mul = d[pfx + 'mulind'] if (pfx + 'mulind') in d else None
        w = bq_weights(d[pfx + 'kern_par'], d[pfx + 'points'], mul)
        if int(w['info'][0]) != 0:
            raise np.linalg.LinAlgError('kernel matrix is not positive definite')
        w1 = {k: w[k][0] for k in ('wm', 'Wc', 'Wcc', 'iK')}
        if symmetric:
            w1 = bqmod.symmetrize_reflective(d[pfx + 'points'], w1)
        for k in ('wm', 'Wc', 'Wcc'):
            out[pfx + k] = w1[k]
        if (pfx + 'iK') in d:
            out[pfx + 'iK'] = w1['iK']
        if str(d[pfx + 'kind']) != 'tp':
            out[pfx + 'model_var'] = w['model_var'][0]
    return out


def weights_reflective(low):
    """(dynamics, measurement): does the forward pass take the compact reflection-symmetric sums for this transform?
    (ssm_weights_reflective; host-side check)"""
    return (bool(lib.ssm_weights_reflective(C.byref(low.desc.tf_dyn))), bool(lib.ssm_weights_reflective(C.byref(low.desc.tf_obs))))


def transform_apply_batched(which, model_id, dim_state, si, par, points, w, time, mean, cov):
    """GPQ / BSQ moment transform of n (mean, cov) columns with one weight set per column (ssm_transform_apply_batched).
    w: dict of device tensors from bq_weights(..., to_host=False) with n parameter vectors; mean (D, n), cov (D, D, n)
    device tensors.  Returns mean_f (E, n), cov_f (E, E, n), cov_fx (E, D, n), status (n,) on the device."""
    points = _c(points)
    D, N = points.shape
    n = mean.shape[1]
    E = {0: dim_state}.get(which)
    if which == 1:
        E = int(w['_dim_out'])
    kw = dict(dtype=torch.float64, device=mean.device)
    mf, cf, cfx = torch.empty((E, n), **kw), torch.empty((E, E, n), **kw), torch.empty((E, D, n), **kw)
    status = torch.empty((n,), dtype=torch.int32, device=mean.device)
    p8 = (C.c_double * 8)(*(list(par) + [0.0] * (8 - len(par))))
    rc = lib.ssm_transform_apply_batched(which, model_id, dim_state, si[0], si[1], p8, N, _ptr(points), _p(w['wm']), _p(w['Wc']),
                                         _p(w['Wcc']), _p(w['model_var']), float(time), _p(mean.contiguous()), _p(cov.contiguous()),
                                         _p(mf), _p(cf), _p(cfx), _p(status), n, n, _stream())
    _lib.check(rc, 'ssm_transform_apply_batched')
    return mf, cf, cfx, status


# ------------------------------------------------------------------------------------------------
# K6: scores
# ------------------------------------------------------------------------------------------------
def _check_score_layout(x, mean, cov, quad):
    """The score kernels are launched with ld = M: every bulk argument must be a dense (..., N, M) array."""
    M = x.shape[-1]
    for name, t in (('x', x), ('mean', mean), ('cov', cov), ('quad', quad)):
        if t is not None and M > 1 and bulk_ld(t) != M:
            raise ValueError('scores: {} must be C-contiguous (trajectory-range views are not supported)'.format(name))


def scores_phase1(x, mean, cov, status=None, want_rmse_acc=True, window=None, out=None, nll_acc=None, quad=None):
    """Per-step packed statistics over trajectories: stats (N, W), W = dx + dx*dx + 3:
    [sum SE | sum d d^T | sum NLL | sum |d| | count]; rmse_acc (dx, M) per-trajectory time-sums of SE.
    window = (k_lo, k_hi) fills only those rows of out = (stats, rmse_acc) (walk the windows first to last).
    nll_acc (M,): optional per-trajectory time-sum of the NLL (continued, not reset, when k_lo > 0).
    quad (N, M): optional output, d' P^-1 d of every scored unit, for scores_phase2(..., quad=quad)."""
    dx, N, M = x.shape
    _check_score_layout(x, mean, cov, quad)
    W = lib.ssm_scores_width(dx)
    if out is not None:
        stats, acc = out
    else:
        stats = torch.empty((N, W), dtype=torch.float64, device=x.device)
        acc = torch.empty((dx, M), dtype=torch.float64, device=x.device) if want_rmse_acc else None
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    rc = lib.ssm_scores_phase1_quad(dx, _p(x), _p(mean), _p(cov), _p(status), _p(stats), _p(acc), _p(nll_acc), _p(quad), M, N, k_lo, k_hi, M, _stream())
    _lib.check(rc, 'ssm_scores_phase1')
    return stats, acc


def scores_phase2(x, mean, cov, mse, status=None, window=None, out=None, lcr_acc=None, quad=None):
    """Per-step sums of the log credibility ratio and of its absolute value: (N, 2).  mse (dx, dx, N).
    window = (k_lo, k_hi) fills only those rows of out (N, 2) and reads only those columns of mse.
    lcr_acc (M,): optional per-trajectory time-sum of the ratio (continued, not reset, when k_lo > 0)."""
    dx, N, M = x.shape
    _check_score_layout(x, mean, cov if quad is None else None, quad)
    lcr = out if out is not None else torch.empty((N, 2), dtype=torch.float64, device=x.device)
    k_lo, k_hi = (0, N) if window is None else (int(window[0]), int(window[1]))
    if quad is not None:   # d' P^-1 d stored by smooth_backward(..., want_quad=True): the covariances are not read (cov may be None)
        rc = lib.ssm_scores_phase2_quad(dx, _p(x), _p(mean), _p(quad), _p(status), _p(mse.contiguous()), _p(lcr), _p(lcr_acc), M, N, k_lo, k_hi, M, _stream())
        _lib.check(rc, 'ssm_scores_phase2')
        return lcr
    rc = lib.ssm_scores_phase2_traj(dx, _p(x), _p(mean), _p(cov), _p(status), _p(mse.contiguous()), _p(lcr), _p(lcr_acc), M, N, k_lo, k_hi, M, _stream())
    _lib.check(rc, 'ssm_scores_phase2')
    return lcr


def bootstrap_var(data, samples=1000, seed=0):
    """Bootstrap variance of the mean of `data` (any shape, squeezed to 1-D like utils.py:236), on the device:
    returns a 0-d device tensor.  Deterministic for a given seed (Philox keyed by (seed, resample index))."""
    d = data.reshape(-1).contiguous()
    means = torch.empty(int(samples), dtype=torch.float64, device=d.device)
    var = torch.empty(1, dtype=torch.float64, device=d.device)
    rc = lib.ssm_bootstrap_var(_p(d), d.numel(), int(samples), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(means), _p(var), _stream())
    _lib.check(rc, 'ssm_bootstrap_var')
    return var[0]
