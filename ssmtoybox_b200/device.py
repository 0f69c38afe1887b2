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
    GQG = _c(G.dot(_c(d['q_cov'])).dot(G.T))  # ssinf.py:279
    keep += [m0, P0, R, GQG]
    s = _lib.SsmDesc()
    s.dyn_model, s.obs_model, s.dx, s.dy = _lib.DYN_IDS[dyn_name], _lib.OBS_IDS[obs_name], dx, dy
    s.dyn_par[0] = float(d.get('dyn_dt', 0.0))
    rl = np.asarray(d.get('radar_loc', [0.0, 0.0]), dtype=np.float64).reshape(-1)
    s.obs_par[0], s.obs_par[1] = float(rl[0]), float(rl[1])
    si = np.asarray(d.get('state_index', []), dtype=np.int64).reshape(-1)
    s.n_state_index = len(si)
    for i, v in enumerate(si):
        s.state_index[i] = int(v)
    s.m0, s.P0, s.GQG, s.R = _ptr(m0), _ptr(P0), _ptr(GQG), _ptr(R)
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


def filter_forward(low, y, store_pred=False, store_cov=True, init_mean=None, init_cov=None, t_offset=None, k0=0,
                   want_last=False, out=None):
    """Run the fused forward pass (ssm_filter) on y (dy, N, M) -> dict of device tensors."""
    dy, N, M = y.shape
    dx = low.dx
    dev = y.device
    _check_bulk(y, (dy,), N, M, 'y')
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
    if want_last:
        o['last_mean'] = torch.empty((dx, M), **kw)
        o['last_cov'] = torch.empty((dx, dx, M), **kw)
    if 'status' not in o:
        o['status'] = torch.empty((M,), dtype=torch.int32, device=dev)
    if init_mean is not None:
        init_mean = init_mean.contiguous()
        init_cov = init_cov.contiguous()
    rc = lib.ssm_filter(C.byref(low.desc), _p(y), _p(o.get('fi_mean')), _p(o.get('fi_cov')), _p(o.get('pr_mean')),
                        _p(o.get('pr_cov')), _p(o.get('pr_xx_cov')), _p(init_mean), _p(init_cov),
                        _p(o.get('last_mean')), _p(o.get('last_cov')), _p(t_offset), int(k0), _p(o['status']),
                        M, N, M, _stream())
    _lib.check(rc, 'ssm_filter')
    return o


def smooth_backward(dx, fwd, out=None):
    """Run the RTS smoother (ssm_smooth) over the arrays stored by filter_forward(store_pred=True)."""
    _, N, M = fwd['fi_mean'].shape
    o = out if out is not None else {}
    if 'sm_mean' not in o:
        o['sm_mean'] = torch.empty_like(fwd['fi_mean'])
        o['sm_cov'] = torch.empty_like(fwd['fi_cov'])
    o['status'] = fwd['status'].clone()
    rc = lib.ssm_smooth(dx, _p(fwd['fi_mean']), _p(fwd['fi_cov']), _p(fwd['pr_mean']), _p(fwd['pr_cov']),
                        _p(fwd['pr_xx_cov']), _p(o['sm_mean']), _p(o['sm_cov']), _p(o['status']), M, N, M, _stream())
    _lib.check(rc, 'ssm_smooth')
    return o


def fp64_peak(n_blocks=148 * 8, n_iters=4096, reps=5):
    """Measured FP64 FMA throughput (FLOP/s) of the device: the roofline denominator for the
    FP64-bound kernels (MEASURED_PEAKS.json carries HBM and bf16 only)."""
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
