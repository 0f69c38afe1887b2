#!/usr/bin/env python
"""Benchmark of the hot path on the configuration BASELINE.json's metric is quoted on.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config C3): 5-D reentry vehicle + radar, GaussianProcessKalman (RBF GPQ,
UT points, the reference's hyper-parameters; quadrature weights: see --weights) forward pass + RTS smoother + error scores,
10^6 trajectories x 500 steps on 8 GPUs = 125 000 trajectories x 500 steps PER GPU (weak scaling:
trajectories are independent, each rank owns a contiguous block, the only collective is the all-reduce
of the packed error statistics).  One "step" = one pass of that hot path over the rank's batch.

  value : filtered trajectory-steps/s, whole job, measurements and truth resident in HBM (CUDA events,
          max over ranks)
  e2e   : the same through the reference-facing Python API with HOST buffers: pinned y / x copied to the
          device, forward_pass + backward_pass + evaluate_performance, scores read back -- all inside the
          timed region
  roofline : dominant kernel = the fused forward pass (FP64-pipe bound); achieved = algorithmic FLOPs per
          launch / its CUDA-event duration; peak = FP64 FMA rate measured in this run by a DFMA
          micro-kernel (MEASURED_PEAKS.json carries HBM and bf16 only)
  cpu_baseline : the UNMODIFIED reference (pip-installed copy under baseline/_ref, imported through oracle/ref_shim.py;
          per-trajectory forward_pass + backward_pass + reset like research/gpq/icinco_demo.py:120-124) on all host
          cores, bounded sample of the same workload; the numpy port of oracle/ only if that copy cannot be imported
          (`kind` says which ran)
  c5    : BASELINE configuration C5 next to the headline: one BSQ NCI sweep point per model (10^6 x 100 per GPU,
          simulate -> filter with in-kernel scoring -> second score phase, nothing materialised)

--impl reference times the reference's own CPU implementation of the path the same way (data from the reference's own
simulators, not timed); --weights own (default) runs the headline with the quadrature weights the public constructor builds,
--weights reference with the golden reference run's weights assigned, and either way the other set is timed next to it in the
same process (`weights_other`); --config c5 makes the C5 sweep point the timed workload of the line.
"""
import argparse
import json
import multiprocessing
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STEPS = 500
TRAJ_PER_GPU = 125000           # 10^6 / 8
FLOP_FILTER = 5756.0            # algorithmic FLOP per trajectory-step, BQ filter, reentry N=11
FLOP_SMOOTH = 902.0
BYTES_FILTER = 8.0 * (2 + 5 + 15 + 5 + 15 + 25)      # read y; write fi_mean, tril(fi_cov), pr_mean, tril(pr_cov), pr_xx_cov (ssm_filter_window_lower)
BYTES_SMOOTH = 8.0 * (5 + 15 + 25 + 5 + 15 + 5 + 5 + 1)  # read pr_mean, tril(pr_cov), pr_xx, fi_mean, tril(fi_cov), x; write d = x - m_s, quad (score-only mode)
WEIGHTS_DESC = {'own': "own: ssm_bq_weights in double-double, projected onto the reflection structure the formulas have in exact "
                       "arithmetic (what GaussianProcessKalman(...) builds by default)",
                'reference': "reference-injected: wm / Wc / Wcc of the reference's own run assigned from tests/golden/c3_reentry_gpq.npz "
                             "(float64 rounding noise included)"}
METRIC = 'filtered trajectory-steps/sec (fp64)'
UNIT = 'trajectory-steps/s'
CONFIG = {'workload': 'C3: reentry 5-D + radar, GPQ (RBF, UT) filter + RTS smoother + scores, '
                      '125000 trajectories x 500 steps per GPU (10^6 x 500 on 8 GPUs)',
          'n_traj_per_gpu': TRAJ_PER_GPU, 'n_steps': N_STEPS,
          'l2': 'inputs and outputs per step (>= 1 GB) are far larger than the 126 MB L2, no explicit flush'}


def golden_c3():
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'c3_reentry_gpq.npz')))


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (baseline/_ref or /root/reference under oracle/ref_shim.py) on all host cores,
# per-trajectory loop exactly like research/gpq/icinco_demo.py:120-124; the oracle port if it cannot be imported
# ------------------------------------------------------------------------------------------------
def _reference_filter(g):
    """The C3 filter built from the REFERENCE's classes (research/gpq/gpq_tracking.py:41-44 on the model of
    research/bsq/bsq_tracking.py:230-261), with the weights of the golden run assigned like bench.build_filter does."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_shim
    if ref_shim.available() is None:
        raise ImportError('no copy of the reference (baseline/_ref, /root/reference)')
    ref_shim.install()
    from ssmtoybox import ssinf, ssmod
    from ssmtoybox.utils import GaussRV
    m0 = np.array([6500, 350, -1.1, -6.1, 0.7])
    dyn = ssmod.ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 1])),
                                           GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 1e-6])), dt=0.1)
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([6374, 0.0]))
    alg = ssinf.GaussianProcessKalman(dyn, obs, g['dyn_kern_par'], g['obs_kern_par'], kernel='rbf', points='ut')
    for tf, p in ((alg.tf_dyn, 'dyn_'), (alg.tf_obs, 'obs_')):
        tf.wm, tf.Wc, tf.Wcc = g[p + 'wm'], g[p + 'Wc'], g[p + 'Wcc']
        tf.model.model_var = float(g[p + 'model_var'])
    return alg, ssmod, GaussRV


def _reference_data(ssmod, GaussRV, n, seed):
    """Truth and measurements from the reference's own simulators (bsq_tracking.py:230-254): Euler-Maruyama at
    dt = 0.05 for 50 s, every second state -> 500 steps."""
    np.random.seed(seed)
    sysm = ssmod.ReentryVehicle2DTransition(GaussRV(5, np.array([6500, 350, -1.8, -6.8, 0.7]), np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0])),
                                            GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 0])))
    obs = ssmod.Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([6374, 0.0]))
    x = sysm.simulate_continuous(duration=N_STEPS * 0.1, dt=0.05, mc_sims=n)
    y = obs.simulate_measurements(x)
    return np.ascontiguousarray(y[:, ::2][:, :N_STEPS])


def _cpu_worker(args):
    """(kind, golden, y or None, n, seed) -> (seconds in forward + backward passes, trajectories completed)."""
    os.environ['OPENBLAS_NUM_THREADS'] = os.environ['OMP_NUM_THREADS'] = '1'
    kind, g, y, n, seed = args
    if kind == 'reference':
        import warnings
        warnings.simplefilter('ignore')
        alg, ssmod, GaussRV = _reference_filter(g)
        if y is None:
            y = _reference_data(ssmod, GaussRV, n, seed)
        t0 = time.perf_counter()
        ok = 0
        for i in range(y.shape[2]):      # research/gpq/icinco_demo.py:120-124
            try:
                alg.forward_pass(y[..., i])
                alg.backward_pass()
                ok += 1
            except (np.linalg.LinAlgError, ValueError):
                pass
            alg.reset()
        return time.perf_counter() - t0, ok
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ssm_oracle as so
    t0 = time.perf_counter()
    fw = so.forward_pass(g, y, backend='lapack')
    so.backward_pass(g, fw, backend='lapack')
    return time.perf_counter() - t0, int((fw['status'] == 0).sum())


def reference_kind():
    """'reference' when the unmodified reference can be imported on this machine, else 'port'."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    try:
        import ref_shim
        return 'reference' if ref_shim.available() else 'port'
    except Exception:
        return 'port'


def cpu_rate(n_traj_per_core=2, cores=None, seed=0, y=None):
    """trajectory-steps/s of the reference's per-trajectory loop (forward + backward pass) on `cores` processes.
    y (2, 500, >= cores * n): measurements of the benchmark itself (the GPU arm passes its Philox data); None: the
    reference arm simulates its own with the reference's simulators (not timed)."""
    g = golden_c3()
    cores = cores or len(os.sched_getaffinity(0))
    kind = reference_kind()
    gl = {k: v for k, v in g.items() if not k.startswith(('fi_', 'pr_', 'sm_')) and k not in ('x', 'y')}
    if y is None and kind == 'port':     # the port has no simulator dependency on the reference: perturbed golden measurements
        rng = np.random.RandomState(seed)
        base = g['y'][:, :, [0] * (cores * n_traj_per_core)]
        y = np.ascontiguousarray(base + rng.randn(*base.shape) * np.sqrt(np.diag(g['r_cov']))[:, None, None] * 0.1)
    jobs = [(kind, gl, None if y is None else np.ascontiguousarray(y[:, :, c * n_traj_per_core:(c + 1) * n_traj_per_core]),
             n_traj_per_core, 1000 * seed + c) for c in range(cores)]
    warm = [(kind, gl, None if y is None else j[2][..., :1], 1, 999983 + c) for c, j in enumerate(jobs)]
    with multiprocessing.get_context('spawn').Pool(cores) as pool:
        pool.map(_cpu_worker, warm)                                   # warm-up: imports, one trajectory each
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    if y is None:   # data generation of the reference arm happens inside the workers: count only the filter time
        wall = max(r[0] for r in res)
    n = cores * n_traj_per_core * N_STEPS
    what = 'the unmodified reference (ssmtoybox v0.1.1a0 under oracle/ref_shim.py), GaussianProcessKalman forward_pass + backward_pass + reset per trajectory' \
        if kind == 'reference' else 'oracle numpy port (per-trajectory loop like the reference), forward + backward pass'
    return n / wall, cores, kind, '{}, {} trajectories x {} steps on {} processes, {:.1f} s, {} completed'.format(
        what, cores * n_traj_per_core, N_STEPS, cores, wall, sum(r[1] for r in res))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals = []
    sample, kind, cores = '', 'port', 1
    per_core = 16
    for i in range(args.warmup + args.steps):
        v, cores, kind, sample = cpu_rate(n_traj_per_core=per_core, seed=i)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * cores * per_core * N_STEPS / value, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': CONFIG,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.  Primary: an NVML polling thread inside this
    process (no start-up latency; the handle is looked up by the PCI bus id of the CUDA device, so
    CUDA_VISIBLE_DEVICES re-mappings do not matter).  Fallback: `nvidia-smi -lms 20` as a child process (the recipe's
    clocks line) -- on a busy 8-GPU box its first sample can arrive after a 170 ms timed region has ended."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.p, self.f, self.thread, self.rows = None, None, None, []
        try:
            import threading
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
            self.stop_flag = threading.Event()

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        r = int(get_reasons(h))
                        self.rows.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, [k for k, b in bits.items() if r & b]))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '20'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            sm = [r[0] for r in self.rows]
            return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.max_mhz,
                    'power_w_max': max([r[1] for r in self.rows] or [0.0]), 'samples': len(self.rows),
                    'reasons': sorted({k for r in self.rows for k in r[2]}), 'source': 'nvml'}
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.strip().lower() == 'active':
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(rows[0][1]) if rows else None,
                'power_w_max': max([float(r[2]) for r in rows if r[2].replace('.', '').isdigit()] or [0.0]),
                'samples': len(rows), 'reasons': sorted(reasons), 'source': 'nvidia-smi'}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_filter(weights='reference'):
    """The C3 filter through the reference-facing API (research/gpq/gpq_tracking.py:41-44 on the model of
    research/bsq/bsq_tracking.py:230-261).  weights='reference': the reference's own weights are assigned from outside
    (the pattern of research/tpq/tpq_ungm.py:114-124): its obs-transform kernel matrix has cond 1e9, so its covariance
    weights are LAPACK rounding noise that no independent evaluation reproduces (DESIGN.md,
    tests/test_gpu_weights_envelope.py).  weights='own' (the bench default): the weights the package computes itself
    (double-double, the correctly rounded values of the reference's formulas, projected onto their exact reflection
    structure -> compact sums in the forward pass)."""
    from ssmtoybox_b200.ssinf import GaussianProcessKalman
    from ssmtoybox_b200.ssmod import ReentryVehicle2DTransition, Radar2DMeasurement
    from ssmtoybox_b200.utils import GaussRV
    g = golden_c3()
    m0 = np.array([6500, 350, -1.1, -6.1, 0.7])
    dyn = ReentryVehicle2DTransition(GaussRV(5, m0, np.diag([1e-6, 1e-6, 1e-6, 1e-6, 1])),
                                     GaussRV(3, cov=np.diag([2.4e-5, 2.4e-5, 1e-6])), dt=0.1)
    obs = Radar2DMeasurement(GaussRV(2, cov=np.diag([1e-6, 0.17e-6])), 5, radar_loc=np.array([6374, 0.0]))
    alg = GaussianProcessKalman(dyn, obs, g['dyn_kern_par'], g['obs_kern_par'], kernel='rbf', points='ut')
    if weights == 'reference':
        for tf, p in ((alg.tf_dyn, 'dyn_'), (alg.tf_obs, 'obs_')):
            tf.wm, tf.Wc, tf.Wcc = g[p + 'wm'], g[p + 'Wc'], g[p + 'Wcc']
            tf.model.model_var = float(g[p + 'model_var'])
    return alg, g


def run_c5_arm(args):
    """--config c5: BSQ NCI calibration sweep (BASELINE configuration C5) as the timed workload: per step one sweep
    point of research.bsq_nci_sweep = simulate -> BayesSardKalman forward pass with in-kernel scoring -> second score
    phase on args.traj trajectories x 100 steps per GPU, nothing materialised but 8 (dx + 1) bytes per unit."""
    import torch
    from ssmtoybox_b200 import device as dv
    from ssmtoybox_b200.dist import Communicator
    from ssmtoybox_b200.research import bsq_nci_sweep as sw
    comm = Communicator.from_env()
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    M = args.traj * comm.world_size
    out = {}
    flop = {'pendulum': 510.0, 'coordturn': 5668.0}
    peak = dv.fp64_peak()
    clocks = ClockSampler(local_rank) if comm.rank == 0 else None
    for model in ('pendulum', 'coordturn'):
        for _ in range(max(args.warmup, 3)):
            sw.bsq_nci_sweep(model, mc_sims=(M,), model_var=(1e-2,), comm=comm, chunk=1 << 19)
        rows = [sw.bsq_nci_sweep(model, mc_sims=(M,), model_var=(1e-2,), comm=comm, chunk=1 << 19)[0] for _ in range(args.steps)]
        sec = float(np.mean([r['seconds'] for r in rows]))
        out[model] = {'ms_per_step': 1e3 * sec, 'value': M * sw.N_STEPS / sec, 'nci': rows[-1]['nci'], 'rmse': rows[-1]['rmse'],
                      'n_failed': rows[-1]['n_failed'], 'fp64_frac': M * sw.N_STEPS / sec * flop[model] / peak / comm.world_size,
                      'kept_bytes_per_rank': rows[-1]['kept_bytes']}
    clk = clocks.stop() if clocks else None
    if comm.rank != 0:
        return
    ct = out['coordturn']
    line = {'metric': METRIC, 'value': ct['value'], 'unit': UNIT, 'n_gpus': comm.world_size, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ct['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic (Philox, simulated in the timed region)',
            'config': {'workload': 'C5: BayesSardKalman NCI calibration sweep point, coordinated turn 5-D + radar (value) and pendulum 2-D, '
                                   '{} trajectories x {} steps per GPU, simulate -> filter with in-kernel scoring -> second score phase'.format(args.traj, sw.N_STEPS),
                       'n_traj_per_gpu': args.traj, 'n_steps': sw.N_STEPS, 'l2': 'no bulk arrays: per unit 8 (dx + 1) bytes are kept'},
            'c5': out, 'roofline': {'bound': 'fp64', 'achieved': ct['value'] * flop['coordturn'] / 1e12 / comm.world_size, 'peak': peak / 1e12, 'unit': 'TFLOP/s',
                                    'frac': ct['fp64_frac'], 'traffic': None, 'flop_per_unit': flop['coordturn']},
            'gpu_launches': 8 * args.steps, 'clocks': clk}
    _emit(line)


def run_gpu_arm(args):
    import torch
    from ssmtoybox_b200 import device as dv, utils as U
    from ssmtoybox_b200.dist import Communicator
    comm = Communicator.from_env()
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    M, N = args.traj, N_STEPS
    alg, g = build_filter(args.weights)
    low = dv.lower(alg._describe())

    # ---- synthetic truth and measurements: Euler-Maruyama at dt = 0.05, every 2nd state (bsq_tracking.py:248-254)
    truth = {'m0': [6500, 350, -1.8, -6.8, 0.7], 'P0': np.diag([1e-6, 1e-6, 1e-6, 1e-6, 0.0]),
             'q_cov': np.diag([2.4e-5, 2.4e-5, 0.0]), 'r_cov': g['r_cov']}
    off = comm.rank * M                                     # weak scaling: every rank owns M trajectories
    x, y = dv.simulate(low, M, N, rng=dv.make_rng(truth, seed=2026, traj_offset=off), mode='continuous', dt=0.05, sub=2,
                       device=dev)
    fwd, sm = {}, {}
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def hot_path(timers=None, low=low):
        """forward pass (stores predictive moments) -> RTS smoother with in-kernel score accumulation ->
        all-reduce of the packed statistics -> second score phase (log credibility ratio) -> all-reduce."""
        e = [ev() for _ in range(4)] if timers is not None else None
        if e: e[0].record()
        dv.filter_forward(low, y, store_pred=True, out=fwd, lower_only=True)   # the score-only smoother reads lower triangles only
        if e: e[1].record()
        dv.smooth_scores(low.dx, fwd, x, out=sm)   # RTS smoother, score-only mode: in-kernel phase-1 statistics, no sm_* stores
        if e: e[2].record()
        sc = U.evaluate_scored(sm, comm=comm, to_host=False)
        if e:
            e[3].record()
            timers.append(e)
        return sc

    for _ in range(max(args.warmup, 3)):
        sc = hot_path()
    torch.cuda.synchronize()
    # The first process on a fresh node runs the HBM-bound smoother up to 25 % slower for its first steps (round-1
    # SCALE N=1 leg: 39.0 against 34.2 ms): keep warming up until three consecutive steps agree to 2 % (at most 15 more)
    n_warm, recent = max(args.warmup, 3), []
    for _ in range(15):
        a0, a1 = ev(), ev()
        a0.record()
        sc = hot_path()
        a1.record()
        a1.synchronize()
        n_warm += 1
        recent = (recent + [a0.elapsed_time(a1)])[-3:]
        if len(recent) == 3 and max(recent) - min(recent) <= 0.02 * min(recent):
            break
    n_failed = int((sm['status'] != 0).sum().item())

    # ---- value: device-resident inputs ---------------------------------------------------------
    clocks = ClockSampler(local_rank) if comm.rank == 0 else None
    timers = []
    comm.barrier()
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        sc = hot_path(timers)
    t1.record()
    torch.cuda.synchronize()
    comm.barrier()
    ms_total = comm.allreduce_max(t0.elapsed_time(t1))
    clk = clocks.stop() if clocks else None
    ms_step = ms_total / args.steps
    value = comm.world_size * M * N / (ms_step * 1e-3)
    k_filter = float(np.mean([e[0].elapsed_time(e[1]) for e in timers]))
    k_smooth = float(np.mean([e[1].elapsed_time(e[2]) for e in timers]))
    k_scores = float(np.mean([e[2].elapsed_time(e[3]) for e in timers]))
    scores = {k: (v.tolist() if hasattr(v, 'tolist') else v) for k, v in sc.items() if k in ('rmse', 'nci', 'nll', 'n_ok')}

    # ---- the same device-resident step with the OTHER weight set, next to the headline ------------------------------
    # (own weights carry their exact reflection structure -> compact sums of the forward pass; weights assigned from a
    # reference run carry its rounding noise instead -> the dense sums of bqmtran.py:175-223 as they stand)
    other = 'reference' if args.weights == 'own' else 'own'
    alg_o, _ = build_filter(other)
    low_o = dv.lower(alg_o._describe())
    for _ in range(3):
        hot_path(low=low_o)
    timers_o = []
    comm.barrier()
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        sc_o = hot_path(timers_o, low=low_o)
    t1.record()
    torch.cuda.synchronize()
    comm.barrier()
    ms_other = comm.allreduce_max(t0.elapsed_time(t1)) / args.steps
    weights_other = {'weights': WEIGHTS_DESC[other], 'compact_sums': list(dv.weights_reflective(low_o)), 'ms_per_step': ms_other,
                     'value': comm.world_size * M * N / (ms_other * 1e-3),
                     'kernel_ms': {'filter_forward': float(np.mean([e[0].elapsed_time(e[1]) for e in timers_o])),
                                   'rts_smoother_with_phase1_scores': float(np.mean([e[1].elapsed_time(e[2]) for e in timers_o])),
                                   'scores_phase2_incl_allreduce': float(np.mean([e[2].elapsed_time(e[3]) for e in timers_o]))},
                     'n_failed_trajectories': int((sm['status'] != 0).sum().item()),
                     'scores': {k: (v.tolist() if hasattr(v, 'tolist') else v) for k, v in sc_o.items() if k in ('rmse', 'nci', 'nll', 'n_ok')}}
    compact = list(dv.weights_reflective(low))
    hot_path()   # leave the headline filter's moments in fwd / sm

    # ---- e2e: host buffers through the reference-facing API --------------------------------------
    yh = torch.empty(y.shape, dtype=torch.float64, pin_memory=True).copy_(y)
    xh = torch.empty(x.shape, dtype=torch.float64, pin_memory=True).copy_(x)
    del fwd, sm
    torch.cuda.empty_cache()

    from ssmtoybox_b200 import mc

    def e2e_step():
        """The user-level call: filter + smoother + scores of one Monte-Carlo batch held in HOST memory.  Inside:
        time-windowed H2D of y (forward) and x (backward) overlapped with the kernels (ssm_filter_window,
        ssm_smooth_window, ssm_scores_phase2_window), scores copied back to the host."""
        return mc.filter_scores(alg, yh, xh, smooth=True, n_windows=args.windows, comm=comm)

    for _ in range(max(1, min(args.warmup, 2))):
        out = e2e_step()
    # like the device-resident path: keep warming up until three consecutive steps agree to 3 % (at most 8 more; every rank
    # takes the same number of steps -- the decision is all-reduced).  The pinned host -> device copy, 63 of the step's
    # ~65 ms, fluctuates from step to step on some boxes (67 / 72 / 102 ms in three consecutive steps of one run; 66 / 82 /
    # 96 ms per step as the mean of runs on three boxes, same code): the last three warm-up times go into the line
    # (`warmup_last_ms`) so that a noisy box is visible as such.
    e2e_warm, recent = max(1, min(args.warmup, 2)), []
    for _ in range(8):
        torch.cuda.synchronize()
        w = time.perf_counter()
        out = e2e_step()
        torch.cuda.synchronize()
        recent = (recent + [comm.allreduce_max((time.perf_counter() - w) * 1e3)])[-3:]
        e2e_warm += 1
        if len(recent) == 3 and max(recent) - min(recent) <= 0.03 * min(recent):
            break
    comm.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        out = e2e_step()
    t1.record()
    torch.cuda.synchronize()
    comm.barrier()
    e2e_ms = comm.allreduce_max(t0.elapsed_time(t1)) / args.steps
    e2e_wall_ms = comm.allreduce_max((time.perf_counter() - w0) * 1e3) / args.steps
    e2e_value = comm.world_size * M * N / (max(e2e_ms, e2e_wall_ms) * 1e-3)
    d2h = 8 * (5 + 4 + 25 * N + N + 1) + 4 * M

    h2d_bytes = int(yh.numel() + xh.numel()) * 8
    y_cpu = yh[:, :, :64 * 48].clone().numpy() if (comm.world_size == 1 and not args.no_cpu) else None
    # ---- configuration C5 next to the headline: one BSQ NCI sweep point per model, 10^6 x 100 per GPU, nothing
    # materialised (simulate -> filter with in-kernel scoring -> second score phase); `python bench.py --config c5` times
    # it as the workload of its own line
    c5 = None
    if not args.no_c5:
        del yh, xh
        torch.cuda.empty_cache()
        from ssmtoybox_b200.research import bsq_nci_sweep as sw
        c5 = {}
        for model, flop in (('pendulum', 510.0), ('coordturn', 5668.0)):
            Mc = 10 ** 6 * comm.world_size
            sw.bsq_nci_sweep(model, mc_sims=(Mc,), model_var=(1e-2,), comm=comm, chunk=1 << 19)
            r = sw.bsq_nci_sweep(model, mc_sims=(Mc,), model_var=(1e-2,), comm=comm, chunk=1 << 19)[0]
            c5[model] = {'n_traj': Mc, 'n_steps': sw.N_STEPS, 'ms': 1e3 * r['seconds'], 'value': r['traj_steps_per_s'], 'unit': UNIT,
                         'nci': r['nci'], 'n_failed': r['n_failed'], 'flop_per_unit': flop}
    if comm.rank != 0:
        return
    # ---- roofline of the dominant kernel -----------------------------------------------------------
    fp64_peak = dv.fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    if c5:
        for v in c5.values():
            v['fp64_frac'] = v['value'] * v['flop_per_unit'] / fp64_peak / comm.world_size
    ach_tf = M * N * FLOP_FILTER / (k_filter * 1e-3) / 1e12
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except (OSError, ValueError):
        pass
    kname = 'BQR (compact reflection-symmetric sums)' if all(compact) else 'BQ (dense sums)'
    exe = prof.get('filter_kernel_executed_flop_per_unit_bqr' if all(compact) else 'filter_kernel_executed_flop_per_unit_bq')
    roofline = {'kernel': 'filter_kernel<DynReentry, ObsRadar<5,0,1>, AXIS_C, 11, %s, GAUSS> (fused forward pass)' % kname,
                'bound': 'fp64', 'achieved': ach_tf, 'peak': fp64_peak / 1e12, 'unit': 'TFLOP/s', 'frac': ach_tf / (fp64_peak / 1e12),
                # frac counts the ALGORITHMIC flops of the reference's dense sums per unit of time; the kernel
                # executes fewer (half-row and reflection-symmetric forms of the same sums): FP64 instructions actually executed
                # per unit from the ncu capture (DFMA = 2), and the share of the FP64 pipe they occupy
                'executed_flop_per_unit': exe,
                'frac_executed': (M * N * exe / (k_filter * 1e-3) / fp64_peak) if exe else None,
                'peak_source': 'measured in this run: DFMA micro-kernel ssm_fp64_peak_kernel (MEASURED_PEAKS.json has no fp64 figure)',
                'flop_per_unit': FLOP_FILTER, 'units_per_launch': M * N, 'launch_ms': k_filter,
                'traffic': prof.get('filter_kernel_dram_bytes_per_launch'),
                'hbm': {'achieved': M * N * BYTES_FILTER / (k_filter * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': M * N * BYTES_FILTER / (k_filter * 1e-3) / 1e9 / hbm_peak, 'bytes_per_unit': BYTES_FILTER,
                        'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback 6650 GB/s'}}
    roofline_smoother = {'kernel': 'smoother_kernel<5, SCORE, KEEP=false> (RTS smoother, score-only mode: in-kernel phase-1 score accumulation, no smoothed arrays stored)', 'bound': 'hbm',
                         'achieved': M * N * BYTES_SMOOTH / (k_smooth * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': M * N * BYTES_SMOOTH / (k_smooth * 1e-3) / 1e9 / hbm_peak, 'bytes_per_unit': BYTES_SMOOTH,
                         'launch_ms': k_smooth, 'traffic': prof.get('smoother_kernel_dram_bytes_per_launch')}
    cpu = None
    if comm.world_size == 1 and not args.no_cpu:
        cores = len(os.sched_getaffinity(0))
        per_core = 48 if reference_kind() == 'reference' else 32
        cores = min(cores, 64)
        v, cores, kind, sample = cpu_rate(n_traj_per_core=per_core, cores=cores, y=y_cpu[:, :, :cores * per_core])
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
               'sample': sample + '; measurements = the first trajectories of this benchmark run (Philox)'}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': comm.world_size, 'steps': args.steps, 'warmup': n_warm,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic (Philox Euler-Maruyama truth + radar measurements, seed 2026, keyed by global trajectory index)',
            'config': dict(CONFIG, n_traj_per_gpu=M),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': max(e2e_ms, e2e_wall_ms),
                    'api': 'ssmtoybox_b200.mc.filter_scores(GaussianProcessKalman, y, x, smooth=True) on pinned host y, x: '
                           'time-windowed H2D (y forward in time, x backward) overlapped with forward pass + RTS smoother + '
                           'scores of the windows that have landed; scores and status read back',
                    'windows': args.windows, 'warmup_steps': e2e_warm, 'warmup_last_ms': recent},
            'gpu_launches': 7 * args.steps,   # filter, NaN fill of failed trajectories, smoother, finalize, MSE factor table, scores phase 2, finalize
            'kernel_ms': {'filter_forward': k_filter, 'rts_smoother_with_phase1_scores': k_smooth, 'scores_phase2_incl_allreduce': k_scores},
            'filter_only_value': comm.world_size * M * N / (k_filter * 1e-3),
            'roofline': roofline, 'roofline_smoother': roofline_smoother, 'cpu_baseline': cpu,
            'clocks': clk, 'n_failed_trajectories': n_failed, 'scores': scores,
            'c5': c5,
            'weights': WEIGHTS_DESC[args.weights], 'compact_sums': compact, 'weights_other': weights_other,
            'warmup_steps_until_stable': n_warm,
            'parity': 'means 1e-9 per step against the unmodified reference on identical inputs, weights included (golden runs with the '
                      "reference's own weights -> dense sums; golden runs with structured weights assigned -> the compact sums of this "
                      "headline: tests/golden/c3_reentry_gpq_structured.npz); un-centred BQ covariances on this model to the reference's own "
                      'float64 noise floor (<= 4x the reference\'s error against a longdouble evaluation: tests/test_gpu_parity.py, '
                      'tests/test_gpu_reflective.py)'}
    _emit(line)


def _emit(line):
    """Print THE one JSON line on the real stdout (fd 1 is pointed at stderr while the job runs so that
    library banners such as 'NCCL version ...' cannot pollute it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--traj', type=int, default=TRAJ_PER_GPU, help='trajectories per GPU (default: the C3 share)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--no-c5', action='store_true', help='skip the configuration-C5 sweep points reported next to the headline')
    ap.add_argument('--windows', type=int, default=20, help='time windows of the host-streaming (e2e) pipeline')
    ap.add_argument('--weights', default='own', choices=['reference', 'own'],
                    help="quadrature weights of the C3 filter: the package's own (default: what the public constructor builds) or "
                         "the reference's values assigned from its golden run; the other set is timed next to it (weights_other)")
    ap.add_argument('--config', default='c3', choices=['c3', 'c5'], help='c3: the headline workload; c5: BSQ NCI sweep point')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    elif args.config == 'c5':
        if args.traj == TRAJ_PER_GPU:
            args.traj = 10 ** 6
        run_c5_arm(args)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    else:
        run_gpu_arm(args)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
