"""Multi-GPU plumbing: one process per GPU, trajectories sharded in contiguous blocks, no data-path
collective.  The only exchange of the Monte-Carlo hot path is the all-reduce of the packed error
statistics (a few hundred kB of fp64, latency-bound): one call for RMSE / MSE / NLL and one more,
N doubles, when the non-credibility index is requested, because the log credibility ratio needs the
GLOBAL per-step MSE matrix (research/gpq/icinco_demo.py:34-40, utils.py:113-120).

torch.distributed is used for the rendezvous and the collective (backend nccl on CUDA tensors over
NVLink / NVSwitch, gloo on CPU tensors in the tests).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous block of ceil(n_total / world_size) trajectories per rank -> (offset, count).
    Philox draws are keyed by the GLOBAL trajectory index (offset + local index), so the simulated
    data do not depend on the number of ranks."""
    per = -(-int(n_total) // int(world_size))
    off = min(rank * per, n_total)
    return off, max(0, min(per, n_total - off))


class Communicator:
    """Thin wrapper over an (optional) torch.distributed process group."""

    def __init__(self, group=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.active else 0
        self.world_size = dist.get_world_size(group) if self.active else 1

    @classmethod
    def from_env(cls, backend=None):
        """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
        ws = int(os.environ.get('WORLD_SIZE', '1'))
        if ws > 1 and not dist.is_initialized():
            if backend is None:
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            if backend == 'nccl':
                torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
            dist.init_process_group(backend=backend)
        return cls()

    def shard(self, n_total):
        return shard_range(n_total, self.rank, self.world_size)

    def allreduce_sum(self, t):
        """In-place sum over ranks of a tensor (fp64 statistics); returns the tensor."""
        if self.active and self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_max(self, value):
        """Max over ranks of a python float (device timings are reported as the max over ranks)."""
        if not (self.active and self.world_size > 1):
            return float(value)
        dev = 'cuda' if dist.get_backend(self.group) == 'nccl' else 'cpu'
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        if self.active and self.world_size > 1:
            dist.barrier(group=self.group)


def finalize_scores(stats, rm, lcr, dx, n_steps):
    """Turn globally reduced packed statistics into scores (host-side arithmetic on N x W numbers),
    following research/gpq/icinco_demo.py:17-52.  stats (N, W) = [sum SE | sum d d^T | sum NLL |
    sum |d| | count], rm (dx,) = sum over trajectories of sqrt(time-mean SE), lcr (N, 2) or None."""
    N = n_steps
    cnt = stats[:, -1]
    n_ok = cnt[0]
    out = {'rmse': rm / n_ok, 'nll': stats[1:, dx + dx * dx].sum() / (N * n_ok),
           'mse': (stats[:, dx:dx + dx * dx] / cnt[:, None]).T.reshape(dx, dx, N),
           'rmse_vs_time': stats[:, dx + dx * dx + 1] / cnt, 'n_ok': n_ok}
    if lcr is not None:
        out['nci'] = lcr[1:, 0].sum() / (N * n_ok)
        out['abs_nci'] = lcr[1:, 1].sum() / (N * n_ok)
    return out
