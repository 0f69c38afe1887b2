"""Developer tool / profiles: host -> device copy bandwidth with N ranks copying AT THE SAME TIME (torchrun), the ceiling of
the end-to-end path of bench.py (every rank streams 3.5 GB of pinned y, x per step).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe_multi.py

Per rank: one pinned buffer of the benchmark's size; (a) one contiguous cudaMemcpyAsync, (b) the windowed strided copies
of mc.filter_scores (20 windows, cudaMemcpy2DAsync, one stream), (c) the same split over two copy streams.  Prints the
slowest rank's time and the aggregate GB/s."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv      # noqa: E402


def main():
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('nccl')
    M, N, C_ = 125000, 500, 7                      # y (2) + x (5) components
    h = torch.empty((C_, N, M), dtype=torch.float64, pin_memory=True)
    h.fill_(1.0)
    d = torch.empty((C_, N, M), dtype=torch.float64, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    nbytes = h.numel() * 8
    wins = [(a, min(a + 25, N)) for a in range(0, N, 25)]

    def window(k0, k1, stream):
        rc = dv.lib.ssm_memcpy2d(C.c_void_p(d.data_ptr() + k0 * M * 8), N * M * 8, C.c_void_p(h.data_ptr() + k0 * M * 8), N * M * 8,
                                 (k1 - k0) * M * 8, C_, 1, C.c_void_p(stream.cuda_stream))
        assert rc == 0

    def contiguous():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)

    def windowed_one():
        for k0, k1 in wins:
            window(k0, k1, s1)

    def windowed_two():
        for i, (k0, k1) in enumerate(wins):
            window(k0, k1, s1 if i % 2 == 0 else s2)

    for name, fn in (('contiguous', contiguous), ('windowed x20, 1 stream', windowed_one), ('windowed x20, 2 streams', windowed_two)):
        best = 1e30
        for _ in range(4):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0), s2.wait_event(e0)
            fn()
            torch.cuda.current_stream().wait_stream(s1), torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        if rank == 0:
            print('ranks=%d  %-24s slowest rank %.1f ms   %.1f GB/s per rank   %.1f GB/s aggregate' % (
                world, name, best, nbytes / best / 1e6, world * nbytes / best / 1e6), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
