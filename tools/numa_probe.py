"""Developer tool: does the placement of the pinned host buffers (NUMA node of the allocating thread) explain the
63 ms vs 90-96 ms spread of the e2e step between boxes?  Prints the topology, then times a 1 GB pinned H2D copy with the
buffer allocated (a) where the process happens to run, (b) on each NUMA node in turn (thread pinned to that node's
cores before cudaHostAlloc + first touch), (c) after nvmlDeviceSetCpuAffinity (the GPU's own node)."""
import glob
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return 'failed: %r' % (e,)


def cpulist(s):
    out = []
    for part in s.strip().split(','):
        if not part:
            continue
        a, _, b = part.partition('-')
        out += list(range(int(a), int(b or a) + 1))
    return out


def h2d_gbs(nbytes=1 << 30, reps=5):
    h = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    h.fill_(1.0)
    d = torch.empty_like(h, device='cuda')
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d.copy_(h, non_blocking=True)
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    del h, d
    torch.cuda.empty_cache()
    torch._C._host_emptyCache() if hasattr(torch._C, '_host_emptyCache') else None
    return nbytes / best / 1e6


def main():
    torch.cuda.set_device(0)
    torch.zeros(1, device='cuda')
    print(sh('nvidia-smi topo -m'))
    print(sh('lscpu | grep -i -E "numa|socket|model name|^cpu\\(s\\)"'))
    allowed = sorted(os.sched_getaffinity(0))
    print('allowed cpus: %d  (%s ... %s)' % (len(allowed), allowed[:4], allowed[-4:]))
    nodes = {}
    for p in sorted(glob.glob('/sys/devices/system/node/node*/cpulist')):
        n = int(p.split('node')[-1].split('/')[0])
        nodes[n] = [c for c in cpulist(open(p).read()) if c in allowed]
    print('numa nodes (allowed cpus per node):', {n: len(c) for n, c in nodes.items()})
    bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), 'pci_bus_id') else None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
    except Exception as e:  # noqa: BLE001
        print('nvml failed', e)
        h = None
    if bus:
        short = bus.lower()[-12:]
        print('gpu0 pci', bus, 'numa_node:', sh('cat /sys/bus/pci/devices/%s/numa_node' % short),
              'local_cpulist:', sh('cat /sys/bus/pci/devices/%s/local_cpulist' % short))
    print('default placement: %.1f GB/s (running on cpu %s)' % (h2d_gbs(), sh('cat /proc/%d/stat | cut -d" " -f39' % os.getpid())))
    for n, cpus in nodes.items():
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        print('thread on node %d: %.1f GB/s' % (n, h2d_gbs()))
    os.sched_setaffinity(0, allowed)
    if h is not None:
        try:
            pynvml.nvmlDeviceSetCpuAffinity(h)
            print('after nvmlDeviceSetCpuAffinity (%d cpus): %.1f GB/s' % (len(os.sched_getaffinity(0)), h2d_gbs()))
        except Exception as e:  # noqa: BLE001
            print('nvmlDeviceSetCpuAffinity failed', e)
    os.sched_setaffinity(0, allowed)


if __name__ == '__main__':
    main()
