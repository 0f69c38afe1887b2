"""Developer tool / profiles: BASELINE configuration C5 -- BSQ NCI calibration sweep, trajectory count 10^3 .. 10^7, on
1 GPU (python tools/time_c5.py) or N GPUs (torchrun --nproc-per-node N tools/time_c5.py).  Prints one table row per
(model, trajectory count): seconds for the whole sweep point (simulate -> BSQ filter with in-kernel scoring -> second
score phase, nothing materialised but 8 (dx + 1) bytes per unit), trajectory-steps/s and the roofline fractions of
SURVEY.md section 8(d) (BQ filter FLOP per step: pendulum 510, coordinated turn 5668)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssmtoybox_b200 import device as dv                      # noqa: E402
from ssmtoybox_b200.dist import Communicator                  # noqa: E402
from ssmtoybox_b200.research import bsq_nci_sweep as sw       # noqa: E402

FLOP = {'pendulum': 510.0, 'coordturn': 5668.0}
BYTES = {'pendulum': 8.0 * (2 + 1 + 3), 'coordturn': 8.0 * (5 + 2 + 6)}    # write+read x, y; write d, quad (read again by phase 2)


def main():
    comm = Communicator.from_env()
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    counts = [int(float(a)) for a in sys.argv[1:]] or [10 ** 3, 10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7]
    peak = dv.fp64_peak() if comm.rank == 0 else 0.0
    rows = []
    for model in ('pendulum', 'coordturn'):
        sw.bsq_nci_sweep(model, mc_sims=(2000,), model_var=(1e-2,), comm=comm)          # warm-up (module load, pools)
        for M in counts:
            r = sw.bsq_nci_sweep(model, mc_sims=(M,), model_var=(1e-2,), comm=comm, chunk=1 << 19)[0]
            r['fp64_frac'] = r['traj_steps_per_s'] * FLOP[model] / peak / comm.world_size if peak else None
            r['n_gpus'] = comm.world_size
            rows.append(r)
            if comm.rank == 0:
                print('{model:10s} gpus={n_gpus} M={mc_sims:<9d} {seconds:8.4f} s  {traj_steps_per_s:10.3e} traj-steps/s  FP64 {f:5.1f} %  NCI {nci:+8.4f}  '
                      'failed {n_failed}  kept {kb:.1f} MB/rank'.format(f=100 * (r['fp64_frac'] or 0), kb=r['kept_bytes'] / 1e6, **r), flush=True)
    if comm.rank == 0:
        print(json.dumps(rows))


if __name__ == '__main__':
    main()
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
