#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md, run under gpurun (1 GPU).
#   1. launch list of the bench command (per-launch device time; cold-cache, serialised: compare SHARES)
#   2. `--set full` capture of the dominant kernel (fused forward pass) at the bench trajectory count
#      (125 000 trajectories -> persistent grid + ticket scheduler, 12 warps/SM; 100 steps keep the replay short)
#   3. `--set full` capture of the smoother (+ in-kernel scores) from a reduced bench run
# Each ncu run follows a plain run of the same command that exited 0.
set -x
mkdir -p gpurun_out
OUT=gpurun_out
TAG=${TAG:-r1c}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python tools/one_launch.py 125000 100 c3_reentry_gpq pred"
$CMD2 > $OUT/plain_filter.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o $OUT/prof_${TAG}_filter $CMD2 > $OUT/ncu_filter.log 2>&1
echo "filter capture rc=$?"
CMD3="python bench.py --steps 1 --warmup 3 --no-cpu --traj 37888"
$CMD3 > $OUT/plain_smoother.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:smoother_kernel -s 3 -c 1 -o $OUT/prof_${TAG}_smoother $CMD3 > $OUT/ncu_smoother.log 2>&1
echo "smoother capture rc=$?"
ls -la $OUT
