#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md, run under gpurun (1 GPU).
#   1. launch list of the bench command (per-launch device time; cold-cache, serialised: compare SHARES)
#   2. one `--set full` capture of the dominant kernel (fused forward pass) and of the smoother
# Each ncu run follows a plain run of the same command that exited 0.
set -x
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r1.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --traj 37888"
$CMD2 > $OUT/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_kernel|smoother_kernel' -s 6 -c 2 -o $OUT/prof_r1 $CMD2 > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la $OUT
