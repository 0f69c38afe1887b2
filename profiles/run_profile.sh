#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md, run under gpurun (1 GPU).
#   1. launch list of the bench command (per-launch device time; cold-cache, serialised: compare SHARES)
#   2. `--set full` capture of the dominant kernel (fused forward pass) at the bench trajectory count
#      (125 000 trajectories -> persistent grid + ticket scheduler, 12 warps/SM; 100 steps keep the replay short):
#      the headline instantiation (own weights: compact reflection-symmetric sums, lower-triangle stores as in the score
#      pipeline) and the dense-sum one (reference weights, full stores)
#   3. `--set full` capture of the score-only smoother (+ in-kernel scores) from a reduced bench run
#   4. `--set full` capture of the scoring forward pass (configuration C5: coordinated turn BSQ, in-kernel scores)
# Each ncu run follows a plain run of the same command that exited 0.
set -x
mkdir -p gpurun_out
OUT=gpurun_out
TAG=${TAG:-r2j}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-c5"
$CMD > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python tools/one_launch.py 125000 100 c3_reentry_gpq:own predlow"
$CMD2 > $OUT/plain_filter.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o $OUT/prof_${TAG}_filter $CMD2 > $OUT/ncu_filter.log 2>&1
echo "filter capture rc=$?"
CMD2D="python tools/one_launch.py 125000 100 c3_reentry_gpq pred"
$CMD2D > $OUT/plain_filter_dense.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o $OUT/prof_${TAG}_filterdense $CMD2D > $OUT/ncu_filter_dense.log 2>&1
echo "dense filter capture rc=$?"
CMD3="python bench.py --steps 1 --warmup 3 --no-cpu --no-c5 --traj 37888"
$CMD3 > $OUT/plain_smoother.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:smoother_kernel -s 3 -c 1 -o $OUT/prof_${TAG}_smoother $CMD3 > $OUT/ncu_smoother.log 2>&1
echo "smoother capture rc=$?"
CMD4="python tools/one_launch.py 121952 100 c4_ct_bsq:own scored"
$CMD4 > $OUT/plain_scored.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o $OUT/prof_${TAG}_scored $CMD4 > $OUT/ncu_scored.log 2>&1
echo "scored filter capture rc=$?"
for k in filter filterdense smoother scored; do
  u=390625; [ $k = smoother ] && u=592000; [ $k = scored ] && u=381100
  python tools/ncu_summary.py $OUT/prof_${TAG}_$k.ncu-rep $u > $OUT/ncu_${k}_${TAG}_summary.txt 2>&1
done
ls -la $OUT
