#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one --set full capture of
# the pass kernels.  Usage under gpurun:  bash profiles/run_ncu.sh <tag> [workload]
set -u
TAG=${1:-r01}
WL=${2:-c3}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:rb_|loss_i8' -s 2 -c 6 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full_$TAG.log
