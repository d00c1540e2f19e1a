#!/bin/bash
# launch list only (ncu --metrics gpu__time_duration.sum).  Usage: bash profiles/run_list.sh <tag> [workload]
TAG=${1:-r01}; WL=${2:-c3}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
