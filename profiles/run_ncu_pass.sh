#!/bin/bash
# ncu --set full on the pass kernels alone (profiles/pass_bench.py).  Usage: bash profiles/run_ncu_pass.sh <tag> [N] [T]
set -u
TAG=${1:-pass}; N=${2:-128}; T=${3:-2e6}
CMD="python profiles/pass_bench.py $N $T 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -c 8 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
