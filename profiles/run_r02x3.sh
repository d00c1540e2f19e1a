#!/bin/bash
# round 2, GPU call x3: ncu --set full of the N x N kernels between the passes (transform series, iteration front)
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'expm_multi_kernel|front_cluster_kernel' --launch-skip 6 -c 4 -o gpurun_out/prof_r02x3 $CMD > gpurun_out/r02x3_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02x3.ncu-rep
tail -n 2 gpurun_out/r02x3_ncu_full.log
exit 0
