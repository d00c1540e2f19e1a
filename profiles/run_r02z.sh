#!/bin/bash
# round 2, GPU call z: ncu --set full of the INT8 gradient kernel inside the bench command (the capture of call y took LOSS launches only)
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout -s KILL 300 $CMD > gpurun_out/r02z_plain.json 2> gpurun_out/r02z_plain.err || exit 1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'grad_i8_kernel' --launch-skip 3 -c 2 -o gpurun_out/prof_r02z $CMD > gpurun_out/r02z_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02z.ncu-rep
tail -n 2 gpurun_out/r02z_ncu_full.log
exit 0
