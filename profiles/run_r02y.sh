#!/bin/bash
# round 2, GPU call y: evidence for profiles/ on the current build -- launch list of the bench command, ncu --set full of the two pass kernels
# captured inside the same command (each only after the command has exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout -s KILL 300 $CMD > gpurun_out/r02y_plain.json 2> gpurun_out/r02y_plain.err || exit 1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02y.csv $CMD > gpurun_out/r02y_ncu_list.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'loss_i8_kernel|grad_i8_kernel' --launch-skip 6 -c 4 -o gpurun_out/prof_r02y $CMD > gpurun_out/r02y_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02y.ncu-rep gpurun_out/launches_r02y.csv
tail -n 2 gpurun_out/r02y_ncu_full.log
exit 0
