#!/bin/bash
# round 2, GPU call q: evidence for profiles/ on the current build -- launch list of the bench command, ncu --set full of the two pass kernels
# captured inside the same command (each only after the command has exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout -s KILL 300 $CMD > gpurun_out/r02q_plain.json 2> gpurun_out/r02q_plain.err || exit 1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02q.csv $CMD > gpurun_out/r02q_ncu_list.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'loss_i8_kernel|grad_i8_kernel' --launch-skip 6 -c 4 -o gpurun_out/prof_r02q $CMD > gpurun_out/r02q_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02q.ncu-rep gpurun_out/launches_r02q.csv
tail -n 2 gpurun_out/r02q_ncu_full.log
exit 0
