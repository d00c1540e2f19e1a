#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:loss_i8_kernel -c 1 --launch-skip 2 -o gpurun_out/r02h_loss profiles/lab/i8_lab 2000000 1 1 > gpurun_out/r02h_ncu_loss.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:grad_i8_kernel -c 1 --launch-skip 2 -o gpurun_out/r02h_grad profiles/lab/i8_lab 2000000 1 1 > gpurun_out/r02h_ncu_grad.log 2>&1
ls -la gpurun_out/r02h*.ncu-rep
exit 0
