#!/bin/bash
# round 2, GPU call c: LOSS epilogue v2 (product log, per-warp stores) -- lab, INT8 tests, then ncu --set full with source of both kernels
mkdir -p gpurun_out
timeout -s KILL 120 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02c_lab.jsonl 2> gpurun_out/r02c_lab.err
echo "lab exit $?" >> gpurun_out/r02c_lab.err
timeout -s KILL 600 python -m pytest tests/test_i8_gpu.py -q -m gpu > gpurun_out/r02c_pytest_i8.log 2>&1
echo "pytest i8 exit $?" >> gpurun_out/r02c_pytest_i8.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:loss_i8_kernel -c 1 --launch-skip 2 -o gpurun_out/r02c_loss profiles/lab/i8_lab 2000000 1 1 > gpurun_out/r02c_ncu_loss.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:grad_i8_kernel -c 1 --launch-skip 2 -o gpurun_out/r02c_grad profiles/lab/i8_lab 2000000 1 1 > gpurun_out/r02c_ncu_grad.log 2>&1
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/r02c_lab.err gpurun_out/r02c_pytest_i8.log gpurun_out/r02c_ncu_loss.log gpurun_out/r02c_ncu_grad.log; do echo "== $f"; tail -n 3 $f; done
exit 0
