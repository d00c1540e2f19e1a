#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_i8_gpu.py -x -q -m gpu > gpurun_out/r02i8t_pytest.log 2>&1; echo "pytest exit $?"
tail -n 5 gpurun_out/r02i8t_pytest.log
exit 0
