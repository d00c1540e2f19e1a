#!/bin/bash
# round 2, GPU call v2 (2 GPUs): multi-GPU tests on the build with the cluster front / new transform kernels
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > gpurun_out/r02v2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02v2_pytest.log
tail -n 4 gpurun_out/r02v2_pytest.log
exit 0
