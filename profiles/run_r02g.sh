#!/bin/bash
# round 2, GPU call g: MMA rate probe (N up to 256), merged-digit MMAs in both INT8 kernels (lab + tests), whitening refinement tests
mkdir -p gpurun_out
timeout -s KILL 120 profiles/lab/pipe_probe > gpurun_out/r02g_pipe_probe.jsonl 2>&1
timeout -s KILL 120 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02g_lab.jsonl 2> gpurun_out/r02g_lab.err
echo "lab exit $?" >> gpurun_out/r02g_lab.err
timeout -s KILL 600 python -m pytest tests/test_i8_gpu.py tests/test_whiten_gpu.py -q -m gpu > gpurun_out/r02g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02g_pytest.log
for f in gpurun_out/r02g_lab.err gpurun_out/r02g_pytest.log; do echo "== $f"; tail -n 6 $f; done
cut -c1-250 gpurun_out/r02g_pipe_probe.jsonl
cut -c1-200 gpurun_out/r02g_lab.jsonl
exit 0
