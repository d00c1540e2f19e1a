#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02o_lab.jsonl 2> gpurun_out/r02o_lab.err
echo "lab exit $?" >> gpurun_out/r02o_lab.err
timeout -s KILL 600 python -m pytest tests/test_i8_gpu.py -q -m gpu > gpurun_out/r02o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02o_pytest.log
for f in gpurun_out/r02o_lab.err gpurun_out/r02o_pytest.log; do echo "== $f"; tail -n 4 $f; done
grep -E "kernel|trace" gpurun_out/r02o_lab.jsonl | cut -c1-200
exit 0
