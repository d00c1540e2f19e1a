#!/bin/bash
# round 2, GPU call u: device block cache keeps the N x T buffers -- whole -m gpu suite and the default bench line
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02u_pytest.log 2>&1; echo "pytest exit $?"
tail -n 4 gpurun_out/r02u_pytest.log
timeout -s KILL 900 python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench exit $?"
grep -v "whiten\|eigh\|centering" gpurun_out/r02u_bench.err | tail -n 48
exit 0
