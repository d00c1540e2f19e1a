#!/bin/bash
# round 2, GPU call f: LOSS v3 (64-sample tiles) + gradient converter v2 (uniform work) -- lab, INT8 tests, bench with parity block
mkdir -p gpurun_out
timeout -s KILL 120 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02f_lab.jsonl 2> gpurun_out/r02f_lab.err
echo "lab exit $?" >> gpurun_out/r02f_lab.err
timeout -s KILL 600 python -m pytest tests/test_i8_gpu.py -q -m gpu > gpurun_out/r02f_pytest_i8.log 2>&1
echo "pytest i8 exit $?" >> gpurun_out/r02f_pytest_i8.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
echo "bench exit $?" >> gpurun_out/r02f_bench.err
for f in gpurun_out/r02f_lab.err gpurun_out/r02f_pytest_i8.log gpurun_out/r02f_bench.err; do echo "== $f"; tail -n 4 $f; done
head -c 600 gpurun_out/r02f_bench.json
exit 0
