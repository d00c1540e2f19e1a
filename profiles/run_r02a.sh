#!/bin/bash
# round 2, GPU call a: new INT8 engines -- lab (layout check, timings, ablations, trace), the GPU test suite, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout -s KILL 300 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02a_lab.jsonl 2> gpurun_out/r02a_lab.err
echo "lab exit $?" >> gpurun_out/r02a_lab.err
timeout -s KILL 900 python -m pytest tests/test_i8_gpu.py -x -q -m gpu > gpurun_out/r02a_pytest_i8.log 2>&1
echo "pytest i8 exit $?" >> gpurun_out/r02a_pytest_i8.log
timeout -s KILL 1200 python -m pytest tests -q -m gpu --deselect tests/test_i8_gpu.py > gpurun_out/r02a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?" >> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_lab.err gpurun_out/r02a_pytest_i8.log gpurun_out/r02a_pytest.log gpurun_out/r02a_bench.err
