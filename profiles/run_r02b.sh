#!/bin/bash
# round 2, GPU call b: gradient kernel under compute-sanitizer first, then lab, INT8 tests, the rest of the suite, a short bench
mkdir -p gpurun_out
timeout -s KILL 240 compute-sanitizer --tool memcheck --print-limit 5 profiles/lab/i8_lab 30000 1 > gpurun_out/r02b_sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/r02b_sanitizer.log
timeout -s KILL 120 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02b_lab.jsonl 2> gpurun_out/r02b_lab.err
echo "lab exit $?" >> gpurun_out/r02b_lab.err
timeout -s KILL 900 python -m pytest tests/test_i8_gpu.py -q -m gpu > gpurun_out/r02b_pytest_i8.log 2>&1
echo "pytest i8 exit $?" >> gpurun_out/r02b_pytest_i8.log
timeout -s KILL 1200 python -m pytest tests -q -x -m gpu --deselect tests/test_i8_gpu.py > gpurun_out/r02b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench exit $?" >> gpurun_out/r02b_bench.err
for f in gpurun_out/r02b_sanitizer.log gpurun_out/r02b_lab.err gpurun_out/r02b_pytest_i8.log gpurun_out/r02b_pytest.log gpurun_out/r02b_bench.err; do echo "== $f"; tail -n 3 $f; done
exit 0
