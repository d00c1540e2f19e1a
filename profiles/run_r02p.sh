#!/bin/bash
# round 2, GPU call p: full GPU suite + bench (1 GPU) on the current build
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/r02p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02p_pytest.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
echo "bench exit $?" >> gpurun_out/r02p_bench.err
for f in gpurun_out/r02p_pytest.log gpurun_out/r02p_bench.err; do echo "== $f"; tail -n 4 $f; done
head -c 300 gpurun_out/r02p_bench.json
exit 0
