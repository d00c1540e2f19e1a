#!/bin/bash
# round 2, GPU call r: the whole -m gpu suite, smoke(), the default bench line on the current build
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/r02r_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke exit $?"
tail -n 3 gpurun_out/r02r_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench exit $?"
tail -n 5 gpurun_out/r02r_bench.err
exit 0
