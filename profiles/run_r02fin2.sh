#!/bin/bash
# round 2, last verification (1 GPU): whole -m gpu suite and smoke() on the final tree
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02fin2_pytest.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/r02fin2_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02fin2_smoke.log 2>&1; echo "smoke exit $?"
tail -n 2 gpurun_out/r02fin2_smoke.log
exit 0
