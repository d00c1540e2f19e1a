#!/bin/bash
# round 2, GPU call t: after the stage-release fix -- the whole -m gpu suite, repeatability of the engines at full size, smoke(), bench
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02t_pytest.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/r02t_pytest.log
timeout -s KILL 900 python profiles/lab/parity_repeat.py 1e7 6 > gpurun_out/r02t_repeat.jsonl 2> gpurun_out/r02t_repeat.err; echo "repeat exit $?"
cut -c1-250 gpurun_out/r02t_repeat.jsonl
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02t_smoke.log 2>&1; echo "smoke exit $?"
timeout -s KILL 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench exit $?"
grep -v "whiten\|eigh\|centering" gpurun_out/r02t_bench.err | tail -n 40
exit 0
