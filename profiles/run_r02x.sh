#!/bin/bash
# round 2, GPU call x: host hand-over gaps of the core loop (PICARD_TRACE_GAPS)
mkdir -p gpurun_out
PICARD_TRACE_GAPS=1 timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench exit $?"
tail -n 8 gpurun_out/r02x_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['non_pass_ms_per_step'])"
exit 0
