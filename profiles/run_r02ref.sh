#!/bin/bash
# round 2: the reference arm of bench.py on the GPU box's host cores (short: 1 warm-up + 2 timed steps)
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02ref.json 2> gpurun_out/r02ref.err; echo "reference exit $?"
cut -c1-900 gpurun_out/r02ref.json; tail -n 3 gpurun_out/r02ref.err
exit 0
