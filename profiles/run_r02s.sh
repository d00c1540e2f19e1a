#!/bin/bash
# round 2, GPU call s: full-size repeatability of the pass kernels (lab mode 2)
mkdir -p gpurun_out
timeout -s KILL 600 profiles/lab/i8_lab 1e7 60 2 > gpurun_out/r02s_lab.jsonl 2> gpurun_out/r02s_lab.err; echo "lab exit $?"
grep -v '"kernel"' gpurun_out/r02s_lab.jsonl | grep -v 'differing_from_first": 0' | cut -c1-400 | tail -n 40
exit 0
