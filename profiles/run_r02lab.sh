#!/bin/bash
# round 2, GPU call lab: LOSS kernel ablations incl. "no MMAs" (what the epilogue alone costs) at T = 1e7
mkdir -p gpurun_out
timeout -s KILL 300 profiles/lab/i8_lab 1e7 5 > gpurun_out/r02lab.jsonl 2> gpurun_out/r02lab.err; echo "lab exit $?"
grep '"loss_i8"\|trace' gpurun_out/r02lab.jsonl | cut -c1-200
exit 0
