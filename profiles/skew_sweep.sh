#!/bin/bash
# A/B of the pair turn-taking switch (PICARD_RB_SKEW: 0 off, 1 on) on the row-block kernels: one JSON line per pass and value.
out=gpurun_out/skew_sweep.jsonl; : > $out
for s in ${@:-0 1}; do
  echo "{\"skew\": $s}" >> $out
  PICARD_RB_SKEW=$s python profiles/pass_bench.py 128 1e7 5 0 loss,gradY >> $out 2>&1
  PICARD_RB_SKEW=$s python profiles/pass_bench.py 256 2e6 5 1 loss,gradY >> $out 2>&1
  PICARD_RB_SKEW=$s python profiles/pass_bench.py 64 4e6 5 0 loss,gradY >> $out 2>&1
done
cat $out
