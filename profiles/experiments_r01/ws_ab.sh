#!/bin/bash
# A/B of the warp-specialised row-block kernels (PICARD_RB_WS=1, default) against the unspecialised ones (PICARD_RB_WS=0).
out=gpurun_out/ws_ab.jsonl; : > $out
for s in 0 1; do
  echo "{\"ws\": $s}" >> $out
  PICARD_RB_WS=$s timeout -s KILL 60 python profiles/pass_bench.py 128 1e7 5 0 loss,gradY,gradY+H >> $out 2>&1
  PICARD_RB_WS=$s timeout -s KILL 60 python profiles/pass_bench.py 256 2e6 5 1 loss,gradY,gradY+H >> $out 2>&1
  PICARD_RB_WS=$s timeout -s KILL 60 python profiles/pass_bench.py 100 1e6 5 0 loss,gradY >> $out 2>&1
done
cat $out
