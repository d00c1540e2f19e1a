#!/bin/bash
# 1/2/4/8-GPU strong-scaling run of bench.py (c3).  Usage under gpurun --gpus 8: bash profiles/run_scaling.sh <tag>
TAG=${1:-r01}
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --no-cpu > gpurun_out/scale_${TAG}_g$N.json 2> gpurun_out/scale_${TAG}_g$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N \
      > gpurun_out/scale_${TAG}_g$N.json 2> gpurun_out/scale_${TAG}_g$N.err
  fi
  echo "N=$N rc=$?"
done
