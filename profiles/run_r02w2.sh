#!/bin/bash
# round 2, GPU call w2 (8 GPUs): c3 strong scaling at 8 and 4 GPUs after the stage-release fix and with the block cache
mkdir -p gpurun_out
export PICARD_TRACE=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02w2_c3_g8.json 2> gpurun_out/r02w2_c3_g8.err
echo "c3 g8 exit $?" >> gpurun_out/r02w2_c3_g8.err
timeout -s KILL 600 $TR --nproc-per-node 4 --master-port 29642 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu --no-parity > gpurun_out/r02w2_c3_g4.json 2> gpurun_out/r02w2_c3_g4.err
echo "c3 g4 exit $?" >> gpurun_out/r02w2_c3_g4.err
timeout -s KILL 600 $TR --nproc-per-node 2 --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-parity --no-e2e > gpurun_out/r02w2_c3_g2.json 2> gpurun_out/r02w2_c3_g2.err
head -c 200 gpurun_out/r02w2_c3_g2.json; echo
for f in gpurun_out/r02w2_c3_g8.err gpurun_out/r02w2_c3_g4.err; do echo "== $f"; tail -n 3 $f; done
for f in gpurun_out/r02w2_c3_g8.json gpurun_out/r02w2_c3_g4.json; do head -c 260 $f; echo; done
exit 0
