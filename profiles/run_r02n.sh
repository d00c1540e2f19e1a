#!/bin/bash
# round 2, GPU call n (8 GPUs): c3 strong scaling at 8 GPUs (fused peer exchange / NCCL), BASELINE configs[3] (c4: N=256, T=5e7, non-ortho exp) at 8 GPUs
mkdir -p gpurun_out
export PICARD_TRACE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02n_c3_g8.json 2> gpurun_out/r02n_c3_g8.err
echo "c3 g8 exit $?" >> gpurun_out/r02n_c3_g8.err
PICARD_NO_P2P=1 timeout -s KILL 600 $TR --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02n_c3_g8_nccl.json 2> gpurun_out/r02n_c3_g8_nccl.err
echo "c3 g8 nccl exit $?" >> gpurun_out/r02n_c3_g8_nccl.err
timeout -s KILL 900 $TR --master-port 29623 bench.py --workload c4 --gpus 8 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02n_c4_g8.json 2> gpurun_out/r02n_c4_g8.err
echo "c4 g8 exit $?" >> gpurun_out/r02n_c4_g8.err
for f in gpurun_out/r02n_c3_g8.err gpurun_out/r02n_c3_g8_nccl.err gpurun_out/r02n_c4_g8.err; do echo "== $f"; tail -n 3 $f; done
for f in gpurun_out/r02n_c3_g8.json gpurun_out/r02n_c3_g8_nccl.json gpurun_out/r02n_c4_g8.json; do head -c 260 $f; echo; done
exit 0
