#!/bin/bash
# round 2, GPU call w3 (8 GPUs): c3 at 8 GPUs with and without one slice of the host cores per rank
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nproc; lscpu | grep -i "model name\|^CPU(s)\|Thread\|NUMA node(s)"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29651 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02w3_aff.json 2> gpurun_out/r02w3_aff.err
echo "aff exit $?"
BENCH_NO_AFFINITY=1 timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29652 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02w3_noaff.json 2> gpurun_out/r02w3_noaff.err
echo "noaff exit $?"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29653 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r02w3_aff2.json 2> gpurun_out/r02w3_aff2.err
for f in gpurun_out/r02w3_aff.json gpurun_out/r02w3_noaff.json gpurun_out/r02w3_aff2.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), round(d['non_pass_ms_per_step'],3), d.get('host_affinity'))"; done
exit 0
