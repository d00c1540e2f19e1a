#!/bin/bash
# round 2, GPU call w4 (8 GPUs): the driver's command at 8 GPUs on the final tree (default flags: e2e + parity)
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02w4_c3_g8.json 2> gpurun_out/r02w4_c3_g8.err
echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02w4_c3_g8.json').read().strip().splitlines()[-1]); e=d['e2e']
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['non_pass_ms_per_step'],3), d['clocks'], 'e2e', round(e['value'],1), e['seconds_all_runs'], 'parity', d['parity']['ok'], d['parity']['worst'], 'cpu', d['cpu_baseline'])"
exit 0
