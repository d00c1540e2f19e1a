#!/bin/bash
# round 2, GPU call v (2 GPUs): multi-GPU tests (with the c3 / c4 shaped two-GPU parity cases), c3 bench at 2 GPUs after the stage-release fix
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > gpurun_out/r02v_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02v_pytest.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02v_bench_g2.json 2> gpurun_out/r02v_bench_g2.err
echo "bench exit $?" >> gpurun_out/r02v_bench_g2.err
for f in gpurun_out/r02v_pytest.log gpurun_out/r02v_bench_g2.err; do echo "== $f"; tail -n 6 $f; done
head -c 400 gpurun_out/r02v_bench_g2.json; echo
exit 0
