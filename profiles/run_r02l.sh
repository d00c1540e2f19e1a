#!/bin/bash
# round 2, GPU call l (2 GPUs): peer-memory allreduce vs NCCL -- multi-GPU tests, c3 bench at 2 GPUs both ways
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_multigpu_gpu.py tests/test_i8_gpu.py::test_second_device_in_one_process -q -m gpu > gpurun_out/r02l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02l_pytest.log
export PICARD_TRACE=1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02l_bench_p2p.json 2> gpurun_out/r02l_bench_p2p.err
echo "bench p2p exit $?" >> gpurun_out/r02l_bench_p2p.err
PICARD_NO_P2P=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02l_bench_nccl.json 2> gpurun_out/r02l_bench_nccl.err
echo "bench nccl exit $?" >> gpurun_out/r02l_bench_nccl.err
for f in gpurun_out/r02l_pytest.log gpurun_out/r02l_bench_p2p.err gpurun_out/r02l_bench_nccl.err; do echo "== $f"; tail -n 5 $f; done
grep -h "peer-memory" gpurun_out/r02l_bench_p2p.err | head -2
head -c 300 gpurun_out/r02l_bench_p2p.json; echo; head -c 300 gpurun_out/r02l_bench_nccl.json
exit 0
