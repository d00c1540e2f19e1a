#!/bin/bash
# round 2, GPU call m (2 GPUs): exchange fused into the pass kernels' tails -- multi-GPU tests, c3 bench at 2 GPUs (fused / NCCL), 1-GPU i8 tests
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_multigpu_gpu.py tests/test_i8_gpu.py -q -m gpu > gpurun_out/r02m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02m_pytest.log
export PICARD_TRACE=1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02m_bench_p2p.json 2> gpurun_out/r02m_bench_p2p.err
echo "bench p2p exit $?" >> gpurun_out/r02m_bench_p2p.err
PICARD_NO_P2P=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02m_bench_nccl.json 2> gpurun_out/r02m_bench_nccl.err
echo "bench nccl exit $?" >> gpurun_out/r02m_bench_nccl.err
for f in gpurun_out/r02m_pytest.log gpurun_out/r02m_bench_p2p.err gpurun_out/r02m_bench_nccl.err; do echo "== $f"; tail -n 3 $f; done
head -c 300 gpurun_out/r02m_bench_p2p.json; echo; head -c 300 gpurun_out/r02m_bench_nccl.json
exit 0
