#!/bin/bash
# round 2, GPU call x2: one-barrier Taylor loop of the transform kernels -- N x N kernel tests, fits, device time of the transforms in the loop
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_small_gpu.py tests/test_point_gpu.py tests/test_fit_gpu.py tests/test_i8_gpu.py -x -q -m gpu > gpurun_out/r02x2_pytest.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/r02x2_pytest.log
PICARD_TRACE_GAPS=1 timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02x2_bench.json 2> gpurun_out/r02x2_bench.err; echo "bench exit $?"
tail -n 6 gpurun_out/r02x2_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02x2_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['non_pass_ms_per_step'], d['passes']['loss'])"
exit 0
