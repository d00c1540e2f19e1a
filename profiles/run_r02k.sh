#!/bin/bash
# round 2, GPU call k: fused LOSS kernel (W' digits in the prologue, reduction + loss + publish in the tail), fused reduction tail
# of the gradient kernel, host polling of mapped scalars; lab, full GPU suite, bench, launch list
mkdir -p gpurun_out
timeout -s KILL 120 profiles/lab/i8_lab 2000000 5 > gpurun_out/r02k_lab.jsonl 2> gpurun_out/r02k_lab.err
echo "lab exit $?" >> gpurun_out/r02k_lab.err
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench exit $?" >> gpurun_out/r02k_bench.err
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02k_launches.csv $CMD > gpurun_out/r02k_ncu_list.log 2>&1
for f in gpurun_out/r02k_lab.err gpurun_out/r02k_pytest.log gpurun_out/r02k_bench.err; do echo "== $f"; tail -n 4 $f; done
grep -E "\"kernel\"|trace" gpurun_out/r02k_lab.jsonl | cut -c1-200
head -c 400 gpurun_out/r02k_bench.json
exit 0
