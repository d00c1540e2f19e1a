#!/bin/bash
# round 2, final GPU call (1 GPU): whole -m gpu suite, smoke(), the default bench line, then the evidence of the final build --
# launch list of the bench command and ncu --set full of the pass kernels and of the N x N kernels between them
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02fin_pytest.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/r02fin_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02fin_smoke.log 2>&1; echo "smoke exit $?"
timeout -s KILL 900 python bench.py > gpurun_out/r02fin_bench.json 2> gpurun_out/r02fin_bench.err; echo "bench exit $?"
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-parity"
timeout -s KILL 300 $CMD > gpurun_out/r02fin_plain.json 2> gpurun_out/r02fin_plain.err || exit 1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02fin.csv $CMD > gpurun_out/r02fin_ncu_list.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'loss_i8_kernel|grad_i8_kernel|front_cluster_kernel|expm_multi_kernel' --launch-skip 12 -c 8 -o gpurun_out/prof_r02fin $CMD > gpurun_out/r02fin_ncu_full.log 2>&1
ls -la gpurun_out/prof_r02fin.ncu-rep gpurun_out/launches_r02fin.csv
tail -n 2 gpurun_out/r02fin_ncu_full.log
exit 0
