#!/bin/bash
# round 2, last GPU call (1 GPU): clean-built final tree -- whole -m gpu suite, smoke(), default bench line
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02fin3_pytest.log 2>&1; echo "pytest exit $?"
tail -n 2 gpurun_out/r02fin3_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02fin3_smoke.log 2>&1; echo "smoke exit $?"
timeout -s KILL 900 python bench.py > gpurun_out/r02fin3_bench.json 2> gpurun_out/r02fin3_bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02fin3_bench.json').read().strip().splitlines()[-1]); e=d['e2e']; r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), round(d['non_pass_ms_per_step'],3), d['clocks']['sm_mhz'], 'e2e', round(e['value'],2), e['seconds_all_runs'], 'parity', d['parity']['ok'], 'roof', round(r['frac'],3), round(r['hbm_frac'],3), r['hbm_peak_source'], r['traffic'])"
exit 0
