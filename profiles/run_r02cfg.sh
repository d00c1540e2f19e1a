#!/bin/bash
# round 2, GPU call cfg (1 GPU): the other BASELINE configs that fit one GPU through bench.py, for the record
mkdir -p gpurun_out
for c in c1 c2 c5; do
  timeout -s KILL 600 python bench.py --workload $c --no-cpu > gpurun_out/r02cfg_$c.json 2> gpurun_out/r02cfg_$c.err; echo "$c exit $?"
  python -c "
import json; d=json.loads(open('gpurun_out/r02cfg_$c.json').read().strip().splitlines()[-1]); e=d.get('e2e') or {}; p=d.get('parity') or {}
print('$c', round(d['value'],1), round(d['ms_per_step'],3), round(d['non_pass_ms_per_step'],3), 'e2e', e.get('value'), e.get('seconds_all_runs'), 'parity', p.get('ok'), p.get('worst'))"
done
exit 0
