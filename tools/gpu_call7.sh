#!/bin/bash
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/c7_pytest.log 2>&1
tail -n 4 $O/c7_pytest.log
(time python bench.py --no-cpu) > $O/c7_bench.json 2> $O/c7_bench.err; tail -n 5 $O/c7_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c7_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'sustained',d['sustained']['value'],d['sustained']['clocks'])
print('e2e',d['e2e']['value'],d['e2e']['link']['frac_of_ceiling'],'i32',d['e2e']['with_i32_web']['value'])
print('roofline',{k:d['roofline'][k] for k in ('frac_throughput','frac_isolated','durations')})
print('c4',d['config4_pairs'])
print('c3',d['config3_bands'])
print('parity',d['parity'], 'clocks', d['clocks'], 'one pair', d['config']['one_pair_per_call']['hot_path_us'])
PY
