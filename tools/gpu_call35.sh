#!/bin/bash
O=gpurun_out
(time python bench.py) > $O/g_bench.json 2> $O/g_bench.err; tail -n 4 $O/g_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/g_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['link']['frac_of_ceiling'],d['e2e']['pairs_per_step'],'i32',d['e2e']['with_i32_web']['value'],d['e2e']['with_i32_web']['frac_of_ceiling'])
print('c4',d['config4_pairs']['resident']['value'],d['config4_pairs']['e2e'],d['config4_pairs']['parity']['equal_reference_golden'])
print('c3',d['config3_bands']['resident'],d['config3_bands']['parity']['bands_equal_reference_golden'],'parity',d['parity'])
PY
