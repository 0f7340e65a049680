#!/bin/bash
O=gpurun_out
(time python bench.py --no-cpu) > $O/c8_bench.json 2> $O/c8_bench.err; tail -n 5 $O/c8_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c8_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'sustained',d['sustained']['value'],d['sustained']['clocks'])
print('e2e',d['e2e']['value'],d['e2e']['link']['frac_of_ceiling'],'i32',d['e2e']['with_i32_web']['value'])
print('roofline',{k:d['roofline'][k] for k in ('frac_throughput','frac_isolated','durations')})
print('c4',d['config4_pairs'])
print('c3',d['config3_bands'])
print('parity',d['parity'], 'clocks', d['clocks'], 'one pair', d['config']['one_pair_per_call']['hot_path_us'])
PY
ARGS="--steps 2 --warmup 3 --pairs 32 --distinct 2 --e2e-pairs 16 --no-cpu --no-extra"
python bench.py $ARGS > $O/c8_plain.json 2> $O/c8_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py $ARGS > $O/c8_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bitslice -s 6 -c 1 -o $O/r02_c2_batch python bench.py $ARGS > $O/c8_ncu2.log 2>&1
ls -la $O/*.ncu-rep $O/r02_launches.csv
