#!/bin/bash
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/c14_pytest.log 2>&1; tail -n 15 $O/c14_pytest.log
python tools/stage_times.py > $O/c14_stages.log 2>&1; cat $O/c14_stages.log
python tests/ladder.py 2 > $O/c14_ladder.md 2> $O/c14_ladder.err; cat $O/c14_ladder.md; tail -n 3 $O/c14_ladder.err
python tests/sweep_configs.py --what sweep --md $O/c14_sweep.md > $O/c14_sweep.jsonl 2> $O/c14_sweep.err; head -n 24 $O/c14_sweep.md; tail -n 3 $O/c14_sweep.err
(time python bench.py --no-cpu) > $O/c14_bench.json 2> $O/c14_bench.err; tail -n 5 $O/c14_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c14_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'sustained',d['sustained']['value'])
print('e2e',d['e2e']['value'],d['e2e']['link'],'i32',d['e2e']['with_i32_web'])
print('roofline',{k:d['roofline'][k] for k in ('frac_throughput','frac_isolated','durations','issue','alu_mix')})
print('c4',d['config4_pairs']['resident'],d['config4_pairs']['e2e']['value'])
print('c3',d['config3_bands']['resident'],d['config3_bands']['parity'])
print('parity',d['parity'], 'one pair', d['config']['one_pair_per_call']['hot_path_us'])
PY
