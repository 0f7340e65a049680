#!/bin/bash
# the round's measurement set on one B200 (run under gpurun); everything lands in gpurun_out/
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/f_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1
python bench.py > $O/f_bench.json 2> $O/f_bench.err
python bench.py --impl reference > $O/f_ref.json 2> $O/f_ref.err
ARGS="--steps 2 --warmup 3 --pairs 32 --distinct 2 --e2e-pairs 16 --no-cpu"
python bench.py $ARGS > $O/f_plain.json 2> $O/f_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/f_launches.csv python bench.py $ARGS > $O/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bitslice|k_pack" -s 6 -c 6 -o $O/prof_final python bench.py $ARGS > $O/f_ncu2.log 2>&1
python tests/sweep_configs.py --what c1,c2,c3,c4,sweep --md $O/f_sweep.md > $O/f_sweep.jsonl 2> $O/f_sweep.err
tail -3 $O/f_pytest.log; cat $O/f_smoke.log; tail -n 2 $O/f_bench.err $O/f_ref.err $O/f_sweep.err; ls -la $O/prof_final.ncu-rep
