#!/bin/bash
# the round's measurement set on one B200 (run under gpurun); everything lands in gpurun_out/ (prefix f_)
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/f_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1
python bench.py > $O/f_bench.json 2> $O/f_bench.err
python bench.py --impl reference > $O/f_ref.json 2> $O/f_ref.err
ARGS="--steps 2 --warmup 3 --pairs 32 --distinct 2 --e2e-pairs 16 --no-cpu --no-extra"
python bench.py $ARGS > $O/f_plain.json 2> $O/f_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/f_launches.csv python bench.py $ARGS > $O/f_ncu1.log 2>&1
# --set full captures are summarised HERE and deleted: gpurun brings back at most 64 MiB and one report is 26 MB
for p in c2 c4 ref30; do
python tools/exp_shapes.py $p --default-only > $O/f_plain_$p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bitslice -s 3 -c 1 -o $O/f_ncu_$p python tools/exp_shapes.py $p --default-only > $O/f_ncu_$p.log 2>&1
python profiles/summarize.py full $O/f_ncu_$p.ncu-rep > $O/f_ncu_$p.md 2>> $O/f_ncu_$p.log
if [ $p = c2 ]; then
  python profiles/make_traffic.py $O/f_ncu_c2.ncu-rep > $O/f_traffic.json 2>> $O/f_ncu_$p.log
  python tools/sass_regions.py $O/f_ncu_c2.ncu-rep > $O/f_sass_regions_c2.md 2>> $O/f_ncu_$p.log
fi
rm -f $O/f_ncu_$p.ncu-rep
done
F=tests/golden/imgs/4-1920x1080
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/f_driver_launches.csv ./timing/stereopar $F/a.png $F/b.png > $O/f_ncu_drv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edges_planes -c 1 -o $O/f_ncu_edges ./timing/stereopar $F/a.png $F/b.png > $O/f_ncu_edges.log 2>&1
python profiles/summarize.py full $O/f_ncu_edges.ncu-rep > $O/f_ncu_edges.md 2>> $O/f_ncu_edges.log; rm -f $O/f_ncu_edges.ncu-rep
python tools/stage_times.py > $O/f_stages.log 2>&1
python tests/sweep_configs.py --what c1,c2,c3,c4,sweep --md $O/f_sweep.md > $O/f_sweep.jsonl 2> $O/f_sweep.err
python tools/exp_shapes.py c2 c4 ref30 c3 w15 w17 c2d32 c2d16 d16w21 small --no-extra 2>&1 | grep -v "direct kernel" > $O/f_shapes.log
python tools/batch_fixture.py > $O/f_batch_fixture.log 2>&1
python tests/ladder.py 3 > $O/f_ladder.md 2> $O/f_ladder.err
tail -3 $O/f_pytest.log; cat $O/f_smoke.log; tail -n 2 $O/f_bench.err $O/f_ref.err $O/f_sweep.err; cat $O/f_shapes.log $O/f_batch_fixture.log; cat $O/f_stages.log; python profiles/summarize.py launches $O/f_driver_launches.csv
