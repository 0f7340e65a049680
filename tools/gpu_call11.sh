#!/bin/bash
# round 2, call 11: size ladder, ncu evidence for the window-21 kernels and the config-2 batch, config-5 sweep
O=gpurun_out
python tests/ladder.py 3 > $O/c11_ladder.md 2> $O/c11_ladder.err; cat $O/c11_ladder.md; tail -n 3 $O/c11_ladder.err
python tools/batch_fixture.py > $O/c11_batch_fixture.log 2>&1; cat $O/c11_batch_fixture.log
for p in c4 ref30 c2; do
python tools/exp_shapes.py $p --default-only > $O/c11_plain_$p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bitslice -s 3 -c 1 -o $O/r02_final_$p python tools/exp_shapes.py $p --default-only > $O/c11_ncu_$p.log 2>&1
tail -n 2 $O/c11_plain_$p.log
done
python tests/sweep_configs.py --what sweep --md $O/c11_sweep.md > $O/c11_sweep.jsonl 2> $O/c11_sweep.err; tail -n 70 $O/c11_sweep.md
ls -la $O/*.ncu-rep
