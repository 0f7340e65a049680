#!/bin/bash
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/c17_pytest.log 2>&1; tail -n 6 $O/c17_pytest.log
for rep in 1 2; do
for lib in tools/_base.so stereomatching_b200/libstereo_b200.so; do
echo "== $lib"; STEREO_B200_LIB=$PWD/$lib python tools/exp_shapes.py c2 c4 ref30 c3 --no-extra 2>&1 | grep -v "direct kernel"
done; done | tee $O/c17_ab.log
python tools/stage_times.py > $O/c17_stages.log 2>&1; cat $O/c17_stages.log
python tests/ladder.py 3 > $O/c17_ladder.md 2> $O/c17_ladder.err; cat $O/c17_ladder.md; tail -n 3 $O/c17_ladder.err
python tools/sweep_runs.py > $O/c17_runs.log 2>&1; cat $O/c17_runs.log
