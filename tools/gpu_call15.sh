#!/bin/bash
O=gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -x -q -k "edge or fixture or f64 or diff or golden or config2 or band or half_word") > $O/c15_pytest.log 2>&1; tail -n 6 $O/c15_pytest.log
python tools/stage_times.py > $O/c15_stages.log 2>&1; cat $O/c15_stages.log
for rep in 1 2; do
for lib in tools/_base.so stereomatching_b200/libstereo_b200.so; do
echo "== $lib"; STEREO_B200_LIB=$PWD/$lib python tools/exp_shapes.py c2 c4 ref30 c3 w15 --no-extra 2>&1 | grep -v "direct kernel"
done; done | tee $O/c15_ab.log
python tools/sweep_runs.py > $O/c15_runs.log 2>&1; cat $O/c15_runs.log
python tools/gpu_dbg.py
