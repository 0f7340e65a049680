#!/bin/bash
O=gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -x -q -k "edge or fixture or f64 or diff or band or config2 or golden") > $O/c19_pytest.log 2>&1; tail -n 6 $O/c19_pytest.log
for rep in 1 2; do
for lib in tools/_new5.so tools/_new6.so; do
echo "== $lib"; STEREO_B200_LIB=$PWD/$lib python tools/exp_shapes.py c2 c4 ref30 c3 --no-extra 2>&1 | grep -v "direct kernel"
done; done | tee $O/c19_ab.log
python tools/stage_times.py > $O/c19_stages.log 2>&1; cat $O/c19_stages.log
F=tests/golden/imgs/4-1920x1080
for i in 1 2 3; do ./timing/stereopar $F/a.png $F/b.png; ./timing/stereopar-ghost $F/a.png $F/b.png; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c19_driver_launches.csv ./timing/stereopar $F/a.png $F/b.png > $O/c19_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c19_driver_launches_ghost.csv ./timing/stereopar-ghost $F/a.png $F/b.png > $O/c19_ncu1g.log 2>&1
python profiles/summarize.py launches $O/c19_driver_launches.csv; python profiles/summarize.py launches $O/c19_driver_launches_ghost.csv
