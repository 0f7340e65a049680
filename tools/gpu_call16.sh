#!/bin/bash
O=gpurun_out
for rep in 1 2; do
for lib in tools/_base.so tools/_unrolled.so stereomatching_b200/libstereo_b200.so; do
echo "== $lib"; STEREO_B200_LIB=$PWD/$lib python tools/exp_shapes.py c2 c4 ref30 --no-extra 2>&1 | grep -v "direct kernel"
done; done | tee $O/c16_ab.log
F=tests/golden/imgs/4-1920x1080
./timing/stereopar $F/a.png $F/b.png
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c16_driver_launches.csv ./timing/stereopar $F/a.png $F/b.png > $O/c16_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c16_driver_launches_ghost.csv ./timing/stereopar-ghost $F/a.png $F/b.png > $O/c16_ncu1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edges_planes -c 1 -o $O/r02_edges_planes ./timing/stereopar $F/a.png $F/b.png > $O/c16_ncu2.log 2>&1
python profiles/summarize.py launches $O/c16_driver_launches.csv; python profiles/summarize.py launches $O/c16_driver_launches_ghost.csv
python tools/sweep_runs.py > $O/c16_runs.log 2>&1; cat $O/c16_runs.log
