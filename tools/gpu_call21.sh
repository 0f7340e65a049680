#!/bin/bash
O=gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py tests/test_gpu_golden_big.py -m gpu -x -q -k "edge or fixture or f64 or diff or band or config or golden or batch") > $O/c21_pytest.log 2>&1; tail -n 6 $O/c21_pytest.log
python tools/stage_times.py > $O/c21_stages.log 2>&1; cat $O/c21_stages.log
F=tests/golden/imgs/4-1920x1080
for i in 1 2 3; do ./timing/stereopar $F/a.png $F/b.png; ./timing/stereopar-ghost $F/a.png $F/b.png; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c21_driver_launches.csv ./timing/stereopar $F/a.png $F/b.png > $O/c21_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c21_driver_launches_ghost.csv ./timing/stereopar-ghost $F/a.png $F/b.png > $O/c21_ncu1g.log 2>&1
python profiles/summarize.py launches $O/c21_driver_launches.csv; python profiles/summarize.py launches $O/c21_driver_launches_ghost.csv
