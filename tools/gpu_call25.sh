#!/bin/bash
O=gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py tests/test_gpu_golden_big.py -m gpu -x -q -k "edge or fixture or f64 or diff or band or config or golden or batch or fuzz") > $O/c25_pytest.log 2>&1; tail -n 6 $O/c25_pytest.log
python tools/edges_time.py; python tools/edges_time.py
F=tests/golden/imgs/4-1920x1080
for i in 1 2 3; do ./timing/stereopar $F/a.png $F/b.png; ./timing/stereopar-ghost $F/a.png $F/b.png; done
