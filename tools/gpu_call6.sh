#!/bin/bash
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/c6_pytest.log 2>&1
tail -n 6 $O/c6_pytest.log
python tools/latency.py > $O/c6_latency.log 2>&1; cat $O/c6_latency.log
for f in 4-1920x1080 5-3840x2160; do for i in 1 2 3; do ./timing/stereopar tests/golden/imgs/$f/a.png tests/golden/imgs/$f/b.png; done; done 2>&1 | tee $O/c6_driver.log
python bench.py --no-cpu > $O/c6_bench.json 2> $O/c6_bench.err; tail -c 1500 $O/c6_bench.json; tail -n 3 $O/c6_bench.err
