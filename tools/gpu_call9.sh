#!/bin/bash
O=gpurun_out
python tools/gpu_dbg.py
python tools/latency.py > $O/c9_latency.log 2>&1; cat $O/c9_latency.log
STEREO_B200_LIB=$PWD/stereomatching_b200/libstereo_b200_dev.so python tools/exp_shapes.py c2 c3 c4 w15 w17 w21d64 ref30 w3 c2d32 > $O/c9_shapes.log 2>&1
grep -v "direct kernel" $O/c9_shapes.log
for f in 4-1920x1080 5-3840x2160; do for i in 1 2 3; do ./timing/stereopar tests/golden/imgs/$f/a.png tests/golden/imgs/$f/b.png; done; done 2>&1 | tee $O/c9_driver.log
