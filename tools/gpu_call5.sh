#!/bin/bash
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/c5_pytest.log 2>&1
tail -n 6 $O/c5_pytest.log
export STEREO_B200_LIB=$PWD/stereomatching_b200/libstereo_b200_dev.so
python tools/exp_shapes.py c2 c3 c4 w15 w17 w21d64 ref30 w3 c2d32 --no-extra > $O/c5_shapes.log 2>&1
grep -v "direct kernel" $O/c5_shapes.log
