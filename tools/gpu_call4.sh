#!/bin/bash
O=gpurun_out
P="c2 c3 c4 w15 w11 w13 w5 ref30 w3 c2d32"
export STEREO_B200_LIB=$PWD/stereomatching_b200/libstereo_b200_dev.so
echo "== cap 4" > $O/c4_cap.log
python tools/exp_shapes.py $P --tm-only >> $O/c4_cap.log 2>&1
echo "== cap 3" >> $O/c4_cap.log
STEREO_B200_LIB=$PWD/stereomatching_b200/libstereo_b200_dev3.so python tools/exp_shapes.py $P --tm-only >> $O/c4_cap.log 2>&1
grep -v "direct kernel" $O/c4_cap.log
