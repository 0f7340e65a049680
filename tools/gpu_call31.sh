#!/bin/bash
STEREO_B200_LIB=$PWD/stereomatching_b200/libstereo_b200_dev.so python tools/exp_shapes.py c2 c4 ref30 c3 c2d16 2>&1 | grep -v "direct kernel" | tee gpurun_out/c31_tr.log
