#!/bin/bash
# round 2, GPU call 1: shape experiment for the 5-plane windows + baseline ncu evidence
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt
python tools/exp_shapes.py > $O/c1_shapes.log 2>&1
tail -n 60 $O/c1_shapes.log
python tools/exp_shapes.py c4 --default-only > $O/c1_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bitslice -s 3 -c 2 -o $O/r02_base_c4 python tools/exp_shapes.py c4 --default-only > $O/c1_ncu_c4.log 2>&1
python tools/exp_shapes.py ref30 --default-only > $O/c1_plain_ref30.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bitslice -s 3 -c 2 -o $O/r02_base_ref30 python tools/exp_shapes.py ref30 --default-only > $O/c1_ncu_ref30.log 2>&1
ls -la $O/*.ncu-rep
