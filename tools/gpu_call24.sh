#!/bin/bash
for rep in 1 2; do
python tools/edges_time.py
for v in b10_r4 b10_r3 b10_r2 b10_r1 b1_r4 b1_r2; do STEREO_B200_LIB=$PWD/tools/_edges_$v.so python tools/edges_time.py; done
done 2>&1 | tee gpurun_out/c24b_edges.log
