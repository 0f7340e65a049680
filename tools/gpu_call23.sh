#!/bin/bash
O=gpurun_out
for N in 2 4 8; do bash tools/gpu_call12.sh $N 2>&1 | tail -n 8 | cut -c1-1500; done
python bench_bands.py --help > /dev/null 2>&1
./timing/stereobatch 2>&1 | tail -n 3
