#!/bin/bash
N=${1:-8}
O=gpurun_out
bash tools/gpu_call12.sh $N
python tools/stage_times.py > $O/c13_stages.log 2>&1; cat $O/c13_stages.log
