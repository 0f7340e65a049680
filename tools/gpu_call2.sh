#!/bin/bash
O=gpurun_out
timeout 600 python tools/exp_shapes.py > $O/c2_shapes.log 2>&1
tail -n 70 $O/c2_shapes.log
