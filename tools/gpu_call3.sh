#!/bin/bash
O=gpurun_out
(time SMB_TM=1 python -m pytest tests -m gpu -x -q) > $O/c3_pytest_tm.log 2>&1
tail -n 15 $O/c3_pytest_tm.log
(time python -m pytest tests/test_gpu_golden_big.py -m gpu -x -q) > $O/c3_pytest_golden_default.log 2>&1
tail -n 8 $O/c3_pytest_golden_default.log
for p in c4; do
EXP_SHAPE=SMB_TM=1 python tools/exp_shapes.py $p --default-only > $O/c3_plain_$p.log 2>&1 && \
EXP_SHAPE=SMB_TM=1 ncu --set full --clock-control none --import-source on -k regex:k_bitslice -s 3 -c 1 -o $O/r02_tm_$p python tools/exp_shapes.py $p --default-only > $O/c3_ncu_$p.log 2>&1
tail -n 2 $O/c3_plain_$p.log
done
ls -la $O/*.ncu-rep
