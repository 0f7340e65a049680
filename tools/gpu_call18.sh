#!/bin/bash
O=gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_big.py -m gpu -x -q) > $O/c18_pytest.log 2>&1; tail -n 6 $O/c18_pytest.log
for rep in 1 2; do
for lib in tools/_base.so tools/_new3.so stereomatching_b200/libstereo_b200.so; do
echo "== $lib"; STEREO_B200_LIB=$PWD/$lib python tools/exp_shapes.py c2 c4 ref30 c3 c2d32 --no-extra 2>&1 | grep -v "direct kernel"
done; done | tee $O/c18_ab.log
