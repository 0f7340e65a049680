#!/bin/bash
bash tools/gpu_call12.sh 8 2>&1 | tail -n 7 | cut -c1-900
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -n 3
