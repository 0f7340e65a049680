#!/bin/bash
(time python tests/fuzz_gpu.py 3000 777) > gpurun_out/c32_fuzz.log 2>&1; tail -n 6 gpurun_out/c32_fuzz.log
