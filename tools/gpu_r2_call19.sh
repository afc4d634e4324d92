#!/bin/bash
# Round 2, call 19 (one GPU): the lock-step GCROT(m,k) of the FEAST / RT-EOM linear solves on the CUDA path.
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "feast or rt_eom" ) > gpurun_out/r2_pytest_gpu_feast.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu_feast.txt
