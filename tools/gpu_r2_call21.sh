#!/bin/bash
# Round 2, call 21 (one GPU): the whole -m gpu suite with the gather path on by default (last GPU seconds of the round).
mkdir -p gpurun_out
( timeout 70 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_o.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_o.txt
