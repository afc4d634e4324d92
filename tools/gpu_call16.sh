#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 300 python tools/micro_edge.py 200 > gpurun_out/micro_edge.txt 2>&1; cat gpurun_out/micro_edge.txt
timeout 300 python tools/profile_pp_virtual.py 25 2 > gpurun_out/pp_virtual.txt 2>&1; cat gpurun_out/pp_virtual.txt
