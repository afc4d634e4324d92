#!/bin/bash
# round-1 call 8: generated V_abcd after the shared-memory index map / two-pass producer
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "virtual or realign or elementwise or ueg" ) > gpurun_out/pytest_gpu_k.txt 2>&1
tail -6 gpurun_out/pytest_gpu_k.txt
timeout 300 python tools/profile_pp_virtual.py 18 2 > gpurun_out/pp_virtual.txt 2>&1
timeout 300 python tools/profile_pp_virtual.py 25 2 >> gpurun_out/pp_virtual.txt 2>&1
cat gpurun_out/pp_virtual.txt
timeout 600 python tools/profile_sweep.py 25 > gpurun_out/sweep_profile_515.txt 2>&1
head -45 gpurun_out/sweep_profile_515.txt
