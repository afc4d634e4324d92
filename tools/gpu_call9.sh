#!/bin/bash
# round-1 call 9: generated ladder with L2 window + per-warp arrive, new split-K cost model
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -6 gpurun_out/pytest_gpu.txt
timeout 300 python tools/profile_pp_virtual.py 25 2 > gpurun_out/pp_virtual.txt 2>&1
PMB_GEN_L2_PERSIST=0 timeout 300 python tools/profile_pp_virtual.py 25 2 >> gpurun_out/pp_virtual.txt 2>&1
cat gpurun_out/pp_virtual.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:contract_ws -c 1 -f -o gpurun_out/pp_gen_v176 python tools/profile_pp_virtual.py 13 1 > gpurun_out/ncu_pp_gen.log 2>&1
tail -3 gpurun_out/ncu_pp_gen.log
timeout 600 python tools/profile_sweep.py 25 > gpurun_out/sweep_profile_515.txt 2>&1
head -40 gpurun_out/sweep_profile_515.txt
