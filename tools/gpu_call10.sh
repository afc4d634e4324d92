#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "virtual" ) > gpurun_out/pytest_gpu_k.txt 2>&1
tail -3 gpurun_out/pytest_gpu_k.txt
timeout 300 python tools/profile_pp_virtual.py 25 2 > gpurun_out/pp_virtual.txt 2>&1
timeout 300 python tools/profile_pp_virtual.py 13 3 >> gpurun_out/pp_virtual.txt 2>&1
cat gpurun_out/pp_virtual.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:contract_ws -c 1 -f -o gpurun_out/pp_gen_v176 python tools/profile_pp_virtual.py 13 1 > gpurun_out/ncu_pp_gen.log 2>&1
tail -2 gpurun_out/ncu_pp_gen.log
