#!/bin/bash
# round-1 call 7: generated V_abcd -- parity tests, generated vs stored ladder, HBM kernels, bench at 515
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -15 gpurun_out/pytest_gpu.txt
timeout 300 python tools/profile_pp_virtual.py 18 2 dense > gpurun_out/pp_virtual.txt 2>&1
timeout 300 python tools/profile_pp_virtual.py 18 2 >> gpurun_out/pp_virtual.txt 2>&1
timeout 300 python tools/profile_pp_virtual.py 25 2 >> gpurun_out/pp_virtual.txt 2>&1
cat gpurun_out/pp_virtual.txt
timeout 300 python tools/bench_hbm.py > gpurun_out/hbm.txt 2>&1; tail -12 gpurun_out/hbm.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
tail -12 gpurun_out/bench_n1.log; cat gpurun_out/bench_n1.json
